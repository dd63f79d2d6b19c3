"""SURVEY 8(f) rows on the device: the functions, cost heads and graph combinators that sit either side of the
dense path, with the reference's names and semantics.

    math_funcs.go   Log :99  LogSigmoid :376  Softmax :479  Sin :511  Cos :599  Norm :201   (+ Tanh, an extension)
    arithmetic.go   AddScaler :184  Div :378  ScaleFirst :495  AddFirst :595  Inverse :772  Pow :869
                    AddLogDomain :1056  SumAllLogDomain :1077
    pool.go         Pool / PoolAll (+R) :21-195        fold.go   Fold / FoldR :19-178
    lin_alg.go      MatMulVec(s) (+R) :14-133          slices.go Split / SplitR :310-335
    cost head       squared_error = Scale(SquaredNorm(Sub(a, y)), .5) as one reduction (north_star "SquaredError")

Everything stays in HBM: pool variables wrap the producing Result's device buffer (`Variable.from_device`), the
temp-variable protocol runs on the backward session's device accumulators (`_Session.take`), and scalars the
reference reads on the host (ScaleFirst's `scaler.Output()[0]`, `MaxAbs()`) are passed to the kernels by device
pointer, so none of these ops synchronises the stream.  No CPU fallback: every op is a call into the C ABI."""
from __future__ import annotations

import math

from . import _lib
from .api import (DeviceVec, Exp, RVariable, ShapeError, Variable, _BinaryResult, _DeviceResult, _down, _down_r,
                  _Elementwise, _in_d, _in_dr, _in_len, _p, _same_len, _TempGrad, _UnaryResult, _zeros_if_none, add,
                  add_r, check, context, mul_r, sum_all, sum_all_r, LinTran)
from .slices import slice_, slice_r


def result_len(r) -> int:
    return _in_len(r)


# --------------------------------------------------------------------------------------------
# elementwise functions whose backward needs the INPUT (Log, LogSigmoid, Sin) or the OUTPUT (Tanh)
# --------------------------------------------------------------------------------------------
class _EwFunc(_Elementwise):
    fwd = bwd = ""
    uses_output = False  # backward reads the forward outputs (T, RT) instead of the inputs (X, RX)

    def apply(self, inp):
        ctx = context()
        x = _in_d(inp)
        o = ctx.empty(x.n)
        check(getattr(ctx.lib, self.fwd)(ctx.h, x.ptr, None, o.ptr, None, x.n))
        return _EwResult(self, inp, o)

    def apply_r(self, rv, inp):
        ctx = context()
        x, xr = _in_d(inp), _in_dr(inp)
        o, ro = ctx.empty(x.n), ctx.empty(x.n)
        check(getattr(ctx.lib, self.fwd)(ctx.h, x.ptr, _p(xr), o.ptr, ro.ptr, x.n))
        return _EwResult(self, inp, o, ro, True)


class _EwResult(_UnaryResult):
    def __init__(self, fn, inp, d, dr=None, is_r=False):
        super().__init__(inp, d, dr, is_r)
        self.fn = fn

    def _bwd(self, sess, dU, grad):
        if self.input.constant(grad):
            return
        ctx = context()
        src = self.d if self.fn.uses_output else _in_d(self.input)
        check(getattr(ctx.lib, self.fn.bwd)(ctx.h, src.ptr, None, dU.ptr, None, dU.n))
        _down(self.input, sess, dU, grad)

    def _bwd_r(self, sess, dU, dUR, rgrad, grad):
        if self.input.constant(rgrad, grad):
            return
        ctx = context()
        if self.fn.uses_output:
            src, srcr = self.d, self.dr
        else:
            src, srcr = _in_d(self.input), _in_dr(self.input)
        check(getattr(ctx.lib, self.fn.bwd)(ctx.h, src.ptr, _p(srcr), dU.ptr, dUR.ptr, dU.n))
        _down_r(self.input, sess, dU, dUR, rgrad, grad)


class Log(_EwFunc):  # math_funcs.go:99-189
    fwd, bwd = "af_log_fwd_r", "af_log_bwd_r"


class LogSigmoid(_EwFunc):  # math_funcs.go:376-477
    fwd, bwd = "af_logsigmoid_fwd_r", "af_logsigmoid_bwd_r"


class Sin(_EwFunc):  # math_funcs.go:511-597
    fwd, bwd = "af_sin_fwd_r", "af_sin_bwd_r"


class Tanh(_EwFunc):
    """EXTENSION: named by BASELINE.json's north_star, absent from the reference."""
    fwd, bwd, uses_output = "af_tanh_fwd_r", "af_tanh_bwd_r", True


class Cos(_Elementwise):  # math_funcs.go:599-607
    def apply(self, inp):
        return Sin().apply(add_scaler(inp, math.pi / 2))

    def apply_r(self, rv, inp):
        return Sin().apply_r(rv, add_scaler_r(inp, math.pi / 2))


# --------------------------------------------------------------------------------------------
# AddScaler, Inverse, Pow, Div
# --------------------------------------------------------------------------------------------
class _PassThrough(_UnaryResult):  # arithmetic.go:212-214,248-252: upstream goes through untouched
    def _bwd(self, sess, dU, grad):
        if not self.input.constant(grad):
            _down(self.input, sess, dU, grad)

    def _bwd_r(self, sess, dU, dUR, rgrad, grad):
        if not self.input.constant(rgrad, grad):
            _down_r(self.input, sess, dU, dUR, rgrad, grad)


def _add_scaler_d(r, f):
    ctx = context()
    x = _in_d(r)
    o = ctx.empty(x.n)
    check(ctx.lib.af_add_scaler(ctx.h, x.ptr, float(f), o.ptr, x.n))
    return o


def add_scaler(r, f):  # arithmetic.go:184-214
    return _PassThrough(r, _add_scaler_d(r, f))


def add_scaler_r(r, f):  # arithmetic.go:216-252 (ROutput is the input's)
    return _PassThrough(r, _add_scaler_d(r, f), _in_dr(r), True)


def _add_dev_d(r, d_f: DeviceVec, sign: float):
    ctx = context()
    x = _in_d(r)
    o = ctx.empty(x.n)
    check(ctx.lib.af_add_dev(ctx.h, x.ptr, d_f.ptr, float(sign), o.ptr, x.n))
    return o


class _Inverse(_UnaryResult):
    def _bwd(self, sess, dU, grad):  # arithmetic.go:800-807
        if not self.input.constant(grad):
            ctx = context()
            check(ctx.lib.af_inverse_bwd_r(ctx.h, self.d.ptr, None, dU.ptr, None, dU.n))
            _down(self.input, sess, dU, grad)

    def _bwd_r(self, sess, dU, dUR, rgrad, grad):  # arithmetic.go:851-867
        if not self.input.constant(rgrad, grad):
            ctx = context()
            check(ctx.lib.af_inverse_bwd_r(ctx.h, self.d.ptr, _p(_in_dr(self.input)), dU.ptr, dUR.ptr, dU.n))
            _down_r(self.input, sess, dU, dUR, rgrad, grad)


def inverse(r):  # arithmetic.go:779-790
    ctx = context()
    x = _in_d(r)
    o = ctx.empty(x.n)
    check(ctx.lib.af_inverse_fwd_r(ctx.h, x.ptr, None, o.ptr, None, x.n))
    return _Inverse(r, o)


def inverse_r(r):  # arithmetic.go:817-837
    ctx = context()
    x = _in_d(r)
    o, ro = ctx.empty(x.n), ctx.empty(x.n)
    check(ctx.lib.af_inverse_fwd_r(ctx.h, x.ptr, _p(_in_dr(r)), o.ptr, ro.ptr, x.n))
    return _Inverse(r, o, ro, True)


class _Pow(_UnaryResult):
    power = 1.0

    def constant(self, *maps):  # arithmetic.go:892-894,944-946
        return self.power != 0 and self.input.constant(*maps)

    def _bwd(self, sess, dU, grad):  # arithmetic.go:896-906
        if not self.constant(grad):
            ctx = context()
            check(ctx.lib.af_pow_bwd_r(ctx.h, _in_d(self.input).ptr, None, self.power, dU.ptr, None, dU.n))
            _down(self.input, sess, dU, grad)

    def _bwd_r(self, sess, dU, dUR, rgrad, grad):  # arithmetic.go:948-966
        if not self.constant(rgrad, grad):
            ctx = context()
            check(ctx.lib.af_pow_bwd_r(ctx.h, _in_d(self.input).ptr, _p(_in_dr(self.input)), self.power, dU.ptr,
                                       dUR.ptr, dU.n))
            _down_r(self.input, sess, dU, dUR, rgrad, grad)


def pow_(r, p):  # arithmetic.go:876-886
    ctx = context()
    x = _in_d(r)
    o = ctx.empty(x.n)
    check(ctx.lib.af_pow_fwd_r(ctx.h, x.ptr, None, float(p), o.ptr, None, x.n))
    res = _Pow(r, o)
    res.power = float(p)
    return res


def pow_r(r, p):  # arithmetic.go:916-937
    ctx = context()
    x = _in_d(r)
    o, ro = ctx.empty(x.n), ctx.empty(x.n)
    check(ctx.lib.af_pow_fwd_r(ctx.h, x.ptr, _p(_in_dr(r)), float(p), o.ptr, ro.ptr, x.n))
    res = _Pow(r, o, ro, True)
    res.power = float(p)
    return res


class _Quotient(_BinaryResult):
    def _bwd(self, sess, dU, grad):  # arithmetic.go:407-423
        ctx = context()
        na, nb = not self.r1.constant(grad), not self.r2.constant(grad)
        if not (na or nb):
            return
        da = ctx.empty(dU.n) if na else None
        db = ctx.empty(dU.n) if nb else None
        check(ctx.lib.af_div_bwd(ctx.h, dU.ptr, self.d.ptr, _in_d(self.r2).ptr, _p(da), _p(db), dU.n))
        if na:
            _down(self.r1, sess, da, grad)
        if nb:
            _down(self.r2, sess, db, grad)


def div(a, b):  # arithmetic.go:385-397
    _same_len(a, b)
    ctx = context()
    x, y = _in_d(a), _in_d(b)
    o = ctx.empty(x.n)
    check(ctx.lib.af_div(ctx.h, x.ptr, y.ptr, o.ptr, x.n))
    return _Quotient(a, b, o)


def div_r(a, b):  # arithmetic.go:425-427
    return mul_r(a, inverse_r(b))


# --------------------------------------------------------------------------------------------
# ScaleFirst, AddFirst: broadcast of a computed scalar (read on the device)
# --------------------------------------------------------------------------------------------
class _ScaleFirst(_DeviceResult):
    def __init__(self, inp, scaler, d, dr=None):
        super().__init__(d, dr, dr is not None)
        self.input, self.scaler = inp, scaler

    def constant(self, *maps):
        return self.input.constant(*maps) and self.scaler.constant(*maps)

    def _bwd(self, sess, dU, grad):  # arithmetic.go:520-533
        ctx = context()
        if not self.scaler.constant(grad):
            down = ctx.zeros(_in_len(self.scaler))
            check(ctx.lib.af_dot(ctx.h, dU.ptr, _in_d(self.input).ptr, None, None, down.ptr, dU.n))
            _down(self.scaler, sess, down, grad)
        if not self.input.constant(grad):  # after the scaler's branch, so the upstream can be scaled in place
            check(ctx.lib.af_scale_dev(ctx.h, dU.ptr, None, _in_d(self.scaler).ptr, None, dU.ptr, None, dU.n))
            _down(self.input, sess, dU, grad)

    def _bwd_r(self, sess, dU, dUR, rgrad, grad):  # arithmetic.go:567-593
        ctx = context()
        x, xr = _in_d(self.input), _in_dr(self.input)
        if not self.scaler.constant(rgrad, grad):
            m = _in_len(self.scaler)
            down, down_r = ctx.zeros(m), ctx.zeros(m)
            check(ctx.lib.af_dot(ctx.h, dU.ptr, x.ptr, None, None, down.ptr, dU.n))
            if xr is not None:
                check(ctx.lib.af_dot(ctx.h, dUR.ptr, x.ptr, dU.ptr, xr.ptr, down_r.ptr, dU.n))
            else:
                check(ctx.lib.af_dot(ctx.h, dUR.ptr, x.ptr, None, None, down_r.ptr, dU.n))
            _down_r(self.scaler, sess, down, down_r, rgrad, grad)
        if not self.input.constant(rgrad, grad):
            check(ctx.lib.af_scale_dev(ctx.h, dU.ptr, dUR.ptr, _in_d(self.scaler).ptr, _p(_in_dr(self.scaler)),
                                       dU.ptr, dUR.ptr, dU.n))
            _down_r(self.input, sess, dU, dUR, rgrad, grad)


def scale_first(inp, scaler):  # arithmetic.go:503-510
    ctx = context()
    x = _in_d(inp)
    o = ctx.empty(x.n)
    check(ctx.lib.af_scale_dev(ctx.h, x.ptr, None, _in_d(scaler).ptr, None, o.ptr, None, x.n))
    return _ScaleFirst(inp, scaler, o)


def scale_first_r(inp, scaler):  # arithmetic.go:544-553
    ctx = context()
    x = _in_d(inp)
    o, ro = ctx.empty(x.n), ctx.empty(x.n)
    check(ctx.lib.af_scale_dev(ctx.h, x.ptr, _p(_in_dr(inp)), _in_d(scaler).ptr, _p(_in_dr(scaler)), o.ptr, ro.ptr,
                               x.n))
    return _ScaleFirst(inp, scaler, o, ro)


class _AddFirst(_DeviceResult):
    def __init__(self, inp, scaler, d, dr=None):
        super().__init__(d, dr, dr is not None)
        self.input, self.scaler = inp, scaler

    def constant(self, *maps):
        return self.input.constant(*maps) and self.scaler.constant(*maps)

    def _bwd(self, sess, dU, grad):  # arithmetic.go:627-639
        ctx = context()
        if not self.scaler.constant(grad):
            su = ctx.zeros(_in_len(self.scaler))
            check(ctx.lib.af_sum_all(ctx.h, dU.ptr, None, su.ptr, None, dU.n))
            _down(self.scaler, sess, su, grad)
        if not self.input.constant(grad):
            _down(self.input, sess, dU, grad)

    def _bwd_r(self, sess, dU, dUR, rgrad, grad):  # arithmetic.go:680-694
        ctx = context()
        if not self.scaler.constant(rgrad, grad):
            m = _in_len(self.scaler)
            su, sur = ctx.zeros(m), ctx.zeros(m)
            check(ctx.lib.af_sum_all(ctx.h, dU.ptr, dUR.ptr, su.ptr, sur.ptr, dU.n))
            _down_r(self.scaler, sess, su, sur, rgrad, grad)
        if not self.input.constant(rgrad, grad):
            _down_r(self.input, sess, dU, dUR, rgrad, grad)


def add_first(v1, v2):  # arithmetic.go:607-619
    return _AddFirst(v1, v2, _add_dev_d(v1, _in_d(v2), 1.0))


def add_first_r(v1, v2):  # arithmetic.go:649-667
    ctx = context()
    n = _in_len(v1)
    xr, fr = _zeros_if_none(_in_dr(v1), n), _in_dr(v2)
    if fr is None:
        ro = xr.copy()
    else:
        ro = ctx.empty(n)
        check(ctx.lib.af_add_dev(ctx.h, xr.ptr, fr.ptr, 1.0, ro.ptr, n))
    return _AddFirst(v1, v2, _add_dev_d(v1, _in_d(v2), 1.0), ro)


# --------------------------------------------------------------------------------------------
# log-domain sums (arithmetic.go:1049-1090); maxVal = max |x| is a constant of the graph, kept on the device
# --------------------------------------------------------------------------------------------
def _max_abs(*rs) -> DeviceVec:
    ctx = context()
    m = ctx.zeros(1)
    for i, r in enumerate(rs):
        x = _in_d(r)
        check(ctx.lib.af_max_abs(ctx.h, x.ptr, m.ptr, int(i > 0), x.n))
    return m


def _shift(r, m, sign, with_r):
    return _PassThrough(r, _add_dev_d(r, m, sign), _in_dr(r) if with_r else None, with_r)


def add_log_domain(v1, v2):  # arithmetic.go:1056-1062
    _same_len(v1, v2)
    m = _max_abs(v1, v2)
    e1, e2 = Exp().apply(_shift(v1, m, -1.0, False)), Exp().apply(_shift(v2, m, -1.0, False))
    return _shift(Log().apply(add(e1, e2)), m, 1.0, False)


def add_log_domain_r(v1, v2):  # arithmetic.go:1065-1072
    _same_len(v1, v2)
    m = _max_abs(v1, v2)
    e1, e2 = Exp().apply_r({}, _shift(v1, m, -1.0, True)), Exp().apply_r({}, _shift(v2, m, -1.0, True))
    return _shift(Log().apply_r({}, add_r(e1, e2)), m, 1.0, True)


def sum_all_log_domain(v):  # arithmetic.go:1077-1082
    m = _max_abs(v)
    return _shift(Log().apply(sum_all(Exp().apply(_shift(v, m, -1.0, False)))), m, 1.0, False)


def sum_all_log_domain_r(v):  # arithmetic.go:1085-1090
    m = _max_abs(v)
    return _shift(Log().apply_r(None, sum_all_r(Exp().apply_r(None, _shift(v, m, -1.0, True)))), m, 1.0, True)


# --------------------------------------------------------------------------------------------
# Norm (math_funcs.go:201-284)
# --------------------------------------------------------------------------------------------
class Norm(_Elementwise):
    def apply(self, inp):
        ctx = context()
        x = _in_d(inp)
        o = ctx.empty(1)
        check(ctx.lib.af_norm_fwd_r(ctx.h, x.ptr, None, o.ptr, None, x.n))
        return _NormResult(inp, o)

    def apply_r(self, rv, inp):
        ctx = context()
        x = _in_d(inp)
        o, ro = ctx.empty(1), ctx.empty(1)
        check(ctx.lib.af_norm_fwd_r(ctx.h, x.ptr, _p(_in_dr(inp)), o.ptr, ro.ptr, x.n))
        return _NormResult(inp, o, ro, True)


class _NormResult(_UnaryResult):
    def _bwd(self, sess, dU, grad):  # math_funcs.go:238-247
        if self.constant(grad):
            return
        ctx = context()
        x = _in_d(self.input)
        d = ctx.empty(x.n)
        check(ctx.lib.af_norm_bwd_r(ctx.h, x.ptr, None, self.d.ptr, None, dU.ptr, None, d.ptr, None, x.n))
        _down(self.input, sess, d, grad)

    def _bwd_r(self, sess, dU, dUR, rgrad, grad):  # math_funcs.go:267-284
        if self.constant(rgrad, grad):
            return
        ctx = context()
        x = _in_d(self.input)
        d, dr = ctx.empty(x.n), ctx.empty(x.n)
        check(ctx.lib.af_norm_bwd_r(ctx.h, x.ptr, _p(_in_dr(self.input)), self.d.ptr, self.dr.ptr, dU.ptr, dUR.ptr,
                                    d.ptr, dr.ptr, x.n))
        _down_r(self.input, sess, d, dr, rgrad, grad)


# --------------------------------------------------------------------------------------------
# Softmax (math_funcs.go:479-509): one fused kernel per direction, batched over the rows of a packed batch
# --------------------------------------------------------------------------------------------
class Softmax:
    """`apply` / `apply_r` are the reference's whole-vector Softmax; `batch` / `batch_r` are
    RFuncBatcher{&Softmax{T}} (batcher.go:57-84): one softmax per sample, all samples in one launch."""

    def __init__(self, temperature: float = 0.0):
        self.temperature = float(temperature)

    def apply(self, inp):
        return self.batch(inp, 1)

    def apply_r(self, rv, inp):
        return self.batch_r(rv, inp, 1)

    def _len(self, inp, n):
        total = _in_len(inp)
        if n < 1 or total % n != 0:
            raise ShapeError(_lib.AF_ESHAPE, "input count does not divide input length")  # batcher.go:41
        return total // n

    def batch(self, inp, n):
        ln = self._len(inp, n)
        ctx = context()
        x = _in_d(inp)
        y = ctx.empty(x.n)
        check(ctx.lib.af_softmax_fwd_r(ctx.h, x.ptr, None, self.temperature, y.ptr, None, ln, n))
        return _SoftmaxResult(self, inp, ln, n, y)

    def batch_r(self, rv, inp, n):
        ln = self._len(inp, n)
        ctx = context()
        x = _in_d(inp)
        y, ry = ctx.empty(x.n), ctx.empty(x.n)
        check(ctx.lib.af_softmax_fwd_r(ctx.h, x.ptr, _p(_in_dr(inp)), self.temperature, y.ptr, ry.ptr, ln, n))
        return _SoftmaxResult(self, inp, ln, n, y, ry, True)


class _SoftmaxResult(_UnaryResult):
    def __init__(self, fn, inp, ln, n, d, dr=None, is_r=False):
        super().__init__(inp, d, dr, is_r)
        self.fn, self.len, self.n = fn, ln, n

    def _bwd(self, sess, dU, grad):
        if self.input.constant(grad):
            return
        ctx = context()
        check(ctx.lib.af_softmax_bwd_r(ctx.h, self.d.ptr, None, self.fn.temperature, dU.ptr, None, self.len, self.n))
        _down(self.input, sess, dU, grad)

    def _bwd_r(self, sess, dU, dUR, rgrad, grad):
        if self.input.constant(rgrad, grad):
            return
        ctx = context()
        check(ctx.lib.af_softmax_bwd_r(ctx.h, self.d.ptr, self.dr.ptr, self.fn.temperature, dU.ptr, dUR.ptr, self.len,
                                       self.n))
        _down_r(self.input, sess, dU, dUR, rgrad, grad)


# --------------------------------------------------------------------------------------------
# squared-error cost head
# --------------------------------------------------------------------------------------------
class _SqErr(_BinaryResult):
    def _bwd(self, sess, dU, grad):
        ctx = context()
        a, y = _in_d(self.r1), _in_d(self.r2)
        for r, sign in ((self.r1, 1.0), (self.r2, -1.0)):
            if not r.constant(grad):
                g = ctx.empty(a.n)
                check(ctx.lib.af_sqerr_bwd_r(ctx.h, a.ptr, None, y.ptr, None, dU.ptr, None, sign, g.ptr, None, a.n))
                _down(r, sess, g, grad)

    def _bwd_r(self, sess, dU, dUR, rgrad, grad):
        ctx = context()
        a, y = _in_d(self.r1), _in_d(self.r2)
        ar, yr = _in_dr(self.r1), _in_dr(self.r2)
        for r, sign in ((self.r1, 1.0), (self.r2, -1.0)):
            if not r.constant(rgrad, grad):
                g, gr = ctx.empty(a.n), ctx.empty(a.n)
                check(ctx.lib.af_sqerr_bwd_r(ctx.h, a.ptr, _p(ar), y.ptr, _p(yr), dU.ptr, dUR.ptr, sign, g.ptr, gr.ptr,
                                             a.n))
                _down_r(r, sess, g, gr, rgrad, grad)


def squared_error(actual, expected):
    """0.5 * ||actual - expected||^2 -- Scale(SquaredNorm(Sub(a, y)), .5) of the reference in one reduction."""
    _same_len(actual, expected)
    ctx = context()
    a, y = _in_d(actual), _in_d(expected)
    o = ctx.empty(1)
    check(ctx.lib.af_sqerr_fwd_r(ctx.h, a.ptr, None, y.ptr, None, o.ptr, None, a.n))
    return _SqErr(actual, expected, o)


def squared_error_r(actual, expected):
    _same_len(actual, expected)
    ctx = context()
    a, y = _in_d(actual), _in_d(expected)
    o, ro = ctx.empty(1), ctx.empty(1)
    check(ctx.lib.af_sqerr_fwd_r(ctx.h, a.ptr, _p(_in_dr(actual)), y.ptr, _p(_in_dr(expected)), o.ptr, ro.ptr, a.n))
    return _SqErr(actual, expected, o, ro, True)


# --------------------------------------------------------------------------------------------
# Split (slices.go:310-335)
# --------------------------------------------------------------------------------------------
def split(n, inp):
    total = _in_len(inp)
    if total % n != 0:
        raise ShapeError(_lib.AF_ESHAPE, "count does not divide input length")
    size = total // n
    return [slice_(inp, i * size, (i + 1) * size) for i in range(n)]


def split_r(n, inp):
    total = _in_len(inp)
    if total % n != 0:
        raise ShapeError(_lib.AF_ESHAPE, "count does not divide input length")
    size = total // n
    return [slice_r(inp, i * size, (i + 1) * size) for i in range(n)]


# --------------------------------------------------------------------------------------------
# temp-variable protocol helpers
# --------------------------------------------------------------------------------------------
def _pool_var(r) -> Variable:
    return Variable.from_device(_in_d(r))


def _pool_rvar(r) -> RVariable:
    pv = _pool_var(r)
    dr = _in_dr(r)
    return RVariable(pv, {pv: dr} if dr is not None else {})


def _marker(pv):
    return {pv: _TempGrad(0)}  # Gradient{v: linalg.Vector{}} at pool.go:107,117,159,174


# --------------------------------------------------------------------------------------------
# Pool / PoolAll (pool.go)
# --------------------------------------------------------------------------------------------
class _Pooled(_DeviceResult):
    def __init__(self, ins, pool_vars, f_output, is_r):
        super().__init__(_in_d(f_output), _in_dr(f_output) if is_r else None, is_r)
        self.inputs, self.pool_vars, self.f_output = ins, pool_vars, f_output

    def constant(self, *maps):  # pool.go:102-112,150-160
        if not self.f_output.constant(*maps):
            return False
        for pv, r in zip(self.pool_vars, self.inputs):
            unused = self.f_output.constant(_marker(pv), None) if self.is_r else self.f_output.constant(_marker(pv))
            if not unused and not r.constant(*maps):
                return False
        return True

    def _bwd(self, sess, dU, grad):  # pool.go:114-136
        consts = []
        for pv, r in zip(self.pool_vars, self.inputs):
            c = self.f_output.constant(_marker(pv)) or r.constant(grad)
            consts.append(c)
            if not c:
                grad[pv] = _TempGrad(pv.size)
        _down(self.f_output, sess, dU, grad)
        ups = [None if c else sess.take(grad, pv) for pv, c in zip(self.pool_vars, consts)]
        for r, c, up in zip(self.inputs, consts, ups):
            if not c:
                _down(r, sess, up, grad)

    def _bwd_r(self, sess, dU, dUR, rgrad, grad):  # pool.go:162-195
        if grad is None:
            grad = {}
        consts = []
        for pv, r in zip(self.pool_vars, self.inputs):
            c = self.f_output.constant(_marker(pv), None) or r.constant(rgrad, grad)
            consts.append(c)
            if not c:
                grad[pv], rgrad[pv] = _TempGrad(pv.size), _TempGrad(pv.size)
        _down_r(self.f_output, sess, dU, dUR, rgrad, grad)
        ups = [None if c else (sess.take(grad, pv), sess.take(rgrad, pv)) for pv, c in zip(self.pool_vars, consts)]
        for r, c, up in zip(self.inputs, consts, ups):
            if not c:
                _down_r(r, sess, up[0], up[1], rgrad, grad)


def pool_all(ins, f):  # pool.go:38-50
    ins = list(ins)
    pvs = [_pool_var(r) for r in ins]
    return _Pooled(ins, pvs, f(list(pvs)), False)


def pool_all_r(ins, f):  # pool.go:53-68
    ins = list(ins)
    prs = [_pool_rvar(r) for r in ins]
    return _Pooled(ins, [p.variable for p in prs], f(prs), True)


def pool(r, f):  # pool.go:21-25
    return pool_all([r], lambda xs: f(xs[0]))


def pool_r(r, f):  # pool.go:28-32
    return pool_all_r([r], lambda xs: f(xs[0]))


# --------------------------------------------------------------------------------------------
# Fold / FoldR (fold.go): the recurrent state never leaves HBM between timesteps
# --------------------------------------------------------------------------------------------
class _Fold(_DeviceResult):
    def __init__(self, pool_vars, intermediate, final, is_r):
        super().__init__(_in_d(final), _in_dr(final) if is_r else None, is_r)
        self.pool, self.intermediate, self.final = pool_vars, intermediate, final

    def _with_temp(self, p, maps, fn):
        """Run fn() with `p` present in the gradient maps (fold.go:45-48,125-134)."""
        live = [m for m in maps if m is not None]
        for m in live:
            m[p] = _TempGrad(p.size)
        try:
            return fn()
        finally:
            for m in live:
                m.pop(p, None)

    def constant(self, *maps):  # fold.go:35-54,113-139
        if not self.final.constant(*maps):
            return False
        for i in range(len(self.pool) - 1, -1, -1):
            if not self.intermediate[i].constant(*maps):
                return False
            if i > 0:
                if self._with_temp(self.pool[i - 1], maps, lambda: self.intermediate[i].constant(*maps)):
                    return True
        return True

    def _bwd(self, sess, dU, grad):  # fold.go:56-84
        if not self.pool:
            _down(self.final, sess, dU, grad)
            return
        state_up = dU
        for i in range(len(self.pool), 0, -1):
            res = self.final if i == len(self.pool) else self.intermediate[i]
            p = self.pool[i - 1]
            grad[p] = _TempGrad(p.size)
            if res.constant(grad):
                del grad[p]
                break
            _down(res, sess, state_up, grad)
            state_up = sess.take(grad, p)
        _down(self.intermediate[0], sess, state_up, grad)

    def _bwd_r(self, sess, dU, dUR, rgrad, grad):  # fold.go:141-178
        if not self.pool:
            _down_r(self.final, sess, dU, dUR, rgrad, grad)
            return
        if grad is None:
            grad = {}
        up, up_r = dU, dUR
        for i in range(len(self.pool), 0, -1):
            res = self.final if i == len(self.pool) else self.intermediate[i]
            p = self.pool[i - 1]
            grad[p], rgrad[p] = _TempGrad(p.size), _TempGrad(p.size)
            if res.constant(rgrad, grad):
                del grad[p]
                del rgrad[p]
                break
            _down_r(res, sess, up, up_r, rgrad, grad)
            up, up_r = sess.take(grad, p), sess.take(rgrad, p)
        _down_r(self.intermediate[0], sess, up, up_r, rgrad, grad)


def fold(state, ins, f):  # fold.go:19-30
    pool_vars, inter = [], []
    for r in ins:
        inter.append(state)
        pv = _pool_var(state)
        pool_vars.append(pv)
        state = f(pv, r)
    return _Fold(pool_vars, inter, state, False)


def fold_r(state, ins, f):  # fold.go:93-107
    pool_vars, inter = [], []
    for r in ins:
        inter.append(state)
        pr = _pool_rvar(state)
        pool_vars.append(pr.variable)
        state = f(pr, r)
    return _Fold(pool_vars, inter, state, True)


# --------------------------------------------------------------------------------------------
# MatMulVecs (lin_alg.go:14-133): LinTran whose matrix is itself a Result; reuses G1-G6 verbatim
# --------------------------------------------------------------------------------------------
class _MatMul(_DeviceResult):
    def __init__(self, mat, mat_var, res, is_r):
        super().__init__(res.d, res.dr if is_r else None, is_r)
        self.mat_in, self.mat_var, self.res = mat, mat_var, res

    def constant(self, *maps):
        return self.res.constant(*maps) and self.mat_in.constant(*maps)

    def _bwd(self, sess, dU, grad):  # lin_alg.go:57-69
        mat_const = self.mat_in.constant(grad)
        if not mat_const:
            grad[self.mat_var] = _TempGrad(self.mat_var.size)
        self.res._bwd(sess, dU, grad)
        if not mat_const:
            _down(self.mat_in, sess, sess.take(grad, self.mat_var), grad)

    def _bwd_r(self, sess, dU, dUR, rgrad, grad):  # lin_alg.go:113-133
        mat_const = self.mat_in.constant(rgrad, grad)
        if not mat_const:
            if grad is None:
                grad = {}
            grad[self.mat_var] = _TempGrad(self.mat_var.size)
            rgrad[self.mat_var] = _TempGrad(self.mat_var.size)
        self.res._bwd_r(sess, dU, dUR, rgrad, grad)
        if not mat_const:
            d, dr = sess.take(grad, self.mat_var), sess.take(rgrad, self.mat_var)
            _down_r(self.mat_in, sess, d, dr, rgrad, grad)


def _matmul_check(mat, rows, cols, vecs) -> int:
    if _in_len(mat) != rows * cols:
        raise ShapeError(_lib.AF_ESHAPE, "invalid matrix data size")  # lin_alg.go:29-31
    if _in_len(vecs) % cols != 0:
        raise ShapeError(_lib.AF_ESHAPE, "invalid vecs size")  # lin_alg.go:32-34
    return _in_len(vecs) // cols


def mat_mul_vecs(mat, rows, cols, vecs):  # lin_alg.go:28-49
    n = _matmul_check(mat, rows, cols, vecs)
    mv = _pool_var(mat)
    return _MatMul(mat, mv, LinTran(mv, rows, cols).batch(vecs, n), False)


def mat_mul_vecs_r(mat, rows, cols, vecs):  # lin_alg.go:84-105
    n = _matmul_check(mat, rows, cols, vecs)
    mr = _pool_rvar(mat)
    mv = mr.variable
    rv = {mv: mr._rdev} if mr._has_r else {}
    return _MatMul(mat, mv, LinTran(mv, rows, cols).batch_r(rv, vecs, n), True)


mat_mul_vec, mat_mul_vec_r = mat_mul_vecs, mat_mul_vecs_r

__all__ = [
    "result_len", "Log", "LogSigmoid", "Sin", "Cos", "Tanh", "Norm", "Softmax", "add_scaler", "add_scaler_r", "inverse",
    "inverse_r", "pow_", "pow_r", "div", "div_r", "scale_first", "scale_first_r", "add_first", "add_first_r",
    "add_log_domain", "add_log_domain_r", "sum_all_log_domain", "sum_all_log_domain_r", "squared_error",
    "squared_error_r", "split", "split_r", "pool", "pool_r", "pool_all", "pool_all_r", "fold", "fold_r", "mat_mul_vecs",
    "mat_mul_vecs_r", "mat_mul_vec", "mat_mul_vec_r",
]
