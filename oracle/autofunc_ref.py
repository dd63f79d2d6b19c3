"""CPU oracle (TEST INFRASTRUCTURE ONLY) -- numpy restatement of unixpickle/autofunc's
dense-network differentiation path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
leg may import this module.  The product (autofunc_b200) never does.

Parity status: the reference is pure Go and no Go toolchain exists in the build image or
on the GPU boxes, so the reference itself cannot be executed.  This restatement is pinned
against every golden vector the reference's own tests hold for the path
(test/lin_tran_test.go:65-97 -> [19,20,36] and [349,131]; fold/aggregate KATs are off the
path) and against the reference's own acceptance procedure (functest.RFuncChecker.FullCheck,
restated in oracle/functest.py) on the reference's fixtures.  Below the 1e-5 level that
those tests resolve, parity is *defined* by this oracle ("parity pinned on KATs + FD
properties; 1e-10 level unpinned by the reference").

Every class cites the reference file:line it follows.  Vectors are 1-D float64 numpy
arrays; a packed batch is sample-major (batcher.go:6-17).  Gradient / RGradient / RVector
are dicts keyed by Variable identity (gradient.go:11,62,93).
"""
from __future__ import annotations

import numpy as np

F64 = np.float64


def vec(x) -> np.ndarray:
    return np.array(x, dtype=F64).reshape(-1)


# --------------------------------------------------------------------------------------
# core data model: variable.go, gradient.go, result.go
# --------------------------------------------------------------------------------------
class Variable:
    """variable.go:19-21 -- a vector whose *identity* is the gradient-map key."""

    __hash__ = object.__hash__

    def __init__(self, v):
        self.vector = vec(v)

    def output(self):  # variable.go:39
        return self.vector

    def constant(self, g):  # variable.go:49-52
        return self not in g

    def propagate_gradient(self, upstream, grad):  # variable.go:43-47
        if self in grad:
            grad[self] += upstream

    # wire format, variable.go:24-37,61-68: LE u64 count + LE f64s
    def serialize(self) -> bytes:
        return np.uint64(len(self.vector)).astype("<u8").tobytes() + self.vector.astype("<f8").tobytes()

    @staticmethod
    def deserialize(b: bytes) -> "Variable":
        n = int(np.frombuffer(b[:8], dtype="<u8")[0])
        return Variable(np.frombuffer(b[8:8 + 8 * n], dtype="<f8").copy())


class RVariable:
    """variable.go:73-123 -- Variable paired with rv[v], or fresh zeros when absent."""

    def __init__(self, v: Variable, rv: dict):
        self.variable = v
        self.routput_vec = rv[v] if v in rv else np.zeros(len(v.vector), dtype=F64)

    def output(self):
        return self.variable.vector

    def routput(self):
        return self.routput_vec

    def constant(self, rg, g):  # variable.go:102-111
        if self.variable in rg:
            return False
        if g is None:
            return True
        return self.variable not in g

    def propagate_rgradient(self, upstream, upstream_r, rgrad, grad):  # variable.go:113-123
        if grad is not None and self.variable in grad:
            grad[self.variable] += upstream
        if self.variable in rgrad:
            rgrad[self.variable] += upstream_r


def new_gradient(variables):  # gradient.go:13-19
    return {v: np.zeros(len(v.vector), dtype=F64) for v in variables}


new_rgradient = new_gradient


def gradient_add(g, g1):  # gradient.go:28-30,108-120
    for k in g:
        g[k] += g1[k]


def gradient_add_to_vars(g, scale):  # gradient.go:40-52
    for v, gv in g.items():
        v.vector += scale * gv


# --------------------------------------------------------------------------------------
# LinTran: lin_tran.go
# --------------------------------------------------------------------------------------
class LinTran:
    """lin_tran.go:17-21.  W is rows x cols row-major; y = W x per packed sample."""

    def __init__(self, data: Variable, rows: int, cols: int):
        self.data, self.rows, self.cols = data, rows, cols

    def _w(self, flat):
        return flat.reshape(self.rows, self.cols)

    # lin_tran.go:24-30, 50-60
    def apply(self, inp):
        if len(inp.output()) != self.cols:
            raise ValueError("input length should be %d but got %d" % (self.cols, len(inp.output())))
        return self.batch(inp, 1)

    def batch(self, inp, n):
        if self.cols * n != len(inp.output()):
            raise ValueError("input length should be %d but got %d" % (self.cols * n, len(inp.output())))
        return _LinTranResult(self, inp, self.multiply(inp.output()))

    # lin_tran.go:33-46, 64-77
    def apply_r(self, rv, inp):
        if len(inp.output()) != self.cols:
            raise ValueError("input length should be %d but got %d" % (self.cols, len(inp.output())))
        return self.batch_r(rv, inp, 1)

    def batch_r(self, rv, inp, n):
        if self.cols * n != len(inp.output()):
            raise ValueError("input length should be %d but got %d" % (self.cols * n, len(inp.output())))
        rdata = RVariable(self.data, rv)
        return _LinTranRResult(self, inp, rdata, self.multiply(inp.output()), self.multiply_r(rdata, inp))

    # G1  lin_tran.go:79-117   Y = X W^T
    def multiply(self, x):
        X = x.reshape(-1, self.cols)
        return (X @ self._w(self.data.vector).T).reshape(-1)

    # G2  lin_tran.go:119-177  RY = X RW^T (beta=0) then += RX W^T (beta=1)
    def multiply_r(self, rdata, rin):
        X = rin.output().reshape(-1, self.cols)
        RX = rin.routput().reshape(-1, self.cols)
        out = X @ self._w(rdata.routput()).T
        out += RX @ self._w(rdata.output()).T
        return out.reshape(-1)

    # G3  lin_tran.go:179-212  gW += U^T X
    def data_gradient(self, upstream, x, grad):
        U = upstream.reshape(-1, self.rows)
        X = x.reshape(-1, self.cols)
        g = grad[self.data]
        g += (U.T @ X).reshape(-1)

    # G4  lin_tran.go:214-270  rgW += U^T RX ; rgW += UR^T X
    def data_gradient_r(self, upstream, upstream_r, x, xr, rg):
        U = upstream.reshape(-1, self.rows)
        UR = upstream_r.reshape(-1, self.rows)
        X = x.reshape(-1, self.cols)
        RX = xr.reshape(-1, self.cols)
        rg += (U.T @ RX).reshape(-1)
        rg += (UR.T @ X).reshape(-1)

    # G5  lin_tran.go:272-306  D = U W
    def input_gradient(self, upstream):
        U = upstream.reshape(-1, self.rows)
        return (U @ self._w(self.data.vector)).reshape(-1)

    # G6  lin_tran.go:308-359  DR = U RW (beta=0) ; += UR W (beta=1)
    def input_gradient_r(self, rdata, upstream, upstream_r):
        U = upstream.reshape(-1, self.rows)
        UR = upstream_r.reshape(-1, self.rows)
        out = UR @ self._w(rdata.output())
        out += U @ self._w(rdata.routput())
        return out.reshape(-1)


class _LinTranResult:  # lin_tran.go:361-386
    def __init__(self, m, inp, out):
        self.matrix, self.input, self.output_vec = m, inp, out

    def output(self):
        return self.output_vec

    def constant(self, g):
        return self.matrix.data.constant(g) and self.input.constant(g)

    def propagate_gradient(self, upstream, grad):  # lin_tran.go:373-382
        if not self.matrix.data.constant(grad):
            self.matrix.data_gradient(upstream, self.input.output(), grad)
        if not self.input.constant(grad):
            self.input.propagate_gradient(self.matrix.input_gradient(upstream), grad)


class _LinTranRResult:  # lin_tran.go:388-425
    def __init__(self, m, inp, rdata, out, rout):
        self.matrix, self.input, self.rdata = m, inp, rdata
        self.output_vec, self.routput_vec = out, rout

    def output(self):
        return self.output_vec

    def routput(self):
        return self.routput_vec

    def constant(self, rg, g):
        return self.input.constant(rg, g) and self.rdata.constant(rg, g)

    def propagate_rgradient(self, upstream, upstream_r, rgrad, grad):  # lin_tran.go:408-425
        m = self.matrix
        if grad is not None and not m.data.constant(grad):
            m.data_gradient(upstream, self.input.output(), grad)
        if m.data in rgrad:
            m.data_gradient_r(upstream, upstream_r, self.input.output(), self.input.routput(), rgrad[m.data])
        if not self.input.constant(rgrad, grad):
            dr = m.input_gradient_r(self.rdata, upstream, upstream_r)
            d = m.input_gradient(upstream)
            self.input.propagate_rgradient(d, dr, rgrad, grad)


# --------------------------------------------------------------------------------------
# LinAdd: lin_add.go   (unbatched in the reference; `BatchedLinAdd` below restates the
# batched form the reference expresses as Add(x, Repeat(b, n)): arithmetic.go:17,
# slices.go:216-308)
# --------------------------------------------------------------------------------------
class LinAdd:
    def __init__(self, var: Variable):
        self.var = var

    def apply(self, inp):  # lin_add.go:13-23
        return _LinAddResult(inp.output() + self.var.vector, self.var, inp)

    def apply_r(self, rv, inp):  # lin_add.go:26-50
        rvar = RVariable(self.var, rv)
        return _LinAddRResult(rvar.output() + inp.output(), rvar.routput() + inp.routput(), rvar, inp)

    # batched semantics == Add(in, Repeat(b, n))
    def batch(self, inp, n):
        if len(inp.output()) != n * len(self.var.vector):
            raise ValueError("input length should be %d but got %d" % (n * len(self.var.vector), len(inp.output())))
        return add(inp, repeat(self.var, n))

    def batch_r(self, rv, inp, n):
        if len(inp.output()) != n * len(self.var.vector):
            raise ValueError("input length should be %d but got %d" % (n * len(self.var.vector), len(inp.output())))
        return add_r(inp, repeat_r(RVariable(self.var, rv), n))


class _LinAddResult:  # lin_add.go:52-73
    def __init__(self, out, var, inp):
        self.output_vec, self.sum_var, self.input = out, var, inp

    def output(self):
        return self.output_vec

    def constant(self, g):
        return self.input.constant(g) and self.sum_var.constant(g)

    def propagate_gradient(self, upstream, grad):
        if self.sum_var in grad:
            grad[self.sum_var] += upstream
        if not self.input.constant(grad):
            self.input.propagate_gradient(upstream, grad)


class _LinAddRResult:  # lin_add.go:75-109
    def __init__(self, out, rout, rvar, inp):
        self.output_vec, self.routput_vec, self.sum_var, self.input = out, rout, rvar, inp

    def output(self):
        return self.output_vec

    def routput(self):
        return self.routput_vec

    def constant(self, rg, g):
        return self.sum_var.constant(rg, g) and self.input.constant(rg, g)

    def propagate_rgradient(self, upstream, upstream_r, rgrad, grad):
        v = self.sum_var.variable
        if grad is not None and v in grad:
            grad[v] += upstream
        if v in rgrad:
            rgrad[v] += upstream_r
        if not self.input.constant(rgrad, grad):
            self.input.propagate_rgradient(upstream, upstream_r, rgrad, grad)


# --------------------------------------------------------------------------------------
# elementwise RFuncs: math_funcs.go
# --------------------------------------------------------------------------------------
class _UnaryResult:
    def __init__(self, out, inp):
        self.output_vec, self.input = out, inp

    def output(self):
        return self.output_vec

    def constant(self, g):
        return self.input.constant(g)


class _UnaryRResult:
    def __init__(self, out, rout, inp):
        self.output_vec, self.routput_vec, self.input = out, rout, inp

    def output(self):
        return self.output_vec

    def routput(self):
        return self.routput_vec

    def constant(self, rg, g):
        return self.input.constant(rg, g)


class _ElementwiseFunc:
    """Elementwise RFunc; also usable as a Batcher since the op is shape-agnostic
    (the reference reaches the same values through RFuncBatcher, batcher.go:57-84)."""

    def batch(self, inp, n):
        if len(inp.output()) % n != 0:
            raise ValueError("input count does not divide input length")
        return self.apply(inp)

    def batch_r(self, rv, inp, n):
        if len(inp.output()) % n != 0:
            raise ValueError("input count does not divide input length")
        return self.apply_r(rv, inp)


class Sigmoid(_ElementwiseFunc):  # math_funcs.go:286-374
    def apply(self, inp):
        return _SigmoidResult(1 / (1 + np.exp(-inp.output())), inp)

    def apply_r(self, rv, inp):
        s = 1 / (1 + np.exp(-inp.output()))
        return _SigmoidRResult(s, s * (1 - s) * inp.routput(), inp)


class _SigmoidResult(_UnaryResult):
    def propagate_gradient(self, upstream, grad):  # math_funcs.go:326-333 (in place)
        if not self.input.constant(grad):
            x = self.output_vec
            upstream *= x * (1 - x)
            self.input.propagate_gradient(upstream, grad)


class _SigmoidRResult(_UnaryRResult):
    def propagate_rgradient(self, upstream, upstream_r, rgrad, grad):  # math_funcs.go:357-374
        if self.input.constant(rgrad, grad):
            return
        x, xr = self.output_vec, self.routput_vec
        p, pr = upstream.copy(), upstream_r.copy()
        upstream[:] = x * (1 - x) * p
        upstream_r[:] = xr * (1 - x) * p - x * xr * p + x * (1 - x) * pr
        self.input.propagate_rgradient(upstream, upstream_r, rgrad, grad)


class Exp(_ElementwiseFunc):  # math_funcs.go:12-97
    def apply(self, inp):
        return _ExpResult(np.exp(inp.output()), inp)

    def apply_r(self, rv, inp):
        e = np.exp(inp.output())
        return _ExpRResult(e, e * inp.routput(), inp)


class _ExpResult(_UnaryResult):
    def propagate_gradient(self, upstream, grad):  # math_funcs.go:56-63
        if not self.input.constant(grad):
            upstream *= self.output_vec
            self.input.propagate_gradient(upstream, grad)


class _ExpRResult(_UnaryRResult):
    def propagate_rgradient(self, upstream, upstream_r, rgrad, grad):  # math_funcs.go:83-97
        if not self.input.constant(rgrad, grad):
            x, xr = self.output_vec, self.routput_vec
            u = upstream.copy()
            upstream[:] = u * x
            upstream_r[:] = upstream_r * x + u * xr
            self.input.propagate_rgradient(upstream, upstream_r, rgrad, grad)


class SquaredNorm(_ElementwiseFunc):  # math_funcs.go:191-199
    def apply(self, inp):
        return sum_all(square(inp))

    def apply_r(self, rv, inp):
        return sum_all_r(square_r(inp))


# --------------------------------------------------------------------------------------
# arithmetic.go (hot subset: Add, Sub, Mul, Scale, Square, SumAll)
# --------------------------------------------------------------------------------------
class _Binary:
    def __init__(self, out, r1, r2):
        self.output_vec, self.r1, self.r2 = out, r1, r2

    def output(self):
        return self.output_vec

    def constant(self, g):
        return self.r1.constant(g) and self.r2.constant(g)


class _BinaryR:
    def __init__(self, out, rout, r1, r2):
        self.output_vec, self.routput_vec, self.r1, self.r2 = out, rout, r1, r2

    def output(self):
        return self.output_vec

    def routput(self):
        return self.routput_vec

    def constant(self, rg, g):
        return self.r1.constant(rg, g) and self.r2.constant(rg, g)


class _Sum(_Binary):  # arithmetic.go:34-49
    def propagate_gradient(self, upstream, grad):
        r2c = self.r2.constant(grad)
        if not self.r1.constant(grad):
            self.r1.propagate_gradient(upstream if r2c else upstream.copy(), grad)
        if not r2c:
            self.r2.propagate_gradient(upstream, grad)


class _SumR(_BinaryR):  # arithmetic.go:79-96
    def propagate_rgradient(self, upstream, upstream_r, rgrad, grad):
        r2c = self.r2.constant(rgrad, grad)
        if not self.r1.constant(rgrad, grad):
            if r2c:
                self.r1.propagate_rgradient(upstream, upstream_r, rgrad, grad)
            else:
                self.r1.propagate_rgradient(upstream.copy(), upstream_r.copy(), rgrad, grad)
        if not r2c:
            self.r2.propagate_rgradient(upstream, upstream_r, rgrad, grad)


def add(r1, r2):  # arithmetic.go:17-23
    return _Sum(r1.output() + r2.output(), r1, r2)


def add_r(r1, r2):  # arithmetic.go:58-65
    return _SumR(r1.output() + r2.output(), r1.routput() + r2.routput(), r1, r2)


class _Diff(_Binary):  # arithmetic.go:128-135
    def propagate_gradient(self, u, g):
        if not self.r1.constant(g):
            self.r1.propagate_gradient(u.copy(), g)
        if not self.r2.constant(g):
            u *= -1
            self.r2.propagate_gradient(u, g)


class _DiffR(_BinaryR):  # arithmetic.go:175-182
    def propagate_rgradient(self, u, ur, rg, g):
        if not self.r1.constant(rg, g):
            self.r1.propagate_rgradient(u.copy(), ur.copy(), rg, g)
        if not self.r2.constant(rg, g):
            u *= -1
            ur *= -1
            self.r2.propagate_rgradient(u, ur, rg, g)


def sub(a, b):  # arithmetic.go:105-117
    return _Diff(a.output() - b.output(), a, b)


def sub_r(a, b):  # arithmetic.go:144-161
    return _DiffR(a.output() - b.output(), a.routput() - b.routput(), a, b)


class _Product(_Binary):  # arithmetic.go:288-306
    def propagate_gradient(self, upstream, grad):
        if not self.r1.constant(grad):
            self.r1.propagate_gradient(upstream * self.r2.output(), grad)
        if not self.r2.constant(grad):
            self.r2.propagate_gradient(upstream * self.r1.output(), grad)


class _ProductR(_BinaryR):  # arithmetic.go:349-376
    def propagate_rgradient(self, upstream, upstream_r, rgrad, grad):
        if not self.r1.constant(rgrad, grad):
            o, orr = self.r2.output(), self.r2.routput()
            self.r1.propagate_rgradient(upstream * o, upstream * orr + upstream_r * o, rgrad, grad)
        if not self.r2.constant(rgrad, grad):
            o, orr = self.r1.output(), self.r1.routput()
            self.r2.propagate_rgradient(upstream * o, upstream * orr + upstream_r * o, rgrad, grad)


def mul(r1, r2):  # arithmetic.go:261-276
    if len(r1.output()) != len(r2.output()):
        raise ValueError("vector sizes do not match")
    return _Product(r1.output() * r2.output(), r1, r2)


def mul_r(r1, r2):  # arithmetic.go:315-335
    if len(r1.output()) != len(r2.output()):
        raise ValueError("vector sizes do not match")
    x, y = r1.output(), r2.output()
    return _ProductR(x * y, x * r2.routput() + r1.routput() * y, r1, r2)


class _Scaled(_UnaryResult):  # arithmetic.go:444-456
    def propagate_gradient(self, upstream, grad):
        if not self.input.constant(grad):
            upstream *= self.scaler
            self.input.propagate_gradient(upstream, grad)


class _ScaledR(_UnaryRResult):  # arithmetic.go:476-493
    def propagate_rgradient(self, upstream, upstream_r, rgrad, grad):
        if not self.input.constant(rgrad, grad):
            upstream *= self.scaler
            upstream_r *= self.scaler
            self.input.propagate_rgradient(upstream, upstream_r, rgrad, grad)


def scale(r, f):  # arithmetic.go:436-442
    res = _Scaled(r.output() * f, r)
    res.scaler = f
    return res


def scale_r(r, f):  # arithmetic.go:466-474
    res = _ScaledR(r.output() * f, r.routput() * f, r)
    res.scaler = f
    return res


class _Square(_UnaryResult):  # arithmetic.go:716-723
    def propagate_gradient(self, upstream, grad):
        if not self.input.constant(grad):
            upstream *= self.input.output() * 2
            self.input.propagate_gradient(upstream, grad)


class _SquareR(_UnaryRResult):  # arithmetic.go:760-770
    def propagate_rgradient(self, upstream, upstream_r, rgrad, grad):
        if not self.input.constant(rgrad, grad):
            x, xr = self.input.output(), self.input.routput()
            upstream_r[:] = 2 * (upstream_r * x + upstream * xr)
            upstream *= x * 2
            self.input.propagate_rgradient(upstream, upstream_r, rgrad, grad)


def square(r):  # arithmetic.go:696-706
    x = r.output()
    return _Square(x * x, r)


def square_r(r):  # arithmetic.go:732-746
    x = r.output()
    return _SquareR(x * x, 2 * x * r.routput(), r)


def _seq_sum(x):
    # arithmetic.go:973-976 sums left to right; np.add.reduce on a 1-D f64 array uses
    # pairwise summation, so restate the loop order exactly with cumsum's sequential add.
    return F64(np.cumsum(x)[-1]) if len(x) else F64(0)


class _SumAll(_UnaryResult):  # arithmetic.go:987-995
    def propagate_gradient(self, upstream, grad):
        if not self.input.constant(grad):
            n = len(self.input.output())
            self.input.propagate_gradient(np.full(n, upstream[0], dtype=F64), grad)


class _SumAllR(_UnaryRResult):  # arithmetic.go:1035-1047
    def propagate_rgradient(self, upstream, upstream_r, rgrad, grad):
        if not self.input.constant(rgrad, grad):
            n = len(self.input.output())
            self.input.propagate_rgradient(np.full(n, upstream[0], dtype=F64),
                                           np.full(n, upstream_r[0], dtype=F64), rgrad, grad)


def sum_all(r):  # arithmetic.go:972-981
    return _SumAll(vec([_seq_sum(r.output())]), r)


def sum_all_r(r):  # arithmetic.go:1009-1021
    return _SumAllR(vec([_seq_sum(r.output())]), vec([_seq_sum(r.routput())]), r)


# --------------------------------------------------------------------------------------
# slices.go: Concat, Repeat, Slice
# --------------------------------------------------------------------------------------
class _Joined:  # slices.go:8-57
    def __init__(self, results):
        self.results = results
        self.output_vec = np.concatenate([r.output() for r in results])

    def output(self):
        return self.output_vec

    def constant(self, g):
        return all(r.constant(g) for r in self.results)

    def propagate_gradient(self, upstream, grad):
        i = 0
        for r in self.results:
            l = len(r.output())
            if not r.constant(grad):
                r.propagate_gradient(upstream[i:i + l], grad)
            i += l


class _JoinedR:  # slices.go:59-119
    def __init__(self, results):
        self.results = results
        self.output_vec = np.concatenate([r.output() for r in results])
        self.routput_vec = np.concatenate([r.routput() for r in results])

    def output(self):
        return self.output_vec

    def routput(self):
        return self.routput_vec

    def constant(self, rg, g):
        return all(r.constant(rg, g) for r in self.results)

    def propagate_rgradient(self, upstream, upstream_r, rgrad, grad):
        i = 0
        for r in self.results:
            l = len(r.output())
            if not r.constant(rgrad, grad):
                r.propagate_rgradient(upstream[i:i + l], upstream_r[i:i + l], rgrad, grad)
            i += l


def concat(*rs):
    return _Joined(list(rs))


def concat_r(*rs):
    return _JoinedR(list(rs))


class _Repeat:  # slices.go:210-251
    def __init__(self, inp, n):
        self.repeated, self.n = inp, n
        self.output_vec = np.tile(inp.output(), n)

    def output(self):
        return self.output_vec

    def constant(self, g):
        return self.repeated.constant(g)

    def propagate_gradient(self, upstream, g):
        if self.repeated.constant(g):
            return
        if len(upstream) != len(self.output_vec):
            raise ValueError("invalid upstream size")
        pl = len(self.repeated.output())
        first = upstream[:pl]
        for i in range(1, self.n):
            first += upstream[i * pl:(i + 1) * pl]
        self.repeated.propagate_gradient(first, g)


class _RepeatR:  # slices.go:253-308
    def __init__(self, inp, n):
        self.repeated, self.n = inp, n
        self.output_vec = np.tile(inp.output(), n)
        self.routput_vec = np.tile(inp.routput(), n)

    def output(self):
        return self.output_vec

    def routput(self):
        return self.routput_vec

    def constant(self, rg, g):
        return self.repeated.constant(rg, g)

    def propagate_rgradient(self, upstream, upstream_r, rg, g):
        if self.repeated.constant(rg, g):
            return
        if len(upstream) != len(self.output_vec):
            raise ValueError("invalid upstream size")
        pl = len(self.repeated.output())
        first, first_r = upstream[:pl], upstream_r[:pl]
        for i in range(1, self.n):
            first += upstream[i * pl:(i + 1) * pl]
            first_r += upstream_r[i * pl:(i + 1) * pl]
        self.repeated.propagate_rgradient(first, first_r, rg, g)


def repeat(inp, n):
    return _Repeat(inp, n)


def repeat_r(inp, n):
    return _RepeatR(inp, n)


class _Sliced:  # slices.go:121-166 (view; backward scatters into zeros)
    def __init__(self, inp, start, end):
        self.input, self.start, self.end = inp, start, end

    def output(self):
        return self.input.output()[self.start:self.end]

    def constant(self, g):
        return self.input.constant(g)

    def propagate_gradient(self, upstream, grad):
        if not self.input.constant(grad):
            down = np.zeros(len(self.input.output()), dtype=F64)
            down[self.start:self.end] = upstream
            self.input.propagate_gradient(down, grad)


class _SlicedR:  # slices.go:168-206
    def __init__(self, inp, start, end):
        self.input, self.start, self.end = inp, start, end

    def output(self):
        return self.input.output()[self.start:self.end]

    def routput(self):
        return self.input.routput()[self.start:self.end]

    def constant(self, rg, g):
        return self.input.constant(rg, g)

    def propagate_rgradient(self, upstream, upstream_r, rg, g):
        if not self.input.constant(rg, g):
            n = len(self.input.output())
            d, dr = np.zeros(n, dtype=F64), np.zeros(n, dtype=F64)
            d[self.start:self.end] = upstream
            dr[self.start:self.end] = upstream_r
            self.input.propagate_rgradient(d, dr, rg, g)


def slice_(inp, start, end):
    return _Sliced(inp, start, end)


def slice_r(inp, start, end):
    return _SlicedR(inp, start, end)


# --------------------------------------------------------------------------------------
# composition: func.go, batcher.go
# --------------------------------------------------------------------------------------
class ComposedRFunc(list):  # func.go:29-45
    def apply(self, inp):
        for f in self:
            inp = f.apply(inp)
        return inp

    def apply_r(self, rv, inp):
        for f in self:
            inp = f.apply_r(rv, inp)
        return inp


ComposedFunc = ComposedRFunc


class ComposedRBatcher(list):  # batcher.go:86-111
    def batch(self, inp, n):
        for b in self:
            inp = b.batch(inp, n)
        return inp

    def batch_r(self, rv, inp, n):
        for b in self:
            inp = b.batch_r(rv, inp, n)
        return inp


ComposedBatcher = ComposedRBatcher


class RFuncBatcher:
    """batcher.go:34-84 -- the naive split / apply / concat adaptor.  (Pool is identity on
    values; it only de-duplicates backward calls, pool.go:5-25, so it is restated here as
    a pass-through that collects the n slices' gradients through _Sliced.)"""

    def __init__(self, f):
        self.f = f

    def batch(self, inp, n):
        l = len(inp.output())
        if l % n != 0:
            raise ValueError("input count does not divide input length")
        s = l // n
        return concat(*[self.f.apply(slice_(inp, i * s, (i + 1) * s)) for i in range(n)])

    def batch_r(self, rv, inp, n):
        l = len(inp.output())
        if l % n != 0:
            raise ValueError("input count does not divide input length")
        s = l // n
        return concat_r(*[self.f.apply_r(rv, slice_r(inp, i * s, (i + 1) * s)) for i in range(n)])


FuncBatcher = RFuncBatcher


# --------------------------------------------------------------------------------------
# workloads: bench/mlp.go, bench/lintran.go  (synthetic inputs from a counter-based
# generator -- Go's math/rand stream is not reproducible without Go)
# --------------------------------------------------------------------------------------
def splitmix_uniform(seed: int, count: int, lo: float = -1.0, hi: float = 1.0) -> np.ndarray:
    """Deterministic U[lo,hi) doubles: splitmix64 of (seed, index) -> 53-bit mantissa.
    The same generator is implemented in autofunc_b200/synth.py for the product side."""
    idx = np.arange(count, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = idx * np.uint64(0x9E3779B97F4A7C15) + np.uint64(seed) * np.uint64(0xD1B54A32D192ED03) + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    u = (z >> np.uint64(11)).astype(F64) * (1.0 / 9007199254740992.0)
    return lo + (hi - lo) * u


class MLP:
    """bench/mlp.go:24-126 with n >= 1 samples: LinTran -> LinAdd -> Sigmoid per layer."""

    def __init__(self, layer_sizes, seed=123, rseed=125, weight_scale=None):
        self.layer_sizes = list(layer_sizes)
        self.variables = []
        self.layers = ComposedRBatcher()
        k = 0
        for i, out_size in enumerate(self.layer_sizes[1:]):
            in_size = self.layer_sizes[i]
            sc = 1.0 if weight_scale is None else weight_scale(in_size)
            w = Variable(splitmix_uniform(seed + 1000 * k, in_size * out_size) * sc)
            b = Variable(splitmix_uniform(seed + 1000 * k + 1, out_size))
            k += 1
            self.layers += [LinTran(w, out_size, in_size), LinAdd(b), Sigmoid()]
            self.variables += [w, b]
        # bench/mlp.go:104-116: one R-vector per variable
        self.rvec = {v: splitmix_uniform(rseed + 1000 * j, len(v.vector)) for j, v in enumerate(self.variables)}

    def make_input(self, n, seed=124):
        return Variable(splitmix_uniform(seed, n * self.layer_sizes[0]))

    def step_r(self, x: Variable, n: int, out_grad=None, out_rgrad=None):
        """bench/mlp.go:43-57 (one iteration, fresh accumulators and upstreams)."""
        res = self.layers.batch_r(self.rvec, RVariable(x, self.rvec), n)
        m = self.layer_sizes[-1] * n
        u = np.ones(m, dtype=F64) if out_grad is None else vec(out_grad).copy()
        ur = np.zeros(m, dtype=F64) if out_rgrad is None else vec(out_rgrad).copy()
        grad, rgrad = new_gradient(self.variables), new_rgradient(self.variables)
        res.propagate_rgradient(u, ur, rgrad, grad)
        return res, grad, rgrad

    def step(self, x: Variable, n: int, out_grad=None):
        """bench/mlp.go:29-41."""
        res = self.layers.batch(x, n)
        m = self.layer_sizes[-1] * n
        u = np.ones(m, dtype=F64) if out_grad is None else vec(out_grad).copy()
        grad = new_gradient(self.variables)
        res.propagate_gradient(u, grad)
        return res, grad


# --------------------------------------------------------------------------------------
# bench/lstm.go: lstmBlock + lstmNet, one lane at a time exactly as the reference runs it
# (the reference bench is single-sample; a batch of B lanes is B independent runs whose
# parameter gradients add -- the same contract seqfunc.MapBatcher gives, map_batcher.go:26-34)
# --------------------------------------------------------------------------------------
class LSTMNet:
    """bench/lstm.go:73-241.  Parameters in the order
    [InWeights, InBiases, InGateWeights, InGateBiases, ForgetGateWeights, ForgetGateBiases,
     OutputGate, OutputGateBiases, OutputWeights, OutputBiases]."""

    def __init__(self, input_size, hidden, output_size, seed=123, rseed=125, weight_scale=1.0):
        self.I, self.H, self.O = input_size, hidden, output_size
        cols = input_size + hidden
        self.variables = []
        for k in range(5):  # bench/lstm.go:102-127 generateLayer: U[-1,1) weights and biases
            rows = hidden if k < 4 else output_size
            self.variables.append(Variable(splitmix_uniform(seed + 1000 * k, rows * cols) * weight_scale))
            self.variables.append(Variable(splitmix_uniform(seed + 1000 * k + 1, rows)))
        v = self.variables
        self.lin = [LinTran(v[2 * k], hidden if k < 4 else output_size, cols) for k in range(5)]
        self.bias = [LinAdd(v[2 * k + 1]) for k in range(5)]
        self.rvec = {p: splitmix_uniform(rseed + 1000 * j, len(p.vector)) for j, p in enumerate(self.variables)}

    # -- non-R: bench/lstm.go:139-198, 215-234 ---------------------------------------------
    def run(self, inputs, upstreams):
        """inputs: list of T vectors (I); upstreams: list of T vectors (O).  Returns (outputs, grad)."""
        s = Sigmoid()
        in_states, out_states, out_state_vars, outputs = [], [], [], []
        for x in inputs:
            in_state = Variable(out_states[-1].output() if out_states else np.zeros(self.H))
            xv = Variable(x)
            inp = concat(in_state, xv)                                      # lstm.go:140
            new = s.apply(self.bias[0].apply(self.lin[0].apply(inp)))       # inState   :143
            mask = s.apply(self.bias[1].apply(self.lin[1].apply(inp)))      # inMask    :144
            forget = s.apply(self.bias[2].apply(self.lin[2].apply(inp)))    # forgetMask:145
            out_state = add(mul(forget, in_state), mul(mask, new))          # :147-150
            osv = Variable(out_state.output())                              # :183
            gate = s.apply(self.bias[3].apply(self.lin[3].apply(concat(in_state, xv))))   # :187-189
            masked = mul(osv, gate)                                         # :191
            res = s.apply(self.bias[4].apply(self.lin[4].apply(concat(masked, xv))))      # :193-194
            in_states.append(in_state); out_states.append(out_state); out_state_vars.append(osv); outputs.append(res)
        grad = new_gradient(self.variables)
        state_grad = np.zeros(self.H)
        for t in range(len(inputs) - 1, -1, -1):                            # :215-234
            grad[out_state_vars[t]] = np.zeros(self.H)
            grad[in_states[t]] = np.zeros(self.H)
            outputs[t].propagate_gradient(vec(upstreams[t]).copy(), grad)
            state_grad = state_grad + grad.pop(out_state_vars[t])
            out_states[t].propagate_gradient(state_grad, grad)
            state_grad = grad.pop(in_states[t])
        return [o.output() for o in outputs], grad

    # -- R: the same network through ConcatR/MulR/AddR/ApplyR, state R-derivative carried the way
    #    seqfunc.MapRBatcher carries it (RVariable{value, R-value}, seqfunc/map_batcher.go:85-91) ----
    def run_r(self, inputs, upstreams, upstreams_r=None):
        s = Sigmoid()
        rv = self.rvec

        def rvar(v, r):
            out = RVariable(v, {})
            out.routput_vec = r
            return out
        in_states, out_states, out_state_vars, outputs = [], [], [], []
        for x in inputs:
            if out_states:
                in_state = rvar(Variable(out_states[-1].output()), out_states[-1].routput())
            else:
                in_state = rvar(Variable(np.zeros(self.H)), np.zeros(self.H))
            xv = RVariable(Variable(x), rv)                                  # input not in rv: zero R
            inp = concat_r(in_state, xv)
            new = s.apply_r(rv, self.bias[0].apply_r(rv, self.lin[0].apply_r(rv, inp)))
            mask = s.apply_r(rv, self.bias[1].apply_r(rv, self.lin[1].apply_r(rv, inp)))
            forget = s.apply_r(rv, self.bias[2].apply_r(rv, self.lin[2].apply_r(rv, inp)))
            out_state = add_r(mul_r(forget, in_state), mul_r(mask, new))
            osv = rvar(Variable(out_state.output()), out_state.routput())
            gate = s.apply_r(rv, self.bias[3].apply_r(rv, self.lin[3].apply_r(rv, concat_r(in_state, xv))))
            masked = mul_r(osv, gate)
            res = s.apply_r(rv, self.bias[4].apply_r(rv, self.lin[4].apply_r(rv, concat_r(masked, xv))))
            in_states.append(in_state); out_states.append(out_state); out_state_vars.append(osv); outputs.append(res)
        grad, rgrad = new_gradient(self.variables), new_rgradient(self.variables)
        sg, sgr = np.zeros(self.H), np.zeros(self.H)
        T = len(inputs)
        for t in range(T - 1, -1, -1):
            for m in (grad, rgrad):
                m[out_state_vars[t].variable] = np.zeros(self.H)
                m[in_states[t].variable] = np.zeros(self.H)
            ur = np.zeros(self.O) if upstreams_r is None else vec(upstreams_r[t]).copy()
            outputs[t].propagate_rgradient(vec(upstreams[t]).copy(), ur, rgrad, grad)
            sg = sg + grad.pop(out_state_vars[t].variable)
            sgr = sgr + rgrad.pop(out_state_vars[t].variable)
            out_states[t].propagate_rgradient(sg, sgr, rgrad, grad)
            sg, sgr = grad.pop(in_states[t].variable), rgrad.pop(in_states[t].variable)
        return [o.output() for o in outputs], [o.routput() for o in outputs], grad, rgrad


# ======================================================================================
# SURVEY 8(f) "next" rows: the callers / cost heads either side of the dense path
# ======================================================================================
# math_funcs.go: Log, LogSigmoid, Sin, Cos, Norm, Softmax
# --------------------------------------------------------------------------------------
class Log(_ElementwiseFunc):  # math_funcs.go:99-189
    def apply(self, inp):
        return _LogResult(np.log(inp.output()), inp)

    def apply_r(self, rv, inp):
        x = inp.output()
        return _LogRResult(np.log(x), inp.routput() / x, inp)


class _LogResult(_UnaryResult):
    def propagate_gradient(self, upstream, grad):  # math_funcs.go:144-153
        if not self.input.constant(grad):
            upstream *= 1 / self.input.output()
            self.input.propagate_gradient(upstream, grad)


class _LogRResult(_UnaryRResult):
    def propagate_rgradient(self, upstream, upstream_r, rgrad, grad):  # math_funcs.go:173-189
        if self.input.constant(rgrad, grad):
            return
        x, xr = self.input.output(), self.input.routput()
        u, ur = upstream.copy(), upstream_r.copy()
        upstream[:] = u / x
        upstream_r[:] = ur / x - u * xr / (x * x)
        self.input.propagate_rgradient(upstream, upstream_r, rgrad, grad)


class LogSigmoid(_ElementwiseFunc):  # math_funcs.go:376-477
    @staticmethod
    def _ls(x):  # math_funcs.go:399-410: the branch avoids the log of a big number
        with np.errstate(over="ignore"):
            return np.where(x > 0, -np.log(1 + np.exp(-x)), x - np.log(1 + np.exp(x)))

    def apply(self, inp):
        return _LogSigmoidResult(self._ls(inp.output()), inp)

    def apply_r(self, rv, inp):
        x = inp.output()
        return _LogSigmoidRResult(self._ls(x), inp.routput() / (1 + np.exp(x)), inp)


class _LogSigmoidResult(_UnaryResult):
    def propagate_gradient(self, upstream, grad):  # math_funcs.go:433-440
        if not self.input.constant(grad):
            upstream /= 1 + np.exp(self.input.output())
            self.input.propagate_gradient(upstream, grad)


class _LogSigmoidRResult(_UnaryRResult):
    def propagate_rgradient(self, upstream, upstream_r, rgrad, grad):  # math_funcs.go:460-477
        if self.input.constant(rgrad, grad):
            return
        x, xr = self.input.output(), self.input.routput()
        u, ur = upstream.copy(), upstream_r.copy()
        ex = np.exp(x)
        s = 1 / (1 + ex)
        upstream[:] = u * s
        upstream_r[:] = ur * s - u * s * s * ex * xr
        self.input.propagate_rgradient(upstream, upstream_r, rgrad, grad)


class Sin(_ElementwiseFunc):  # math_funcs.go:511-597
    def apply(self, inp):
        return _SinResult(np.sin(inp.output()), inp)

    def apply_r(self, rv, inp):
        x = inp.output()
        return _SinRResult(np.sin(x), np.cos(x) * inp.routput(), inp)


class _SinResult(_UnaryResult):
    def propagate_gradient(self, upstream, grad):  # math_funcs.go:556-563
        if not self.input.constant(grad):
            upstream *= np.cos(self.input.output())
            self.input.propagate_gradient(upstream, grad)


class _SinRResult(_UnaryRResult):
    def propagate_rgradient(self, upstream, upstream_r, rgrad, grad):  # math_funcs.go:583-597
        if self.input.constant(rgrad, grad):
            return
        c = np.cos(self.input.output())
        cd = -self.output_vec * self.input.routput()
        upstream_r[:] = upstream_r * c + upstream * cd
        upstream *= c
        self.input.propagate_rgradient(upstream, upstream_r, rgrad, grad)


class Cos(_ElementwiseFunc):  # math_funcs.go:599-607
    def apply(self, inp):
        return Sin().apply(add_scaler(inp, np.pi / 2))

    def apply_r(self, rv, inp):
        return Sin().apply_r(rv, add_scaler_r(inp, np.pi / 2))


class Tanh(_ElementwiseFunc):
    """EXTENSION named by BASELINE.json's north_star; the reference has no Tanh.
    Defined by the same pattern as Sigmoid (math_funcs.go:286-374) with t' = 1 - t^2."""

    def apply(self, inp):
        return _TanhResult(np.tanh(inp.output()), inp)

    def apply_r(self, rv, inp):
        t = np.tanh(inp.output())
        return _TanhRResult(t, (1 - t * t) * inp.routput(), inp)


class _TanhResult(_UnaryResult):
    def propagate_gradient(self, upstream, grad):
        if not self.input.constant(grad):
            t = self.output_vec
            upstream *= 1 - t * t
            self.input.propagate_gradient(upstream, grad)


class _TanhRResult(_UnaryRResult):
    def propagate_rgradient(self, upstream, upstream_r, rgrad, grad):
        if self.input.constant(rgrad, grad):
            return
        t, tr = self.output_vec, self.routput_vec
        u = upstream.copy()
        upstream[:] = (1 - t * t) * u
        upstream_r[:] = (1 - t * t) * upstream_r - 2 * t * tr * u
        self.input.propagate_rgradient(upstream, upstream_r, rgrad, grad)


class Norm(_ElementwiseFunc):  # math_funcs.go:201-284
    def apply(self, inp):
        return _NormResult(vec([np.sqrt(_seq_sum(inp.output() ** 2))]), inp)

    def apply_r(self, rv, inp):
        x = inp.output()
        out = np.sqrt(_seq_sum(x ** 2))
        return _NormRResult(vec([out]), vec([(1 / out) * _seq_sum(x * inp.routput())]), inp)


class _NormResult(_UnaryResult):
    def propagate_gradient(self, u, g):  # math_funcs.go:238-247
        if not self.constant(g):
            self.input.propagate_gradient(self.input.output() * (u[0] / self.output_vec[0]), g)


class _NormRResult(_UnaryRResult):
    def propagate_rgradient(self, u, ur, rg, g):  # math_funcs.go:267-284
        if self.constant(rg, g):
            return
        o, ro = self.output_vec[0], self.routput_vec[0]
        s = u[0] / o
        sr = -u[0] * ro / (o * o) + ur[0] / o
        x = self.input.output()
        self.input.propagate_rgradient(x * s, x * sr + self.input.routput() * s, rg, g)


class Softmax(_ElementwiseFunc):  # math_funcs.go:479-509 (whole-vector; composed exactly as the reference)
    def __init__(self, temperature: float = 0.0):
        self.temperature = float(temperature)

    def apply(self, inp):
        scaled = inp
        if self.temperature != 0 and self.temperature != 1:
            scaled = scale(inp, 1 / self.temperature)
        exps = Exp().apply(scaled)
        return pool(exps, lambda e: scale_first(e, inverse(sum_all(e))))

    def apply_r(self, rv, inp):
        scaled = inp
        if self.temperature != 0 and self.temperature != 1:
            scaled = scale_r(inp, 1 / self.temperature)
        exps = Exp().apply_r(rv, scaled)
        # math_funcs.go:506-507 sums the OUTER exps, not the pooled input: kept as written
        return pool_r(exps, lambda e: scale_first_r(exps, inverse_r(sum_all_r(exps))))

    # the batched form is RFuncBatcher{Softmax} (batcher.go:57-84): one softmax per sample
    def batch(self, inp, n):
        return RFuncBatcher(self).batch(inp, n)

    def batch_r(self, rv, inp, n):
        return RFuncBatcher(self).batch_r(rv, inp, n)


# --------------------------------------------------------------------------------------
# arithmetic.go: AddScaler, Div, ScaleFirst, AddFirst, Inverse, Pow, log-domain sums
# --------------------------------------------------------------------------------------
class _AddScaler(_UnaryResult):  # arithmetic.go:184-214
    def propagate_gradient(self, upstream, grad):
        self.input.propagate_gradient(upstream, grad)


class _AddScalerR(_UnaryRResult):  # arithmetic.go:216-252
    def propagate_rgradient(self, upstream, upstream_r, rgrad, grad):
        self.input.propagate_rgradient(upstream, upstream_r, rgrad, grad)


def add_scaler(r, f):
    return _AddScaler(r.output() + f, r)


def add_scaler_r(r, f):
    return _AddScalerR(r.output() + f, r.routput(), r)


class _Quotient(_Binary):  # arithmetic.go:378-423
    def propagate_gradient(self, u, g):
        den = self.r2.output()
        if not self.r1.constant(g):
            self.r1.propagate_gradient(u / den, g)
        if not self.r2.constant(g):
            u *= -self.output_vec / den
            self.r2.propagate_gradient(u, g)


def div(a, b):
    return _Quotient(a.output() / b.output(), a, b)


def div_r(a, b):  # arithmetic.go:425-427
    return mul_r(a, inverse_r(b))


class _ScaleFirst:  # arithmetic.go:495-533
    def __init__(self, inp, scaler):
        self.input, self.scaler = inp, scaler
        self.output_vec = inp.output() * scaler.output()[0]

    def output(self):
        return self.output_vec

    def constant(self, g):
        return self.input.constant(g) and self.scaler.constant(g)

    def propagate_gradient(self, upstream, grad):
        if not self.scaler.constant(grad):
            down = np.zeros(len(self.scaler.output()), dtype=F64)
            down[0] = np.dot(upstream, self.input.output())
            self.scaler.propagate_gradient(down, grad)
        if not self.input.constant(grad):
            upstream *= self.scaler.output()[0]
            self.input.propagate_gradient(upstream, grad)


class _ScaleFirstR:  # arithmetic.go:535-593
    def __init__(self, inp, scaler):
        self.input, self.scaler = inp, scaler
        f, fr = scaler.output()[0], scaler.routput()[0]
        self.output_vec = inp.output() * f
        self.routput_vec = inp.output() * fr + inp.routput() * f

    def output(self):
        return self.output_vec

    def routput(self):
        return self.routput_vec

    def constant(self, rg, g):
        return self.input.constant(rg, g) and self.scaler.constant(rg, g)

    def propagate_rgradient(self, upstream, upstream_r, rgrad, grad):
        if not self.scaler.constant(rgrad, grad):
            down = np.zeros(len(self.scaler.output()), dtype=F64)
            down_r = np.zeros(len(self.scaler.output()), dtype=F64)
            down[0] = np.dot(upstream, self.input.output())
            down_r[0] = np.dot(upstream_r, self.input.output()) + np.dot(upstream, self.input.routput())
            self.scaler.propagate_rgradient(down, down_r, rgrad, grad)
        if not self.input.constant(rgrad, grad):
            f, fr = self.scaler.output()[0], self.scaler.routput()[0]
            upstream_r[:] = upstream_r * f + upstream * fr
            upstream *= f
            self.input.propagate_rgradient(upstream, upstream_r, rgrad, grad)


def scale_first(inp, scaler):
    return _ScaleFirst(inp, scaler)


def scale_first_r(inp, scaler):
    return _ScaleFirstR(inp, scaler)


class _AddFirst:  # arithmetic.go:595-639
    def __init__(self, v1, v2):
        self.input, self.scaler = v1, v2
        self.output_vec = v1.output() + v2.output()[0]

    def output(self):
        return self.output_vec

    def constant(self, g):
        return self.input.constant(g) and self.scaler.constant(g)

    def propagate_gradient(self, upstream, g):
        if not self.scaler.constant(g):
            su = np.zeros(len(self.scaler.output()), dtype=F64)
            su[0] = _seq_sum(upstream)
            self.scaler.propagate_gradient(su, g)
        if not self.input.constant(g):
            self.input.propagate_gradient(upstream, g)


class _AddFirstR:  # arithmetic.go:641-694
    def __init__(self, v1, v2):
        self.input, self.scaler = v1, v2
        self.output_vec = v1.output() + v2.output()[0]
        self.routput_vec = v1.routput() + v2.routput()[0]

    def output(self):
        return self.output_vec

    def routput(self):
        return self.routput_vec

    def constant(self, rg, g):
        return self.input.constant(rg, g) and self.scaler.constant(rg, g)

    def propagate_rgradient(self, upstream, upstream_r, rg, g):
        if not self.scaler.constant(rg, g):
            su = np.zeros(len(self.scaler.output()), dtype=F64)
            sur = np.zeros(len(self.scaler.output()), dtype=F64)
            su[0], sur[0] = _seq_sum(upstream), _seq_sum(upstream_r)
            self.scaler.propagate_rgradient(su, sur, rg, g)
        if not self.input.constant(rg, g):
            self.input.propagate_rgradient(upstream, upstream_r, rg, g)


def add_first(v1, v2):
    return _AddFirst(v1, v2)


def add_first_r(v1, v2):
    return _AddFirstR(v1, v2)


class _Inverse(_UnaryResult):  # arithmetic.go:772-807
    def propagate_gradient(self, upstream, grad):
        if not self.input.constant(grad):
            o = self.output_vec
            upstream *= -o * o
            self.input.propagate_gradient(upstream, grad)


class _InverseR(_UnaryRResult):  # arithmetic.go:809-867
    def propagate_rgradient(self, upstream, upstream_r, rgrad, grad):
        if self.input.constant(rgrad, grad):
            return
        o = self.output_vec
        sq = -(o * o)
        u = upstream.copy()
        upstream[:] = sq * u
        upstream_r[:] = sq * upstream_r + -2 * (sq * o) * self.input.routput() * u
        self.input.propagate_rgradient(upstream, upstream_r, rgrad, grad)


def inverse(r):
    return _Inverse(1 / r.output(), r)


def inverse_r(r):
    rec = 1 / r.output()
    return _InverseR(rec, -(rec * rec) * r.routput(), r)


class _Pow(_UnaryResult):  # arithmetic.go:869-906
    power = 1.0

    def constant(self, g):
        return self.power != 0 and self.input.constant(g)

    def propagate_gradient(self, upstream, grad):
        if not self.constant(grad):
            if self.power != 1:
                upstream *= self.power * np.power(self.input.output(), self.power - 1)
            self.input.propagate_gradient(upstream, grad)


class _PowR(_UnaryRResult):  # arithmetic.go:908-966
    power = 1.0

    def constant(self, rg, g):
        return self.power != 0 and self.input.constant(rg, g)

    def propagate_rgradient(self, upstream, upstream_r, rgrad, grad):
        if not self.constant(rgrad, grad):
            if self.power != 1:
                x, xr, p = self.input.output(), self.input.routput(), self.power
                d1 = p * np.power(x, p - 1)
                d2 = p * (p - 1) * np.power(x, p - 2) * xr
                u = upstream.copy()
                upstream[:] = u * d1
                upstream_r[:] = upstream_r * d1 + u * d2
            self.input.propagate_rgradient(upstream, upstream_r, rgrad, grad)


def pow_(r, p):
    res = _Pow(np.power(r.output(), p), r)
    res.power = float(p)
    return res


def pow_r(r, p):
    x = r.output()
    rout = p * np.power(x, p - 1) * r.routput() if p != 0 else np.zeros(len(x), dtype=F64)
    res = _PowR(np.power(x, p), rout, r)
    res.power = float(p)
    return res


def _max_abs(x):
    return float(np.max(np.abs(x))) if len(x) else 0.0


def add_log_domain(v1, v2):  # arithmetic.go:1056-1062 (maxVal is the max ABSOLUTE value, a constant)
    m = max(_max_abs(v1.output()), _max_abs(v2.output()))
    e1 = Exp().apply(add_scaler(v1, -m))
    e2 = Exp().apply(add_scaler(v2, -m))
    return add_scaler(Log().apply(add(e1, e2)), m)


def add_log_domain_r(v1, v2):  # arithmetic.go:1065-1072
    m = max(_max_abs(v1.output()), _max_abs(v2.output()))
    e1 = Exp().apply_r({}, add_scaler_r(v1, -m))
    e2 = Exp().apply_r({}, add_scaler_r(v2, -m))
    return add_scaler_r(Log().apply_r({}, add_r(e1, e2)), m)


def sum_all_log_domain(v):  # arithmetic.go:1077-1082
    m = _max_abs(v.output())
    return add_scaler(Log().apply(sum_all(Exp().apply(add_scaler(v, -m)))), m)


def sum_all_log_domain_r(v):  # arithmetic.go:1085-1090
    m = _max_abs(v.output())
    return add_scaler_r(Log().apply_r(None, sum_all_r(Exp().apply_r(None, add_scaler_r(v, -m)))), m)


def squared_error(actual, expected):
    """north_star's "SquaredError": 0.5*||a - y||^2 = Scale(SquaredNorm(Sub(a, y)), .5)."""
    return scale(SquaredNorm().apply(sub(actual, expected)), 0.5)


def squared_error_r(actual, expected):
    return scale_r(SquaredNorm().apply_r({}, sub_r(actual, expected)), 0.5)


def split(n, inp):  # slices.go:310-321
    if len(inp.output()) % n != 0:
        raise ValueError("count does not divide input length")
    size = len(inp.output()) // n
    return [slice_(inp, i * size, (i + 1) * size) for i in range(n)]


def split_r(n, inp):  # slices.go:324-335
    if len(inp.output()) % n != 0:
        raise ValueError("count does not divide input length")
    size = len(inp.output()) // n
    return [slice_r(inp, i * size, (i + 1) * size) for i in range(n)]


# --------------------------------------------------------------------------------------
# pool.go: Pool / PoolAll (+R) -- the temp-variable protocol
# --------------------------------------------------------------------------------------
class _Pooled:  # pool.go:94-136
    def __init__(self, ins, f):
        self.inputs = list(ins)
        self.pool_vars = [Variable(r.output()) for r in ins]
        for pv, r in zip(self.pool_vars, ins):
            pv.vector = r.output()  # aliased like &Variable{Vector: in.Output()}
        self.f_output = f(list(self.pool_vars))

    def output(self):
        return self.f_output.output()

    def constant(self, g):
        if not self.f_output.constant(g):
            return False
        for pv, r in zip(self.pool_vars, self.inputs):
            if not self.f_output.constant({pv: np.zeros(0)}) and not r.constant(g):
                return False
        return True

    def propagate_gradient(self, upstream, grad):
        consts = []
        for pv, r in zip(self.pool_vars, self.inputs):
            c = self.f_output.constant({pv: np.zeros(0)}) or r.constant(grad)
            consts.append(c)
            if not c:
                grad[pv] = np.zeros(len(pv.vector), dtype=F64)
        self.f_output.propagate_gradient(upstream, grad)
        ups = [None] * len(self.pool_vars)
        for i, pv in enumerate(self.pool_vars):
            if not consts[i]:
                ups[i] = grad.pop(pv)
        for i, c in enumerate(consts):
            if not c:
                self.inputs[i].propagate_gradient(ups[i], grad)


class _PooledR:  # pool.go:138-195
    def __init__(self, ins, f):
        self.inputs = list(ins)
        self.pool_vars = []
        pool_res = []
        for r in ins:
            pv = Variable(r.output())
            pv.vector = r.output()
            self.pool_vars.append(pv)
            pool_res.append(RVariable(pv, {pv: r.routput()}))
        self.f_output = f(pool_res)

    def output(self):
        return self.f_output.output()

    def routput(self):
        return self.f_output.routput()

    def constant(self, rg, g):
        if not self.f_output.constant(rg, g):
            return False
        for pv, r in zip(self.pool_vars, self.inputs):
            if not self.f_output.constant({pv: np.zeros(0)}, None) and not r.constant(rg, g):
                return False
        return True

    def propagate_rgradient(self, upstream, upstream_r, rgrad, grad):
        if grad is None:
            grad = {}
        consts = []
        for pv, r in zip(self.pool_vars, self.inputs):
            c = self.f_output.constant({pv: np.zeros(0)}, None) or r.constant(rgrad, grad)
            consts.append(c)
            if not c:
                grad[pv] = np.zeros(len(pv.vector), dtype=F64)
                rgrad[pv] = np.zeros(len(pv.vector), dtype=F64)
        self.f_output.propagate_rgradient(upstream, upstream_r, rgrad, grad)
        ups, ups_r = [None] * len(consts), [None] * len(consts)
        for i, pv in enumerate(self.pool_vars):
            if not consts[i]:
                ups[i], ups_r[i] = grad.pop(pv), rgrad.pop(pv)
        for i, c in enumerate(consts):
            if not c:
                self.inputs[i].propagate_rgradient(ups[i], ups_r[i], rgrad, grad)


def pool(r, f):  # pool.go:21-25
    return _Pooled([r], lambda ins: f(ins[0]))


def pool_r(r, f):  # pool.go:28-32
    return _PooledR([r], lambda ins: f(ins[0]))


def pool_all(ins, f):  # pool.go:38-50
    return _Pooled(ins, f)


def pool_all_r(ins, f):  # pool.go:53-68
    return _PooledR(ins, f)


# --------------------------------------------------------------------------------------
# fold.go: Fold / FoldR
# --------------------------------------------------------------------------------------
class _Fold:  # fold.go:5-84
    def __init__(self, state, ins, f):
        self.pool, self.intermediate = [], []
        for r in ins:
            self.intermediate.append(state)
            pv = Variable(state.output())
            pv.vector = state.output()
            self.pool.append(pv)
            state = f(pv, r)
        self.final = state

    def output(self):
        return self.final.output()

    def constant(self, g):  # fold.go:35-54
        if not self.final.constant(g):
            return False
        for i in range(len(self.pool) - 1, -1, -1):
            if not self.intermediate[i].constant(g):
                return False
            if i > 0:
                p = self.pool[i - 1]
                g[p] = np.zeros(len(p.vector), dtype=F64)
                c = self.intermediate[i].constant(g)
                del g[p]
                if c:
                    return True
        return True

    def propagate_gradient(self, u, g):  # fold.go:56-84
        if not self.pool:
            self.final.propagate_gradient(u, g)
            return
        state_up = u
        for i in range(len(self.pool), 0, -1):
            res = self.final if i == len(self.pool) else self.intermediate[i]
            p = self.pool[i - 1]
            g[p] = np.zeros(len(p.vector), dtype=F64)
            if res.constant(g):
                del g[p]
                break
            res.propagate_gradient(state_up, g)
            state_up = g.pop(p)
        if self.intermediate:
            self.intermediate[0].propagate_gradient(state_up, g)


class _FoldR:  # fold.go:86-178
    def __init__(self, state, ins, f):
        self.pool, self.intermediate = [], []
        for r in ins:
            self.intermediate.append(state)
            pv = Variable(state.output())
            pv.vector = state.output()
            self.pool.append(pv)
            state = f(RVariable(pv, {pv: state.routput()}), r)
        self.final = state

    def output(self):
        return self.final.output()

    def routput(self):
        return self.final.routput()

    def constant(self, rg, g):  # fold.go:113-139
        if not self.final.constant(rg, g):
            return False
        for i in range(len(self.pool) - 1, -1, -1):
            if not self.intermediate[i].constant(rg, g):
                return False
            if i > 0:
                p = self.pool[i - 1]
                if g is not None:
                    g[p] = np.zeros(len(p.vector), dtype=F64)
                rg[p] = np.zeros(len(p.vector), dtype=F64)
                c = self.intermediate[i].constant(rg, g)
                if g is not None:
                    del g[p]
                del rg[p]
                if c:
                    return True
        return True

    def propagate_rgradient(self, u, ur, rg, g):  # fold.go:141-178
        if not self.pool:
            self.final.propagate_rgradient(u, ur, rg, g)
            return
        if g is None:
            g = {}
        state_up, state_up_r = u, ur
        for i in range(len(self.pool), 0, -1):
            res = self.final if i == len(self.pool) else self.intermediate[i]
            p = self.pool[i - 1]
            g[p] = np.zeros(len(p.vector), dtype=F64)
            rg[p] = np.zeros(len(p.vector), dtype=F64)
            if res.constant(rg, g):
                del g[p]
                del rg[p]
                break
            res.propagate_rgradient(state_up, state_up_r, rg, g)
            state_up, state_up_r = g.pop(p), rg.pop(p)
        if self.intermediate:
            self.intermediate[0].propagate_rgradient(state_up, state_up_r, rg, g)


def fold(state, ins, f):
    return _Fold(state, ins, f)


def fold_r(state, ins, f):
    return _FoldR(state, ins, f)


# --------------------------------------------------------------------------------------
# lin_alg.go: MatMulVecs -- LinTran with a computed matrix
# --------------------------------------------------------------------------------------
class _MatMul:  # lin_alg.go:9-69
    def __init__(self, mat, rows, cols, vecs):
        if len(mat.output()) != rows * cols:
            raise ValueError("invalid matrix data size")
        if len(vecs.output()) % cols != 0:
            raise ValueError("invalid vecs size")
        n = len(vecs.output()) // cols
        self.mat_in = mat
        self.mat_var = Variable(mat.output())
        self.mat_var.vector = mat.output()
        self.res = LinTran(self.mat_var, rows, cols).batch(vecs, n)

    def output(self):
        return self.res.output()

    def constant(self, g):
        return self.res.constant(g) and self.mat_in.constant(g)

    def propagate_gradient(self, upstream, grad):
        mat_const = self.mat_in.constant(grad)
        if not mat_const:
            grad[self.mat_var] = np.zeros(len(self.mat_var.vector), dtype=F64)
        self.res.propagate_gradient(upstream, grad)
        if not mat_const:
            self.mat_in.propagate_gradient(grad.pop(self.mat_var), grad)


class _MatMulR:  # lin_alg.go:71-133
    def __init__(self, mat, rows, cols, vecs):
        if len(mat.output()) != rows * cols:
            raise ValueError("invalid matrix data size")
        if len(vecs.output()) % cols != 0:
            raise ValueError("invalid vecs size")
        n = len(vecs.output()) // cols
        self.mat_in = mat
        self.mat_var = Variable(mat.output())
        self.mat_var.vector = mat.output()
        self.res = LinTran(self.mat_var, rows, cols).batch_r({self.mat_var: mat.routput()}, vecs, n)

    def output(self):
        return self.res.output()

    def routput(self):
        return self.res.routput()

    def constant(self, rg, g):
        return self.res.constant(rg, g) and self.mat_in.constant(rg, g)

    def propagate_rgradient(self, upstream, upstream_r, rgrad, grad):
        mat_const = self.mat_in.constant(rgrad, grad)
        if not mat_const:
            if grad is None:
                grad = {}
            grad[self.mat_var] = np.zeros(len(self.mat_var.vector), dtype=F64)
            rgrad[self.mat_var] = np.zeros(len(self.mat_var.vector), dtype=F64)
        self.res.propagate_rgradient(upstream, upstream_r, rgrad, grad)
        if not mat_const:
            self.mat_in.propagate_rgradient(grad.pop(self.mat_var), rgrad.pop(self.mat_var), rgrad, grad)


def mat_mul_vecs(mat, rows, cols, vecs):
    return _MatMul(mat, rows, cols, vecs)


def mat_mul_vecs_r(mat, rows, cols, vecs):
    return _MatMulR(mat, rows, cols, vecs)


mat_mul_vec, mat_mul_vec_r = mat_mul_vecs, mat_mul_vecs_r


def result_len(r) -> int:
    return len(r.output())
