"""Host-side mirror of the reference's operator interface for the dense differentiation path,
on top of the C ABI (include/autofunc_b200.h).

Same names, argument meaning and error behaviour as the Go interfaces it stands in for
(Go itself is not available in the build image; go/b200 holds the cgo shim of the same shape):

    Func{Apply}            func.go:7-9        -> .apply(in)
    RFunc{ApplyR}          func.go:24-27      -> .apply_r(rv, in)
    Batcher{Batch}         batcher.go:5-20    -> .batch(in, n)
    RBatcher{BatchR}       batcher.go:24-29   -> .batch_r(rv, in, n)
    Result / RResult       result.go:10-77    -> .output() .routput() .constant() .propagate_gradient()
                                                 .propagate_rgradient(upstream, upstream_r, rgrad, grad)
    Variable / RVariable   variable.go:19-123
    Gradient / RGradient / RVector  gradient.go:11,62,93 -> dict keyed by Variable identity

Design: values, R-values, upstreams and parameter-gradient accumulators stay in HBM between
operators.  A Result produced here holds device buffers and only downloads when `.output()` is
called.  Device Results recognise each other (as slices.go:148 recognises *Variable) and hand
device buffers down the backward recursion; host arrays are touched only when a host caller
enters (`propagate_*` with numpy upstreams) and when parameter gradients are flushed with `+=`
into the caller's Gradient maps, which happens before that outermost call returns so the
reference's temp-variable protocol (pool.go:114-136) keeps working.

There is no CPU fallback: every op goes through libautofunc_b200.so and raises if it is absent.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref

import numpy as np

from . import _lib
from ._lib import AutofuncError, ShapeError, check  # noqa: F401  (re-exported)

F64 = np.float64


# --------------------------------------------------------------------------------------------
# context and device buffers
# --------------------------------------------------------------------------------------------
class Context:
    """One device + one stream (SURVEY 8b threading)."""

    def __init__(self, device: int | None = None, stream: int | None = None):
        self.lib = _lib.load()
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        h = C.c_void_p()
        if stream is None:
            check(self.lib.af_ctx_create(device, C.byref(h)))
        else:
            check(self.lib.af_ctx_create_on_stream(device, C.c_void_p(stream), C.byref(h)))
        self.h = h
        self.device = device
        self._pool: dict[int, list[int]] = {}
        self._closed = False
        # A/B runs of the tuning knobs over unmodified callers (tests, tools): AF_OPTIONS="sk_rotate=1,big_cfg=0"
        for item in filter(None, os.environ.get("AF_OPTIONS", "").split(",")):
            name, _, value = item.partition("=")
            check(self.lib.af_set_option(self.h, name.strip().encode(), int(value)))

    # size-bucketed free lists: operators allocate their outputs per call as the reference does
    # with make() (lin_tran.go:81), but cudaMalloc/cudaFree would synchronise the stream
    def _alloc(self, n: int) -> tuple[int, int]:
        cap = max(32, 1 << (max(n, 1) - 1).bit_length()) if n <= (1 << 20) else (n + (1 << 18) - 1) // (1 << 18) * (1 << 18)
        fl = self._pool.get(cap)
        if fl:
            return fl.pop(), cap
        p = C.c_void_p()
        check(self.lib.af_malloc(self.h, cap, C.byref(p)))
        return p.value, cap

    def _release(self, ptr: int, cap: int) -> None:
        if not self._closed:
            self._pool.setdefault(cap, []).append(ptr)

    def empty(self, n: int) -> "DeviceVec":
        return DeviceVec(self, int(n))

    def zeros(self, n: int) -> "DeviceVec":
        v = DeviceVec(self, int(n))
        check(self.lib.af_zero(self.h, v.ptr, v.n))
        return v

    def upload(self, arr) -> "DeviceVec":
        a = np.ascontiguousarray(arr, dtype=F64).reshape(-1)
        v = DeviceVec(self, a.size)
        check(self.lib.af_upload(self.h, v.ptr, a.ctypes.data, a.size))
        return v

    def sync(self) -> None:
        check(self.lib.af_sync(self.h))

    def launch_count(self) -> int:
        return int(self.lib.af_launch_count(self.h))

    def set_option(self, name: str, value: int) -> None:
        check(self.lib.af_set_option(self.h, name.encode(), int(value)))

    # ---- data parallelism (af_dp_*): Gradient.Add across ranks, NCCL behind the C ABI ----
    def dp_init(self, nranks: int, rank: int, unique_id: bytes) -> None:
        """Join the communicator identified by `unique_id` (the 128 bytes rank 0 got from parallel.dp_unique_id())."""
        buf = C.create_string_buffer(bytes(unique_id), _lib.AF_DP_ID_BYTES)
        check(self.lib.af_dp_init(self.h, int(nranks), int(rank), buf))

    def dp_info(self) -> tuple[int, int, int]:
        n, r, v = C.c_int(), C.c_int(), C.c_int()
        check(self.lib.af_dp_info(self.h, C.byref(n), C.byref(r), C.byref(v)))
        return n.value, r.value, v.value

    def dp_calls(self) -> int:
        return int(self.lib.af_dp_calls(self.h))

    def allreduce_sum(self, v: "DeviceVec") -> None:
        check(self.lib.af_allreduce_sum(self.h, v.ptr, len(v)))

    def set_gemm_path(self, mode: int) -> None:
        check(self.lib.af_set_gemm_path(self.h, mode))

    def record(self) -> "Graph":
        """`with ctx.record() as g:` records every ABI-level launch issued on this context inside the block into a
        CUDA graph instead of running it; `g.launch()` replays the sequence (same buffers, same sizes) as one
        submission.  For launch-bound chains: the reference's own benchmarks run n = 1 and n = 10."""
        return Graph(self)

    def trim(self) -> None:
        self.sync()
        for fl in self._pool.values():
            for p in fl:
                self.lib.af_free(self.h, p)
        self._pool.clear()

    def close(self) -> None:
        if not self._closed:
            self.trim()
            self._closed = True
            self.lib.af_ctx_destroy(self.h)


class Graph:
    """A recorded launch sequence (af_graph_begin / af_graph_end / af_graph_launch)."""

    def __init__(self, ctx: Context):
        self.ctx, self.h = ctx, None

    def __enter__(self):
        check(self.ctx.lib.af_graph_begin(self.ctx.h))
        return self

    def __exit__(self, exc_type, exc, tb):
        if exc_type is not None:  # abandon the recording, let the exception through
            self.ctx.lib.af_graph_end(self.ctx.h, None)
            return False
        h = C.c_void_p()
        check(self.ctx.lib.af_graph_end(self.ctx.h, C.byref(h)))
        self.h = h
        return False

    @property
    def launches(self) -> int:
        return int(self.ctx.lib.af_graph_launches(self.h)) if self.h else 0

    def launch(self) -> None:
        if self.h is None:
            raise AutofuncError(_lib.AF_EINVAL, "the graph was not recorded")
        check(self.ctx.lib.af_graph_launch(self.h))

    def close(self) -> None:
        if self.h is not None:
            self.ctx.lib.af_graph_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            if not self.ctx._closed:
                self.close()
        except Exception:
            pass


class DeviceVec:
    """A float64 vector resident in HBM."""

    __slots__ = ("ctx", "n", "ptr", "_cap", "_owner", "__weakref__")

    def __init__(self, ctx: Context, n: int, ptr: int | None = None, owner=None):
        self.ctx, self.n = ctx, n
        if ptr is None:
            self.ptr, self._cap = ctx._alloc(n)
            self._owner = None
        else:  # a view into someone else's allocation
            self.ptr, self._cap, self._owner = ptr, 0, owner

    def __del__(self):
        try:
            if self._owner is None and self._cap:
                self.ctx._release(self.ptr, self._cap)
        except Exception:
            pass

    def __len__(self):
        return self.n

    def view(self, start: int, end: int) -> "DeviceVec":
        return DeviceVec(self.ctx, end - start, self.ptr + 8 * start, owner=self)

    def numpy(self) -> np.ndarray:
        out = np.empty(self.n, dtype=F64)
        if self.n:
            check(self.ctx.lib.af_download(self.ctx.h, out.ctypes.data, self.ptr, self.n))
        return out

    def copy(self) -> "DeviceVec":
        v = DeviceVec(self.ctx, self.n)
        check(self.ctx.lib.af_copy(self.ctx.h, v.ptr, self.ptr, self.n))
        return v

    def __iadd__(self, other):  # `grad[v] += upstream` (variable.go:45) when the map's vectors live in HBM
        o = other if isinstance(other, DeviceVec) else self.ctx.upload(other)
        if o.n != self.n:
            raise ShapeError(_lib.AF_ESHAPE, "vector sizes do not match")
        check(self.ctx.lib.af_axpy(self.ctx.h, 1.0, o.ptr, self.ptr, self.n))
        return self


_default_ctx: Context | None = None


def context() -> Context:
    """The process-wide default context (cuda:LOCAL_RANK)."""
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context()
    return _default_ctx


def set_context(ctx: Context | None) -> None:
    global _default_ctx
    _default_ctx = ctx


def _p(v: "DeviceVec | None"):
    return None if v is None else v.ptr


# --------------------------------------------------------------------------------------------
# Variable / RVariable / Gradient maps
# --------------------------------------------------------------------------------------------
class Variable:
    """variable.go:19-21.  `.vector` is the host value (numpy float64); a device copy is cached and
    refreshed when the host value changes -- call `set()` / `mark_dirty()` after mutating large
    vectors in place (vectors up to 64Ki entries are compared automatically).

    The device copy can also be the authoritative one: pool variables wrap a Result's output without
    leaving HBM (`from_device`, the analogue of `&Variable{Vector: in.Output()}` at pool.go:42,
    fold.go:24, lin_alg.go:37), and `add_scaled_device` is gradient.go:40-52's Axpy on the device.
    `.vector` then downloads lazily, the first time a host caller looks."""

    __hash__ = object.__hash__
    _AUTO_CHECK = 1 << 16

    def __init__(self, v):
        self._vector = np.array(v, dtype=F64).reshape(-1)
        self.size = self._vector.size
        self._dev: DeviceVec | None = None
        self._snapshot: np.ndarray | None = None
        self._host_stale = False

    @classmethod
    def from_device(cls, d: DeviceVec) -> "Variable":
        v = cls.__new__(cls)
        v._vector, v.size, v._dev, v._snapshot, v._host_stale = None, d.n, d, None, True
        return v

    @property
    def vector(self) -> np.ndarray:
        if self._host_stale:
            h = self._dev.numpy()
            if self._vector is None:
                self._vector = h
            else:
                self._vector[:] = h
            self._host_stale = False
            self._snapshot = self._vector.copy() if self.size <= self._AUTO_CHECK else None
        return self._vector

    @vector.setter
    def vector(self, arr) -> None:
        self._vector = np.array(arr, dtype=F64).reshape(-1)
        self.size = self._vector.size
        self._dev, self._host_stale = None, False

    def __len__(self):
        return self.size

    def set(self, arr) -> None:
        self.vector[:] = arr
        self._dev = None

    def mark_dirty(self) -> None:
        _ = self.vector  # pull a device-side update down before the host copy becomes authoritative
        self._dev = None

    def device(self) -> DeviceVec:
        if self._host_stale:
            return self._dev
        small = self.size <= self._AUTO_CHECK
        if self._dev is not None and small and not np.array_equal(self._snapshot, self._vector):
            self._dev = None
        if self._dev is None:
            self._dev = context().upload(self._vector)
            self._snapshot = self._vector.copy() if small else None
        return self._dev

    def add_scaled_device(self, scale: float, g: DeviceVec) -> None:
        """v += scale * g without leaving HBM (gradient.go:40-52)."""
        d = self.device()
        if g.n != d.n:
            raise ShapeError(_lib.AF_ESHAPE, "vector sizes do not match")
        ctx = context()
        check(ctx.lib.af_axpy(ctx.h, float(scale), g.ptr, d.ptr, d.n))
        self._host_stale = True

    # Result interface ------------------------------------------------------------------
    def output(self):
        return self.vector

    def constant(self, g):
        return self not in g

    def propagate_gradient(self, upstream, grad):  # variable.go:43-47
        if self in grad:
            grad[self] += upstream if isinstance(grad[self], DeviceVec) else _host(upstream)

    # device-side protocol
    def _d(self):
        return self.device()

    def _bwd(self, sess, dU, grad):
        if self in grad:
            sess.accumulate(grad, self, dU)

    # wire format (variable.go:24-37,61-68)
    def serialize(self) -> bytes:
        return np.uint64(self.size).astype("<u8").tobytes() + self.vector.astype("<f8").tobytes()

    @staticmethod
    def deserialize(b: bytes) -> "Variable":
        # DeserializeVariable (variable.go:24-37) returns an error on a short read; so does this
        if len(b) < 8:
            raise ValueError(f"serialized Variable is truncated: {len(b)} bytes, the length prefix alone takes 8")
        n = int(np.frombuffer(b[:8], dtype="<u8")[0])
        if (len(b) - 8) // 8 < n:
            raise ValueError(f"serialized Variable is truncated: the prefix announces {n} values, {(len(b) - 8) // 8} present")
        return Variable(np.frombuffer(b[8:8 + 8 * n], dtype="<f8").copy())


class RVariable:
    """variable.go:73-123."""

    def __init__(self, v: Variable, rv: dict):
        self.variable = v
        self._has_r = v in rv
        r = rv[v] if self._has_r else None
        self._rdev: DeviceVec | None = r if isinstance(r, DeviceVec) else None
        self.routput_vec = np.asarray(r, dtype=F64) if (self._has_r and self._rdev is None) else None

    def output(self):
        return self.variable.vector

    def routput(self):
        if self.routput_vec is None:
            if self._rdev is not None:
                self.routput_vec = self._rdev.numpy()
            else:  # variable.go:85-91: fresh zeros
                self.routput_vec = np.zeros(self.variable.size, dtype=F64)
        return self.routput_vec

    def constant(self, rg, g):  # variable.go:102-111
        if self.variable in rg:
            return False
        if g is None:
            return True
        return self.variable not in g

    def propagate_rgradient(self, upstream, upstream_r, rgrad, grad):  # variable.go:113-123
        v = self.variable
        if grad is not None and v in grad:
            grad[v] += upstream if isinstance(grad[v], DeviceVec) else _host(upstream)
        if v in rgrad:
            rgrad[v] += upstream_r if isinstance(rgrad[v], DeviceVec) else _host(upstream_r)

    def _d(self):
        return self.variable.device()

    def _dr(self):
        """Device R-vector, or None meaning identically zero (the product skips those terms)."""
        if not self._has_r:
            return None
        if self._rdev is None:
            self._rdev = context().upload(self.routput_vec)
        return self._rdev

    def _bwd_r(self, sess, dU, dUR, rgrad, grad):
        v = self.variable
        if grad is not None and v in grad:
            sess.accumulate(grad, v, dU)
        if v in rgrad:
            sess.accumulate(rgrad, v, dUR)


def new_gradient(variables):  # gradient.go:13-19
    return {v: np.zeros(v.size, dtype=F64) for v in variables}


new_rgradient = new_gradient


def gradient_add(g, g1):  # gradient.go:28-30
    for k in g:
        g[k] += g1[k]


def gradient_add_to_vars(g, scale):  # gradient.go:40-52
    for v, gv in g.items():
        if isinstance(gv, DeviceVec):
            v.add_scaled_device(scale, gv)
        else:
            v.vector += scale * gv
            v.mark_dirty()


def gradient_zero(g):  # gradient.go:22-24,100-106
    for gv in g.values():
        if isinstance(gv, DeviceVec):
            check(gv.ctx.lib.af_zero(gv.ctx.h, gv.ptr, gv.n))
        else:
            gv[:] = 0


def gradient_scale(g, f):  # gradient.go:33-36,122-126
    for gv in g.values():
        if isinstance(gv, DeviceVec):
            check(gv.ctx.lib.af_scale(gv.ctx.h, gv.ptr, float(f), gv.n))
        else:
            gv *= f


def gradient_copy(g):  # gradient.go:54-57,128-136
    if isinstance(g, DeviceGradient):
        return DeviceGradient._from_items((k, v.copy()) for k, v in g.items())
    return {k: v.copy() for k, v in g.items()}


class DeviceGradient(dict):
    """A Gradient / RGradient (gradient.go:11,62) whose vectors stay in HBM: backward passes accumulate
    straight into them (no per-call flush to the host), `add_to_vars` is the SGD step on the device, and
    nothing is downloaded until `host()` is asked for.  Keys are Variables, values DeviceVecs."""

    def __init__(self, variables=()):
        super().__init__()
        for v in variables:
            self[v] = context().zeros(v.size)

    @classmethod
    def _from_items(cls, items):
        g = cls()
        for k, v in items:
            g[k] = v
        return g

    def zero(self):
        gradient_zero(self)

    def scale(self, f):
        gradient_scale(self, f)

    def add(self, g1):
        gradient_add(self, g1)

    def add_to_vars(self, scale):
        gradient_add_to_vars(self, scale)

    def copy(self):
        return gradient_copy(self)

    def host(self):
        return {k: _host(v) for k, v in self.items()}


new_device_gradient = new_device_rgradient = DeviceGradient


class _TempGrad:
    """Placeholder a device node inserts as `grad[tmp]` for the reference's temp-variable protocol
    (pool.go:114-136, fold.go:64-80, lin_alg.go:57-68).  Device nodes accumulate into the session's
    HBM accumulator for `tmp`; a foreign host Result that does `grad[tmp] += upstream` materialises a
    host vector here; `_Session.take` returns the sum of both as a DeviceVec."""

    def __init__(self, n):
        self.n, self.host = n, None

    def __len__(self):
        return self.n

    def __iadd__(self, x):
        x = np.asarray(x, dtype=F64)
        self.host = x.copy() if self.host is None else self.host + x
        return self


def _host(x) -> np.ndarray:
    return x.numpy() if isinstance(x, DeviceVec) else np.asarray(x, dtype=F64)


def _dev(x) -> DeviceVec:
    return x if isinstance(x, DeviceVec) else context().upload(x)


# --------------------------------------------------------------------------------------------
# backward session: device accumulators for parameter gradients, flushed with += on exit
# --------------------------------------------------------------------------------------------
class _Session:
    def __init__(self):
        self.acc: dict[tuple[int, int], tuple[dict, Variable, DeviceVec]] = {}

    def accumulator(self, gmap, v) -> DeviceVec:
        cur = gmap.get(v)
        if isinstance(cur, DeviceVec):  # DeviceGradient: the map's own vector is the accumulator
            return cur
        key = (id(gmap), id(v))
        ent = self.acc.get(key)
        if ent is None:
            ent = (gmap, v, context().zeros(v.size))
            self.acc[key] = ent
        return ent[2]

    def take(self, gmap, v) -> DeviceVec:
        """Remove temp variable `v` from `gmap` and return everything accumulated for it, on the device."""
        ent = self.acc.pop((id(gmap), id(v)), None)
        cur = gmap.pop(v)
        d = ent[2] if ent is not None else None
        host = cur.host if isinstance(cur, _TempGrad) else cur
        if isinstance(host, DeviceVec):
            d = host if d is None else d.__iadd__(host)
        elif host is not None:
            up = context().upload(host)
            d = up if d is None else d.__iadd__(up)
        return d if d is not None else context().zeros(v.size)

    def accumulate(self, gmap, v, dU: DeviceVec | None):
        if dU is None:
            return
        a = self.accumulator(gmap, v)
        ctx = context()
        check(ctx.lib.af_axpy(ctx.h, 1.0, dU.ptr, a.ptr, a.n))

    def flush(self):
        for gmap, v, a in self.acc.values():
            if v in gmap:
                cur = gmap[v]
                cur += a if isinstance(cur, DeviceVec) else a.numpy()
                gmap[v] = cur
        self.acc.clear()


class _DeviceResult:
    """Base of every Result/RResult produced by this module."""

    def __init__(self, d: DeviceVec, dr: DeviceVec | None = None, is_r: bool = False):
        self.d, self.dr, self.is_r = d, dr, is_r
        self._host_out = None
        self._host_rout = None

    def output(self):
        if self._host_out is None:
            self._host_out = self.d.numpy()
        return self._host_out

    def routput(self):
        if self._host_rout is None:
            self._host_rout = self.dr.numpy() if self.dr is not None else np.zeros(self.d.n, dtype=F64)
        return self._host_rout

    def _d(self):
        return self.d

    def _dr(self):
        return self.dr

    # host entry points --------------------------------------------------------------------
    def propagate_gradient(self, upstream, grad):
        if len(upstream) != self.d.n:
            raise ShapeError(_lib.AF_ESHAPE, "invalid upstream size")
        sess = _Session()
        self._bwd(sess, _dev(upstream), grad)
        sess.flush()

    def propagate_rgradient(self, upstream, upstream_r, rgrad, grad):
        if len(upstream) != self.d.n or len(upstream_r) != self.d.n:
            raise ShapeError(_lib.AF_ESHAPE, "invalid upstream size")
        sess = _Session()
        self._bwd_r(sess, _dev(upstream), _dev(upstream_r), rgrad, grad)
        sess.flush()


def _in_d(r) -> DeviceVec:
    return r._d() if hasattr(r, "_d") else context().upload(r.output())


def _in_dr(r) -> DeviceVec | None:
    if hasattr(r, "_dr"):
        return r._dr()
    return context().upload(r.routput())


def _down(r, sess, dU, grad):
    """Continue the non-R recursion into input `r` (device node, Variable or foreign host Result)."""
    if hasattr(r, "_bwd"):
        r._bwd(sess, dU, grad)
    else:
        sess.flush()
        r.propagate_gradient(dU.numpy(), grad)


def _down_r(r, sess, dU, dUR, rgrad, grad):
    if hasattr(r, "_bwd_r"):
        r._bwd_r(sess, dU, dUR, rgrad, grad)
    else:
        sess.flush()
        r.propagate_rgradient(dU.numpy(), dUR.numpy(), rgrad, grad)


# --------------------------------------------------------------------------------------------
# LinTran (lin_tran.go)
# --------------------------------------------------------------------------------------------
class LinTran:
    def __init__(self, data: Variable, rows: int, cols: int):
        self.data, self.rows, self.cols = data, int(rows), int(cols)

    def _check(self, inp, n):
        got = len(inp.output()) if not hasattr(inp, "_d") else inp._d().n
        if self.cols * n != got:  # lin_tran.go:26-27,52-54: same message as the reference's panic
            raise ShapeError(_lib.AF_ESHAPE, "input length should be %d but got %d" % (self.cols * n, got))

    def apply(self, inp):
        self._check(inp, 1)
        return self.batch(inp, 1)

    def apply_r(self, rv, inp):
        self._check(inp, 1)
        return self.batch_r(rv, inp, 1)

    def batch(self, inp, n):
        self._check(inp, n)
        ctx = context()
        x = _in_d(inp)
        y = ctx.empty(n * self.rows)
        check(ctx.lib.af_lintran_fwd(ctx.h, self.data.device().ptr, x.ptr, y.ptr, self.rows, self.cols, n))
        return _LinTranResult(self, inp, x, n, y)

    def batch_r(self, rv, inp, n):
        self._check(inp, n)
        ctx = context()
        rdata = RVariable(self.data, rv)
        x, xr = _in_d(inp), _in_dr(inp)
        y, ry = ctx.empty(n * self.rows), ctx.empty(n * self.rows)
        check(ctx.lib.af_lintran_fwd_r(ctx.h, rdata._d().ptr, _p(rdata._dr()), x.ptr, _p(xr), y.ptr, ry.ptr,
                                       self.rows, self.cols, n))
        return _LinTranRResult(self, inp, rdata, x, xr, n, y, ry)


class _LinTranResult(_DeviceResult):
    def __init__(self, m, inp, x, n, y):
        super().__init__(y)
        self.matrix, self.input, self.x, self.n = m, inp, x, n

    def constant(self, g):
        return self.matrix.data.constant(g) and self.input.constant(g)

    def _bwd(self, sess, dU, grad):  # lin_tran.go:373-382 -- both gradients from ONE call (af_lintran_backward)
        ctx, m = context(), self.matrix
        acc = sess.accumulator(grad, m.data) if not m.data.constant(grad) else None
        need_d = not self.input.constant(grad)
        if acc is None and not need_d:
            return
        into = sess.accumulator(grad, self.input) if (need_d and isinstance(self.input, Variable) and self.input in grad) else None
        d = into if into is not None else (ctx.empty(self.n * m.cols) if need_d else None)
        check(ctx.lib.af_lintran_backward(ctx.h, dU.ptr, m.data.device().ptr if need_d else None, self.x.ptr, _p(acc), _p(d),
                                          m.rows, m.cols, self.n, 1 if into is not None else 0))
        if need_d and into is None:
            _down(self.input, sess, d, grad)


class _LinTranRResult(_DeviceResult):
    def __init__(self, m, inp, rdata, x, xr, n, y, ry):
        super().__init__(y, ry, True)
        self.matrix, self.input, self.rdata, self.x, self.xr, self.n = m, inp, rdata, x, xr, n

    def constant(self, rg, g):
        return self.input.constant(rg, g) and self.rdata.constant(rg, g)

    def _bwd_r(self, sess, dU, dUR, rgrad, grad):  # lin_tran.go:408-425 -- all four gradients from ONE call
        ctx, m = context(), self.matrix
        gacc = sess.accumulator(grad, m.data) if (grad is not None and not m.data.constant(grad)) else None
        racc = sess.accumulator(rgrad, m.data) if m.data in rgrad else None
        need_d = not self.input.constant(rgrad, grad)
        if gacc is None and racc is None and not need_d:
            return
        # the input is an RVariable present in both maps: its PropagateRGradient (variable.go:113-123) is `+=` into the
        # accumulators, which the call does itself
        v = self.input.variable if isinstance(self.input, RVariable) else None
        direct = need_d and v is not None and grad is not None and v in grad and v in rgrad
        if direct:
            d, dr = sess.accumulator(grad, v), sess.accumulator(rgrad, v)
        elif need_d:
            d, dr = ctx.empty(self.n * m.cols), ctx.empty(self.n * m.cols)
        else:
            d = dr = None
        check(ctx.lib.af_lintran_backward_r(ctx.h, dU.ptr, dUR.ptr, self.rdata._d().ptr if need_d else None,
                                            _p(self.rdata._dr()) if need_d else None, self.x.ptr, _p(self.xr), _p(gacc), _p(racc),
                                            _p(d), _p(dr), m.rows, m.cols, self.n, 1 if direct else 0))
        if need_d and not direct:
            _down_r(self.input, sess, d, dr, rgrad, grad)


# --------------------------------------------------------------------------------------------
# LinAdd (lin_add.go) -- also a Batcher: the packed form is Add(x, Repeat(b, n))
# --------------------------------------------------------------------------------------------
class LinAdd:
    def __init__(self, var: Variable):
        self.var = var

    def _n(self, inp, n):
        got = _in_len(inp)
        if got != n * self.var.size:
            raise ShapeError(_lib.AF_ESHAPE, "input length should be %d but got %d" % (n * self.var.size, got))

    def apply(self, inp):
        return self.batch(inp, 1)

    def apply_r(self, rv, inp):
        return self.batch_r(rv, inp, 1)

    def batch(self, inp, n):
        self._n(inp, n)
        ctx = context()
        x = _in_d(inp)
        y = ctx.empty(x.n)
        check(ctx.lib.af_bias_fwd(ctx.h, x.ptr, self.var.device().ptr, y.ptr, self.var.size, n))
        return _LinAddResult(self.var, inp, n, y)

    def batch_r(self, rv, inp, n):
        self._n(inp, n)
        ctx = context()
        rvar = RVariable(self.var, rv)
        x, xr = _in_d(inp), _in_dr(inp)
        y, ry = ctx.empty(x.n), ctx.empty(x.n)
        check(ctx.lib.af_bias_fwd_r(ctx.h, x.ptr, _p(xr), rvar._d().ptr, _p(rvar._dr()), y.ptr, ry.ptr,
                                    self.var.size, n))
        return _LinAddRResult(rvar, inp, n, y, ry)


def _in_len(inp) -> int:
    return inp._d().n if hasattr(inp, "_d") else len(inp.output())


class _LinAddResult(_DeviceResult):
    def __init__(self, var, inp, n, y):
        super().__init__(y)
        self.sum_var, self.input, self.n = var, inp, n

    def constant(self, g):
        return self.input.constant(g) and self.sum_var.constant(g)

    def _bwd(self, sess, dU, grad):  # lin_add.go:66-73
        ctx = context()
        if self.sum_var in grad:
            acc = sess.accumulator(grad, self.sum_var)
            check(ctx.lib.af_bias_grad(ctx.h, dU.ptr, acc.ptr, acc.n, self.n))
        if not self.input.constant(grad):
            _down(self.input, sess, dU, grad)


class _LinAddRResult(_DeviceResult):
    def __init__(self, rvar, inp, n, y, ry):
        super().__init__(y, ry, True)
        self.sum_var, self.input, self.n = rvar, inp, n

    def constant(self, rg, g):
        return self.sum_var.constant(rg, g) and self.input.constant(rg, g)

    def _bwd_r(self, sess, dU, dUR, rgrad, grad):  # lin_add.go:94-109
        ctx, v = context(), self.sum_var.variable
        gacc = sess.accumulator(grad, v) if (grad is not None and v in grad) else None
        racc = sess.accumulator(rgrad, v) if v in rgrad else None
        if gacc is not None or racc is not None:
            check(ctx.lib.af_bias_grad_r(ctx.h, dU.ptr, dUR.ptr, _p(gacc), _p(racc), v.size, self.n))
        if not self.input.constant(rgrad, grad):
            _down_r(self.input, sess, dU, dUR, rgrad, grad)


# --------------------------------------------------------------------------------------------
# elementwise RFuncs (math_funcs.go); each is also an RBatcher
# --------------------------------------------------------------------------------------------
class _Elementwise:
    def batch(self, inp, n):
        if _in_len(inp) % n != 0:  # batcher.go:41
            raise ShapeError(_lib.AF_ESHAPE, "input count does not divide input length")
        return self.apply(inp)

    def batch_r(self, rv, inp, n):
        if _in_len(inp) % n != 0:
            raise ShapeError(_lib.AF_ESHAPE, "input count does not divide input length")
        return self.apply_r(rv, inp)


class _UnaryResult(_DeviceResult):
    def __init__(self, inp, d, dr=None, is_r=False):
        super().__init__(d, dr, is_r)
        self.input = inp

    def constant(self, *maps):
        return self.input.constant(*maps)


class Sigmoid(_Elementwise):  # math_funcs.go:286-374
    def apply(self, inp):
        ctx = context()
        x = _in_d(inp)
        s = ctx.empty(x.n)
        check(ctx.lib.af_sigmoid_fwd(ctx.h, x.ptr, s.ptr, x.n))
        return _SigmoidResult(inp, s)

    def apply_r(self, rv, inp):
        ctx = context()
        x, xr = _in_d(inp), _in_dr(inp)
        s, rs = ctx.empty(x.n), ctx.empty(x.n)
        check(ctx.lib.af_sigmoid_fwd_r(ctx.h, x.ptr, _p(xr), s.ptr, rs.ptr, x.n))
        return _SigmoidResult(inp, s, rs, True)


class _SigmoidResult(_UnaryResult):
    def _bwd(self, sess, dU, grad):
        if not self.input.constant(grad):
            ctx = context()
            check(ctx.lib.af_sigmoid_bwd(ctx.h, self.d.ptr, dU.ptr, dU.n))
            _down(self.input, sess, dU, grad)

    def _bwd_r(self, sess, dU, dUR, rgrad, grad):
        if self.input.constant(rgrad, grad):
            return
        ctx = context()
        check(ctx.lib.af_sigmoid_bwd_r(ctx.h, self.d.ptr, self.dr.ptr, dU.ptr, dUR.ptr, dU.n))
        _down_r(self.input, sess, dU, dUR, rgrad, grad)


class Exp(_Elementwise):  # math_funcs.go:12-97
    def apply(self, inp):
        ctx = context()
        x = _in_d(inp)
        e = ctx.empty(x.n)
        check(ctx.lib.af_exp_fwd(ctx.h, x.ptr, e.ptr, x.n))
        return _ExpResult(inp, e)

    def apply_r(self, rv, inp):
        ctx = context()
        x, xr = _in_d(inp), _in_dr(inp)
        e, re_ = ctx.empty(x.n), ctx.empty(x.n)
        check(ctx.lib.af_exp_fwd_r(ctx.h, x.ptr, _p(xr), e.ptr, re_.ptr, x.n))
        return _ExpResult(inp, e, re_, True)


class _ExpResult(_UnaryResult):
    def _bwd(self, sess, dU, grad):
        if not self.input.constant(grad):
            ctx = context()
            check(ctx.lib.af_exp_bwd(ctx.h, self.d.ptr, dU.ptr, dU.n))
            _down(self.input, sess, dU, grad)

    def _bwd_r(self, sess, dU, dUR, rgrad, grad):
        if not self.input.constant(rgrad, grad):
            ctx = context()
            check(ctx.lib.af_exp_bwd_r(ctx.h, self.d.ptr, self.dr.ptr, dU.ptr, dUR.ptr, dU.n))
            _down_r(self.input, sess, dU, dUR, rgrad, grad)


class SquaredNorm(_Elementwise):  # math_funcs.go:191-199
    def apply(self, inp):
        return sum_all(square(inp))

    def apply_r(self, rv, inp):
        return sum_all_r(square_r(inp))


# --------------------------------------------------------------------------------------------
# arithmetic.go hot subset
# --------------------------------------------------------------------------------------------
class _BinaryResult(_DeviceResult):
    def __init__(self, r1, r2, d, dr=None, is_r=False):
        super().__init__(d, dr, is_r)
        self.r1, self.r2 = r1, r2

    def constant(self, *maps):
        return self.r1.constant(*maps) and self.r2.constant(*maps)


def _same_len(r1, r2):
    if _in_len(r1) != _in_len(r2):
        raise ShapeError(_lib.AF_ESHAPE, "vector sizes do not match")  # arithmetic.go:265


def _binary(fn_name, r1, r2):
    ctx = context()
    a, b = _in_d(r1), _in_d(r2)
    out = ctx.empty(a.n)
    check(getattr(ctx.lib, fn_name)(ctx.h, a.ptr, b.ptr, out.ptr, a.n))
    return out


def _zeros_if_none(v, n):
    return v if v is not None else context().zeros(n)


class _Sum(_BinaryResult):
    def _bwd(self, sess, dU, grad):  # arithmetic.go:34-49
        r2c = self.r2.constant(grad)
        if not self.r1.constant(grad):
            _down(self.r1, sess, dU if r2c else dU.copy(), grad)
        if not r2c:
            _down(self.r2, sess, dU, grad)

    def _bwd_r(self, sess, dU, dUR, rgrad, grad):  # arithmetic.go:79-96
        r2c = self.r2.constant(rgrad, grad)
        if not self.r1.constant(rgrad, grad):
            if r2c:
                _down_r(self.r1, sess, dU, dUR, rgrad, grad)
            else:
                _down_r(self.r1, sess, dU.copy(), dUR.copy(), rgrad, grad)
        if not r2c:
            _down_r(self.r2, sess, dU, dUR, rgrad, grad)


def add(r1, r2):
    _same_len(r1, r2)
    return _Sum(r1, r2, _binary("af_add", r1, r2))


def add_r(r1, r2):
    _same_len(r1, r2)
    ctx = context()
    n = _in_len(r1)
    a, b = _zeros_if_none(_in_dr(r1), n), _zeros_if_none(_in_dr(r2), n)
    outr = ctx.empty(n)
    check(ctx.lib.af_add(ctx.h, a.ptr, b.ptr, outr.ptr, n))
    return _Sum(r1, r2, _binary("af_add", r1, r2), outr, True)


class _Diff(_BinaryResult):
    def _bwd(self, sess, dU, grad):  # arithmetic.go:128-135
        ctx = context()
        if not self.r1.constant(grad):
            _down(self.r1, sess, dU.copy(), grad)
        if not self.r2.constant(grad):
            check(ctx.lib.af_scale(ctx.h, dU.ptr, -1.0, dU.n))
            _down(self.r2, sess, dU, grad)

    def _bwd_r(self, sess, dU, dUR, rgrad, grad):  # arithmetic.go:175-182
        ctx = context()
        if not self.r1.constant(rgrad, grad):
            _down_r(self.r1, sess, dU.copy(), dUR.copy(), rgrad, grad)
        if not self.r2.constant(rgrad, grad):
            check(ctx.lib.af_scale(ctx.h, dU.ptr, -1.0, dU.n))
            check(ctx.lib.af_scale(ctx.h, dUR.ptr, -1.0, dUR.n))
            _down_r(self.r2, sess, dU, dUR, rgrad, grad)


def sub(a, b):
    _same_len(a, b)
    return _Diff(a, b, _binary("af_sub", a, b))


def sub_r(a, b):
    _same_len(a, b)
    ctx = context()
    n = _in_len(a)
    ar, br = _zeros_if_none(_in_dr(a), n), _zeros_if_none(_in_dr(b), n)
    outr = ctx.empty(n)
    check(ctx.lib.af_sub(ctx.h, ar.ptr, br.ptr, outr.ptr, n))
    return _Diff(a, b, _binary("af_sub", a, b), outr, True)


class _Product(_BinaryResult):
    def _bwd(self, sess, dU, grad):  # arithmetic.go:288-306
        ctx = context()
        for me, other in ((self.r1, self.r2), (self.r2, self.r1)):
            if not me.constant(grad):
                d = ctx.empty(dU.n)
                check(ctx.lib.af_mul_bwd(ctx.h, dU.ptr, None, _in_d(other).ptr, None, d.ptr, None, dU.n))
                _down(me, sess, d, grad)

    def _bwd_r(self, sess, dU, dUR, rgrad, grad):  # arithmetic.go:349-376
        ctx = context()
        for me, other in ((self.r1, self.r2), (self.r2, self.r1)):
            if not me.constant(rgrad, grad):
                d, dr = ctx.empty(dU.n), ctx.empty(dU.n)
                check(ctx.lib.af_mul_bwd(ctx.h, dU.ptr, dUR.ptr, _in_d(other).ptr, _p(_in_dr(other)), d.ptr, dr.ptr, dU.n))
                _down_r(me, sess, d, dr, rgrad, grad)


def mul(r1, r2):
    _same_len(r1, r2)
    return _Product(r1, r2, _binary("af_mul", r1, r2))


def mul_r(r1, r2):
    _same_len(r1, r2)
    ctx = context()
    a, b = _in_d(r1), _in_d(r2)
    outr = ctx.empty(a.n)
    check(ctx.lib.af_mul_r(ctx.h, a.ptr, _p(_in_dr(r1)), b.ptr, _p(_in_dr(r2)), outr.ptr, a.n))
    return _Product(r1, r2, _binary("af_mul", r1, r2), outr, True)


class _Scaled(_UnaryResult):
    scaler = 1.0

    def _bwd(self, sess, dU, grad):  # arithmetic.go:444-456
        if not self.input.constant(grad):
            ctx = context()
            check(ctx.lib.af_scale(ctx.h, dU.ptr, self.scaler, dU.n))
            _down(self.input, sess, dU, grad)

    def _bwd_r(self, sess, dU, dUR, rgrad, grad):  # arithmetic.go:476-493
        if not self.input.constant(rgrad, grad):
            ctx = context()
            check(ctx.lib.af_scale(ctx.h, dU.ptr, self.scaler, dU.n))
            check(ctx.lib.af_scale(ctx.h, dUR.ptr, self.scaler, dUR.n))
            _down_r(self.input, sess, dU, dUR, rgrad, grad)


def scale(r, f):
    ctx = context()
    out = _in_d(r).copy()
    check(ctx.lib.af_scale(ctx.h, out.ptr, float(f), out.n))
    res = _Scaled(r, out)
    res.scaler = float(f)
    return res


def scale_r(r, f):
    ctx = context()
    out = _in_d(r).copy()
    check(ctx.lib.af_scale(ctx.h, out.ptr, float(f), out.n))
    outr = _zeros_if_none(_in_dr(r), out.n).copy()
    check(ctx.lib.af_scale(ctx.h, outr.ptr, float(f), outr.n))
    res = _Scaled(r, out, outr, True)
    res.scaler = float(f)
    return res


class _Square(_UnaryResult):
    def _bwd(self, sess, dU, grad):  # arithmetic.go:716-723
        if not self.input.constant(grad):
            ctx = context()
            check(ctx.lib.af_square_bwd_r(ctx.h, _in_d(self.input).ptr, None, dU.ptr, None, dU.n))
            _down(self.input, sess, dU, grad)

    def _bwd_r(self, sess, dU, dUR, rgrad, grad):  # arithmetic.go:760-770
        if not self.input.constant(rgrad, grad):
            ctx = context()
            check(ctx.lib.af_square_bwd_r(ctx.h, _in_d(self.input).ptr, _p(_in_dr(self.input)), dU.ptr, dUR.ptr, dU.n))
            _down_r(self.input, sess, dU, dUR, rgrad, grad)


def square(r):
    ctx = context()
    x = _in_d(r)
    out = ctx.empty(x.n)
    check(ctx.lib.af_square_fwd_r(ctx.h, x.ptr, None, out.ptr, None, x.n))
    return _Square(r, out)


def square_r(r):
    ctx = context()
    x = _in_d(r)
    out, outr = ctx.empty(x.n), ctx.empty(x.n)
    check(ctx.lib.af_square_fwd_r(ctx.h, x.ptr, _p(_in_dr(r)), out.ptr, outr.ptr, x.n))
    return _Square(r, out, outr, True)


class _SumAll(_UnaryResult):
    def _fill(self, src: DeviceVec, n: int) -> DeviceVec:
        ctx = context()
        out = ctx.empty(n)
        check(ctx.lib.af_fill_dev(ctx.h, out.ptr, src.ptr, n))  # the upstream scalar never leaves the device
        return out

    def _bwd(self, sess, dU, grad):  # arithmetic.go:987-995
        if not self.input.constant(grad):
            _down(self.input, sess, self._fill(dU, _in_len(self.input)), grad)

    def _bwd_r(self, sess, dU, dUR, rgrad, grad):  # arithmetic.go:1035-1047
        if not self.input.constant(rgrad, grad):
            n = _in_len(self.input)
            _down_r(self.input, sess, self._fill(dU, n), self._fill(dUR, n), rgrad, grad)


def sum_all(r):
    ctx = context()
    x = _in_d(r)
    out = ctx.empty(1)
    check(ctx.lib.af_sum_all(ctx.h, x.ptr, None, out.ptr, None, x.n))
    return _SumAll(r, out)


def sum_all_r(r):
    ctx = context()
    x = _in_d(r)
    xr = _zeros_if_none(_in_dr(r), x.n)
    out, outr = ctx.empty(1), ctx.empty(1)
    check(ctx.lib.af_sum_all(ctx.h, x.ptr, xr.ptr, out.ptr, outr.ptr, x.n))
    return _SumAll(r, out, outr, True)


# --------------------------------------------------------------------------------------------
# composition (func.go, batcher.go)
# --------------------------------------------------------------------------------------------
class ComposedRFunc(list):
    def apply(self, inp):
        for f in self:
            inp = f.apply(inp)
        return inp

    def apply_r(self, rv, inp):
        for f in self:
            inp = f.apply_r(rv, inp)
        return inp


ComposedFunc = ComposedRFunc


class ComposedRBatcher(list):
    def batch(self, inp, n):
        for b in self:
            inp = b.batch(inp, n)
        return inp

    def batch_r(self, rv, inp, n):
        for b in self:
            inp = b.batch_r(rv, inp, n)
        return inp


ComposedBatcher = ComposedRBatcher


# --------------------------------------------------------------------------------------------
# fused dense layer: LinTran -> LinAdd -> Sigmoid as one RBatcher (one kernel per direction)
# --------------------------------------------------------------------------------------------
class DenseSigmoid:
    """Drop-in for the triple {&LinTran{W}, &LinAdd{b}, Sigmoid{}} of bench/mlp.go:75-83."""

    def __init__(self, weights: Variable, biases: Variable, rows: int, cols: int, activation: bool = True):
        self.w, self.b, self.rows, self.cols, self.activation = weights, biases, int(rows), int(cols), activation
        if weights.size != rows * cols or biases.size != rows:
            raise ShapeError(_lib.AF_ESHAPE, "parameter sizes do not match rows x cols")

    def _check(self, inp, n):
        got = _in_len(inp)
        if self.cols * n != got:
            raise ShapeError(_lib.AF_ESHAPE, "input length should be %d but got %d" % (self.cols * n, got))

    def apply(self, inp):
        return self.batch(inp, 1)

    def apply_r(self, rv, inp):
        return self.batch_r(rv, inp, 1)

    def batch(self, inp, n):
        self._check(inp, n)
        ctx = context()
        x = _in_d(inp)
        s = ctx.empty(n * self.rows)
        check(ctx.lib.af_dense_fwd(ctx.h, self.w.device().ptr, None, self.b.device().ptr, None, x.ptr, None, s.ptr,
                                   None, self.rows, self.cols, n, int(self.activation)))
        return _DenseResult(self, inp, None, None, x, None, n, s, None)

    def batch_r(self, rv, inp, n):
        self._check(inp, n)
        ctx = context()
        rw, rb = RVariable(self.w, rv), RVariable(self.b, rv)
        x, xr = _in_d(inp), _in_dr(inp)
        s, rs = ctx.empty(n * self.rows), ctx.empty(n * self.rows)
        check(ctx.lib.af_dense_fwd(ctx.h, rw._d().ptr, _p(rw._dr()), rb._d().ptr, _p(rb._dr()), x.ptr, _p(xr), s.ptr,
                                   rs.ptr, self.rows, self.cols, n, int(self.activation)))
        return _DenseResult(self, inp, rw, rb, x, xr, n, s, rs)


class _DenseResult(_DeviceResult):
    def __init__(self, layer, inp, rw, rb, x, xr, n, s, rs):
        super().__init__(s, rs, rs is not None)
        self.layer, self.input, self.rw, self.rb, self.x, self.xr, self.n = layer, inp, rw, rb, x, xr, n

    def constant(self, *maps):
        if self.is_r:
            return self.input.constant(*maps) and self.rw.constant(*maps) and self.rb.constant(*maps)
        return self.input.constant(*maps) and self.layer.w.constant(*maps) and self.layer.b.constant(*maps)

    def _bwd(self, sess, dU, grad, skip_act=False):
        ctx, L = context(), self.layer
        if self.constant(grad):
            return
        if L.activation and not skip_act:
            check(ctx.lib.af_sigmoid_bwd(ctx.h, self.d.ptr, dU.ptr, dU.n))
        if L.b in grad:
            check(ctx.lib.af_bias_grad(ctx.h, dU.ptr, sess.accumulator(grad, L.b).ptr, L.rows, self.n))
        if L.w in grad:
            check(ctx.lib.af_lintran_wgrad(ctx.h, dU.ptr, self.x.ptr, sess.accumulator(grad, L.w).ptr, L.rows, L.cols, self.n))
        if not self.input.constant(grad):
            self._input_grad(sess, dU, None, None, grad)

    def _bwd_r(self, sess, dU, dUR, rgrad, grad, skip_act=False):
        ctx, L = context(), self.layer
        if self.constant(rgrad, grad):
            return
        if L.activation and not skip_act:
            check(ctx.lib.af_sigmoid_bwd_r(ctx.h, self.d.ptr, self.dr.ptr, dU.ptr, dUR.ptr, dU.n))
        gb = sess.accumulator(grad, L.b) if (grad is not None and L.b in grad) else None
        rgb = sess.accumulator(rgrad, L.b) if L.b in rgrad else None
        if gb is not None or rgb is not None:
            check(ctx.lib.af_bias_grad_r(ctx.h, dU.ptr, dUR.ptr, _p(gb), _p(rgb), L.rows, self.n))
        gw = sess.accumulator(grad, L.w) if (grad is not None and L.w in grad) else None
        rgw = sess.accumulator(rgrad, L.w) if L.w in rgrad else None
        if gw is not None or rgw is not None:
            check(ctx.lib.af_lintran_wgrad_r(ctx.h, dU.ptr, dUR.ptr, self.x.ptr, _p(self.xr), _p(gw), _p(rgw),
                                             L.rows, L.cols, self.n))
        if not self.input.constant(rgrad, grad):
            self._input_grad(sess, dU, dUR, rgrad, grad)

    def _input_grad(self, sess, dU, dUR, rgrad, grad):
        """G5/G6 with the previous DenseSigmoid's sigmoid backward fused into the epilogue."""
        ctx, L = context(), self.layer
        prev = self.input
        fuse = isinstance(prev, _DenseResult) and prev.layer.activation
        r = dUR is not None
        d = ctx.empty(self.n * L.cols)
        dr = ctx.empty(self.n * L.cols) if r else None
        w = self.rw._d() if r else L.w.device()
        check(ctx.lib.af_dense_igrad(ctx.h, dU.ptr, _p(dUR), w.ptr, _p(self.rw._dr()) if r else None,
                                     prev.d.ptr if fuse else None, _p(prev.dr) if (fuse and r) else None,
                                     d.ptr, _p(dr), L.rows, L.cols, self.n))
        if fuse:  # the layer above already applied prev's sigmoid backward in its epilogue
            if r:
                prev._bwd_r(sess, d, dr, rgrad, grad, skip_act=True)
            else:
                prev._bwd(sess, d, grad, skip_act=True)
        elif r:
            _down_r(prev, sess, d, dr, rgrad, grad)
        else:
            _down(prev, sess, d, grad)


def fuse_dense(layers):
    """Rewrite {LinTran, LinAdd, Sigmoid} runs of a composed function into DenseSigmoid layers."""
    out, i = ComposedRBatcher(), 0
    layers = list(layers)
    while i < len(layers):
        a = layers[i]
        if (isinstance(a, LinTran) and i + 1 < len(layers) and isinstance(layers[i + 1], LinAdd)
                and layers[i + 1].var.size == a.rows):
            act = i + 2 < len(layers) and isinstance(layers[i + 2], Sigmoid)
            out.append(DenseSigmoid(a.data, layers[i + 1].var, a.rows, a.cols, activation=act))
            i += 3 if act else 2
        else:
            out.append(a)
            i += 1
    return out


# --------------------------------------------------------------------------------------------
# whole-network plan (af_mlp_*): bench/mlp.go batched, everything resident
# --------------------------------------------------------------------------------------------
class MLPPlan:
    def __init__(self, layer_sizes, batch: int, ctx: Context | None = None):
        self.ctx = ctx or context()
        self.sizes = [int(s) for s in layer_sizes]
        self.L = len(self.sizes) - 1
        self.batch = int(batch)
        arr = (C.c_int64 * len(self.sizes))(*self.sizes)
        h = C.c_void_p()
        check(self.ctx.lib.af_mlp_create(self.ctx.h, arr, self.L, self.batch, C.byref(h)))
        self.h = h

    def param_len(self, idx):
        l = idx // 2
        return self.sizes[l + 1] if idx & 1 else self.sizes[l] * self.sizes[l + 1]

    def set_param(self, idx, value):
        a = np.ascontiguousarray(value, dtype=F64).reshape(-1)
        check(self.ctx.lib.af_mlp_set_param(self.h, idx, a.ctypes.data, a.size))

    def set_rvector(self, idx, value):
        if value is None:
            check(self.ctx.lib.af_mlp_set_rvector(self.h, idx, None, 0))
            return
        a = np.ascontiguousarray(value, dtype=F64).reshape(-1)
        check(self.ctx.lib.af_mlp_set_rvector(self.h, idx, a.ctypes.data, a.size))

    def _vec(self, ptr, n):
        return DeviceVec(self.ctx, n, ptr, owner=self)

    def grad(self, idx):
        return self._vec(self.ctx.lib.af_mlp_grad_ptr(self.h, idx), self.param_len(idx))

    def rgrad(self, idx):
        return self._vec(self.ctx.lib.af_mlp_rgrad_ptr(self.h, idx), self.param_len(idx))

    def input(self):
        return self._vec(self.ctx.lib.af_mlp_input_ptr(self.h), self.batch * self.sizes[0])

    def output(self):
        return self._vec(self.ctx.lib.af_mlp_output_ptr(self.h), self.batch * self.sizes[-1])

    def routput(self):
        return self._vec(self.ctx.lib.af_mlp_routput_ptr(self.h), self.batch * self.sizes[-1])

    def upstream(self):
        return self._vec(self.ctx.lib.af_mlp_upstream_ptr(self.h), self.batch * self.sizes[-1])

    def upstream_r(self):
        return self._vec(self.ctx.lib.af_mlp_upstream_r_ptr(self.h), self.batch * self.sizes[-1])

    def zero_grads(self):
        check(self.ctx.lib.af_mlp_zero_grads(self.h))

    def add_to_vars(self, scale: float):
        """Gradient.AddToVars (gradient.go:40-52) on the device: param += scale * grad."""
        check(self.ctx.lib.af_mlp_add_to_vars(self.h, float(scale)))

    def step_fresh(self, with_r=True):
        """One bench iteration in one call: fresh accumulators, outGrad = 1 / outRGrad = 0, forward + backward."""
        check(self.ctx.lib.af_mlp_step_fresh(self.h, int(with_r)))

    # ---- data parallel (gradient.go:28-30 across the context's communicator) ----
    def set_dp(self, mode: int):
        check(self.ctx.lib.af_mlp_set_dp(self.h, int(mode)))

    def step_dp(self, with_r=True):
        """af_mlp_step(with_r, backprop) from accumulators the caller zeroed, then the sum over ranks."""
        check(self.ctx.lib.af_mlp_step_dp(self.h, int(with_r)))

    def step_fresh_dp(self, with_r=True):
        check(self.ctx.lib.af_mlp_step_fresh_dp(self.h, int(with_r)))

    def dp_join(self):
        check(self.ctx.lib.af_mlp_dp_join(self.h))

    def set_host_shard(self, enable=True):
        check(self.ctx.lib.af_mlp_set_host_shard(self.h, int(bool(enable))))

    def grad_slab(self):
        """The contiguous {grad | rgrad} block of the bound slot (layer blocks gW, gb, rgW, rgb adjacent)."""
        p = C.c_void_p()
        n = self.ctx.lib.af_mlp_grad_slab(self.h, C.byref(p))
        return self._vec(p.value, int(n))

    def get_param(self, idx) -> np.ndarray:
        out = np.empty(self.param_len(idx), dtype=F64)
        check(self.ctx.lib.af_mlp_get_param(self.h, idx, out.ctypes.data, out.size))
        return out

    def step(self, with_r=True, backprop=True):
        check(self.ctx.lib.af_mlp_step(self.h, int(with_r), int(backprop)))

    def step_host(self, x, with_r=True, upstream=None, upstream_r=None, out=None, rout=None, grads=None, rgrads=None):
        """The reference-facing call with HOST buffers (numpy arrays; pinned or pageable)."""
        n_out = self.batch * self.sizes[-1]
        out = np.empty(n_out, dtype=F64) if out is None else out
        rout = np.empty(n_out, dtype=F64) if (rout is None and with_r) else rout
        P = 2 * self.L
        grads = [np.empty(self.param_len(i), dtype=F64) for i in range(P)] if grads is None else grads
        rgrads = [np.empty(self.param_len(i), dtype=F64) for i in range(P)] if (rgrads is None and with_r) else rgrads
        HP = C.c_void_p * P
        g_arr = HP(*[g.ctypes.data for g in grads])
        rg_arr = HP(*[g.ctypes.data for g in rgrads]) if rgrads is not None else None
        ptr = lambda a: None if a is None else a.ctypes.data
        check(self.ctx.lib.af_mlp_step_host(self.h, int(with_r), ptr(x), ptr(upstream), ptr(upstream_r), ptr(out),
                                            ptr(rout), g_arr, rg_arr))
        return out, rout, grads, rgrads

    def submit_host(self, x, with_r=True, upstream=None, upstream_r=None, out=None, rout=None, grads=None, rgrads=None):
        """Pipelined step_host: queue the step and return; results are valid after the matching wait().
        All arrays must stay alive (and should be pinned) until then."""
        P = 2 * self.L
        HP = C.c_void_p * P
        g_arr = HP(*[g.ctypes.data for g in grads]) if grads is not None else None
        rg_arr = HP(*[g.ctypes.data for g in rgrads]) if rgrads is not None else None
        ptr = lambda a: None if a is None else a.ctypes.data
        check(self.ctx.lib.af_mlp_submit_host(self.h, int(with_r), ptr(x), ptr(upstream), ptr(upstream_r), ptr(out),
                                              ptr(rout), g_arr, rg_arr))

    def wait(self):
        check(self.ctx.lib.af_mlp_wait(self.h))

    def close(self):
        if self.h is not None:
            self.ctx.lib.af_mlp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
