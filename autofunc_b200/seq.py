"""seqfunc.MapBatcher / MapRBatcher (seqfunc/map_batcher.go:14-140) over device Batchers.

Sequence lists are the reference's `[][]linalg.Vector`: a list of lanes, each a list of per-timestep
vectors (lanes may have different lengths).  Timestep t of every live lane is packed into one vector
(`join_time`, seqfunc/map.go:262-270) and handed to `B.batch(in, n)` -- the call that feeds the batched
device kernels -- and the packed result is split back per lane (`append_split`, map.go:272-284)."""
from __future__ import annotations

import numpy as np

from . import api

F64 = np.float64


def max_sequence_len(seqs):
    return max((len(s) for s in seqs), default=0)


def join_time(seqs, t):
    parts = [np.asarray(s[t], dtype=F64) for s in seqs if len(s) > t]
    return (np.concatenate(parts) if parts else np.zeros(0)), len(parts)


def append_split(seqs, t, dest, joined):
    _, n = join_time(seqs, t)
    if n == 0:
        return
    if len(joined) % n != 0:
        raise api.ShapeError(-1, "cannot evenly split vector")
    size = len(joined) // n
    k = 0
    for lane, s in enumerate(seqs):
        if len(s) > t:
            dest[lane].append(np.array(joined[k * size:(k + 1) * size], dtype=F64))
            k += 1


class ConstResult:
    """seqfunc/construct.go: a constant sequence list (gradient sink)."""

    def __init__(self, seqs):
        self.seqs = [[np.asarray(v, dtype=F64) for v in s] for s in seqs]

    def output_seqs(self):
        return self.seqs

    def propagate_gradient(self, u, g):
        pass


class ConstRResult(ConstResult):
    def routput_seqs(self):
        return [[np.zeros_like(v) for v in s] for s in self.seqs]

    def propagate_rgradient(self, u, ur, rg, g):
        pass


class VarResult:
    """seqfunc/construct.go VarResult: sequences of Variables, so upstreams reach their gradients."""

    def __init__(self, var_seqs):
        self.vars = var_seqs

    def output_seqs(self):
        return [[v.vector for v in s] for s in self.vars]

    def propagate_gradient(self, u, g):
        for lane, s in enumerate(self.vars):
            for t, v in enumerate(s):
                if v in g and lane < len(u) and t < len(u[lane]):
                    g[v] += u[lane][t]


class VarRResult(VarResult):
    def __init__(self, rv, var_seqs):
        super().__init__(var_seqs)
        self.rv = rv

    def routput_seqs(self):
        return [[np.asarray(self.rv[v], dtype=F64) if v in self.rv else np.zeros_like(v.vector) for v in s] for s in self.vars]

    def propagate_rgradient(self, u, ur, rg, g):
        for lane, s in enumerate(self.vars):
            for t, v in enumerate(s):
                if lane < len(u) and t < len(u[lane]):
                    if g is not None and v in g:
                        g[v] += u[lane][t]
                    if v in rg:
                        rg[v] += ur[lane][t]


class MapBatcher:
    def __init__(self, b):
        self.b = b

    def apply_seqs(self, r):
        """seqfunc/map_batcher.go:26-48.  A packed device list (anything with `.packed()`) never touches the host."""
        if hasattr(r, "packed"):
            return _map_packed(self.b, r, None, False)
        seqs = r.output_seqs()
        res = _MapBatcherResult(r, [[] for _ in seqs])
        for t in range(max_sequence_len(seqs)):
            joined, n = join_time(seqs, t)
            in_var = api.Variable(joined)
            out = self.b.batch(in_var, n)
            res.pool.append(in_var)
            res.results.append(out)
            append_split(seqs, t, res.output, out.output())
        return res


class _MapBatcherResult:
    def __init__(self, inp, output):
        self.input, self.output, self.pool, self.results = inp, output, [], []

    def output_seqs(self):
        return self.output

    def propagate_gradient(self, u, g):  # map_batcher.go:50-62: temp-variable protocol per timestep
        down = [[] for _ in u]
        for t in range(max_sequence_len(u)):
            upstream, _ = join_time(u, t)
            pv = self.pool[t]
            g[pv] = np.zeros(pv.vector.size, dtype=F64)
            self.results[t].propagate_gradient(upstream, g)
            append_split(u, t, down, g[pv])
            del g[pv]
        self.input.propagate_gradient(down, g)


class MapRBatcher:
    def __init__(self, b):
        self.b = b

    def apply_seqs(self, r):
        return MapBatcher(self.b).apply_seqs(r)

    def apply_seqs_r(self, rv, r):  # map_batcher.go:76-99
        if hasattr(r, "packed"):
            return _map_packed(self.b, r, rv, True)
        seqs, rseqs = r.output_seqs(), r.routput_seqs()
        res = _MapBatcherRResult(r, [[] for _ in seqs], [[] for _ in seqs])
        for t in range(max_sequence_len(seqs)):
            joined, n = join_time(seqs, t)
            joined_r, _ = join_time(rseqs, t)
            v = api.Variable(joined)
            in_var = api.RVariable(v, {v: joined_r})
            out = self.b.batch_r(rv, in_var, n)
            res.pool.append(v)
            res.results.append(out)
            append_split(seqs, t, res.output, out.output())
            append_split(rseqs, t, res.routput, out.routput())
        return res


class _MapBatcherRResult:
    def __init__(self, inp, output, routput):
        self.input, self.output, self.routput, self.pool, self.results = inp, output, routput, [], []

    def output_seqs(self):
        return self.output

    def routput_seqs(self):
        return self.routput

    def propagate_rgradient(self, u, ur, rg, g):  # map_batcher.go:117-140
        if g is None:
            g = {}
        down, down_r = [[] for _ in u], [[] for _ in u]
        for t in range(max_sequence_len(u)):
            upstream, _ = join_time(u, t)
            upstream_r, _ = join_time(ur, t)
            pv = self.pool[t]
            rg[pv] = np.zeros(pv.vector.size, dtype=F64)
            g[pv] = np.zeros(pv.vector.size, dtype=F64)
            self.results[t].propagate_rgradient(upstream, upstream_r, rg, g)
            append_split(u, t, down, g[pv])
            append_split(ur, t, down_r, rg[pv])
            del g[pv]
            del rg[pv]
        self.input.propagate_rgradient(down, down_r, rg, g)


# ======================================================================================================
# Device path: sequence lists packed time-major in ONE HBM buffer
# ======================================================================================================
class PackedSeqs:
    """A `[][]linalg.Vector` sequence list laid out the way the batcher wants to read it: timestep-major, and within
    a timestep the vectors of the lanes that are still alive, in lane order -- exactly the vector `joinTime` builds
    on the host for every step (seqfunc/map.go:262-270).  Packing happens once (one H2D copy); after that step t is
    a VIEW of the buffer, so mapping a Batcher over the list moves no data and splitting the result back
    (`appendSplit`, map.go:272-284) is deferred until a host caller asks for `to_host()`."""

    def __init__(self, lens, widths, data: api.DeviceVec):
        self.lens = [int(n) for n in lens]                     # lane lengths
        self.T = max(self.lens, default=0)
        self.counts = [sum(1 for n in self.lens if n > t) for t in range(self.T)]  # live lanes per step
        self.widths = [int(w) for w in widths]                 # vector size per step
        self.offsets = np.concatenate([[0], np.cumsum([c * w for c, w in zip(self.counts, self.widths)])]).astype(int)
        self.data = data
        if data.n != int(self.offsets[-1]):
            raise api.ShapeError(-1, "packed buffer does not match the lane lengths")

    @staticmethod
    def pack_host(seqs) -> tuple[list, list, np.ndarray]:
        lens = [len(s) for s in seqs]
        T = max(lens, default=0)
        parts, widths = [], []
        for t in range(T):
            joined, n = join_time(seqs, t)
            if len(joined) % n != 0 or any(len(s[t]) != len(joined) // n for s in seqs if len(s) > t):
                raise api.ShapeError(-1, "cannot evenly split vector")
            parts.append(joined)
            widths.append(len(joined) // n)
        return lens, widths, (np.concatenate(parts) if parts else np.zeros(0))

    @classmethod
    def from_host(cls, seqs) -> "PackedSeqs":
        lens, widths, flat = cls.pack_host(seqs)
        return cls(lens, widths, api.context().upload(flat))

    def like(self, widths, data) -> "PackedSeqs":
        return PackedSeqs(self.lens, widths, data)

    def step(self, t) -> tuple[api.DeviceVec, int]:
        return self.data.view(int(self.offsets[t]), int(self.offsets[t + 1])), self.counts[t]

    def unpack_host(self, flat) -> list:
        out = [[] for _ in self.lens]
        for t in range(self.T):
            append_split_lens(self.lens, t, out, flat[self.offsets[t]:self.offsets[t + 1]])
        return out

    def to_host(self) -> list:
        return self.unpack_host(self.data.numpy())


def append_split_lens(lens, t, dest, joined):
    n = sum(1 for ln in lens if ln > t)
    if n == 0:
        return
    size = len(joined) // n
    k = 0
    for lane, ln in enumerate(lens):
        if ln > t:
            dest[lane].append(np.array(joined[k * size:(k + 1) * size], dtype=F64))
            k += 1


def _packed_upstream(shape: PackedSeqs, widths, u) -> api.DeviceVec:
    """Upstream sequence lists arrive from a host caller as [][]Vector, or already packed from a device caller."""
    if isinstance(u, PackedSeqs):
        return u.data
    if isinstance(u, api.DeviceVec):
        return u
    _, w, flat = PackedSeqs.pack_host(u)
    if [len(s) for s in u] != shape.lens or w != list(widths):
        raise api.ShapeError(-1, "upstream sequence shape mismatch")
    return api.context().upload(flat)


class PackedConstResult:
    """seqfunc.ConstResult (seqfunc/construct.go) over a packed device list: a gradient sink."""

    def __init__(self, packed: PackedSeqs):
        self.p = packed

    @classmethod
    def from_host(cls, seqs):
        return cls(PackedSeqs.from_host(seqs))

    def packed(self):
        return self.p

    def packed_r(self):
        return None  # identically zero

    def output_seqs(self):
        return self.p.to_host()

    def routput_seqs(self):
        return [[np.zeros_like(v) for v in s] for s in self.p.to_host()]

    def constant(self, *maps):
        return True

    def propagate_gradient(self, u, g):
        pass

    def propagate_rgradient(self, u, ur, rg, g):
        pass


class PackedVarResult:
    """seqfunc.VarResult (seqfunc/construct.go) for a packed list: ONE Variable holds the whole packed buffer, so the
    sequence gradient arrives as one vector with the same time-major layout."""

    def __init__(self, var: api.Variable, lens, widths, rv=None):
        self.var, self.rv = var, rv
        self.p = PackedSeqs(lens, widths, var.device())

    def packed(self):
        return self.p

    def packed_r(self):
        if self.rv is None or self.var not in self.rv:
            return None
        return self.p.like(self.p.widths, api.RVariable(self.var, self.rv)._dr())

    def output_seqs(self):
        return self.p.to_host()

    def routput_seqs(self):
        pr = self.packed_r()
        return pr.to_host() if pr is not None else [[np.zeros_like(v) for v in s] for s in self.p.to_host()]

    def constant(self, *maps):
        return all(m is None or self.var not in m for m in maps)

    def propagate_gradient(self, u, g):
        if self.var in g:
            up = _packed_upstream(self.p, self.p.widths, u)
            g[self.var] += up if isinstance(g[self.var], api.DeviceVec) else api._host(up)

    def propagate_rgradient(self, u, ur, rg, g):
        if g is not None:
            self.propagate_gradient(u, g)
        self.propagate_gradient(ur, rg)


class _PackedMapResult:
    """Result of mapping a Batcher over a packed list: per-step device Results, outputs packed time-major."""

    def __init__(self, inp, shape: PackedSeqs, pool, results, out: PackedSeqs, out_r: PackedSeqs | None):
        self.input, self.shape, self.pool, self.results, self.out, self.out_r = inp, shape, pool, results, out, out_r

    def packed(self):
        return self.out

    def packed_r(self):
        return self.out_r

    def output_seqs(self):
        return self.out.to_host()

    def routput_seqs(self):
        return self.out_r.to_host()

    def constant(self, *maps):
        return self.input.constant(*maps) and all(r.constant(*maps) for r in self.results)

    # map_batcher.go:50-62 / :117-140 with the temp-variable protocol running on ONE backward session for all
    # timesteps: parameter gradients accumulate in HBM across the whole sequence and are flushed once.
    def propagate_gradient(self, u, g):
        up = _packed_upstream(self.out, self.out.widths, u)
        sess = api._Session()
        down = api.context().empty(self.shape.data.n)
        for t in range(self.shape.T):
            pv = self.pool[t]
            g[pv] = api._TempGrad(pv.size)
            a, b = int(self.out.offsets[t]), int(self.out.offsets[t + 1])
            api._down(self.results[t], sess, up.view(a, b).copy(), g)
            d = sess.take(g, pv)
            _copy_at(down, int(self.shape.offsets[t]), d)
        sess.flush()
        self.input.propagate_gradient(self.shape.like(self.shape.widths, down), g)

    def propagate_rgradient(self, u, ur, rg, g):
        if g is None:
            g = {}
        up = _packed_upstream(self.out, self.out.widths, u)
        up_r = _packed_upstream(self.out, self.out.widths, ur)
        sess = api._Session()
        ctx = api.context()
        down, down_r = ctx.empty(self.shape.data.n), ctx.empty(self.shape.data.n)
        for t in range(self.shape.T):
            pv = self.pool[t]
            g[pv], rg[pv] = api._TempGrad(pv.size), api._TempGrad(pv.size)
            a, b = int(self.out.offsets[t]), int(self.out.offsets[t + 1])
            api._down_r(self.results[t], sess, up.view(a, b).copy(), up_r.view(a, b).copy(), rg, g)
            _copy_at(down, int(self.shape.offsets[t]), sess.take(g, pv))
            _copy_at(down_r, int(self.shape.offsets[t]), sess.take(rg, pv))
        sess.flush()
        like = self.shape.like
        self.input.propagate_rgradient(like(self.shape.widths, down), like(self.shape.widths, down_r), rg, g)


def _copy_at(dst: api.DeviceVec, offset: int, src: api.DeviceVec):
    ctx = api.context()
    if src.n:
        api.check(ctx.lib.af_copy(ctx.h, dst.ptr + 8 * offset, src.ptr, src.n))


def _map_packed(b, r, rv, with_r):
    shape = r.packed()
    shape_r = r.packed_r() if with_r else None
    ctx = api.context()
    pool, results, outs, outs_r, widths = [], [], [], [], []
    for t in range(shape.T):
        x, n = shape.step(t)
        pv = api.Variable.from_device(x)
        if with_r:
            xr = shape_r.step(t)[0] if shape_r is not None else None
            res = b.batch_r(rv, api.RVariable(pv, {pv: xr} if xr is not None else {}), n)
        else:
            res = b.batch(pv, n)
        pool.append(pv)
        results.append(res)
        d = api._in_d(res)
        if d.n % n != 0:
            raise api.ShapeError(-1, "cannot evenly split vector")
        widths.append(d.n // n)
        outs.append(d)
        if with_r:
            outs_r.append(api._in_dr(res))
    total = sum(d.n for d in outs)
    out = ctx.empty(total)
    out_r = ctx.zeros(total) if with_r else None
    off = 0
    for i, d in enumerate(outs):
        _copy_at(out, off, d)
        if with_r and outs_r[i] is not None:
            _copy_at(out_r, off, outs_r[i])
        off += d.n
    return _PackedMapResult(r, shape, pool, results, shape.like(widths, out),
                            shape.like(widths, out_r) if with_r else None)
