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
