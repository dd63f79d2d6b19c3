"""Device-resident Concat / Slice / Repeat (slices.go:13-335) for the operator API.

Concat joins whole vectors end to end (slices.go:16-33); for a packed batch whose parts must be
interleaved per sample -- the LSTM's concat(state, x) per lane -- use `concat_batched`, the batched
form built on af_copy_2d."""
from __future__ import annotations

import numpy as np

from . import _lib
from .api import (DeviceVec, ShapeError, _DeviceResult, _down, _down_r, _in_d, _in_dr, _in_len, _p, check, context)


def _copy_into(dst: DeviceVec, offset: int, src: DeviceVec):
    ctx = context()
    check(ctx.lib.af_copy(ctx.h, dst.ptr + 8 * offset, src.ptr, src.n))


def _zeros_like_len(n):
    return context().zeros(n)


class _Joined(_DeviceResult):
    def __init__(self, results, d, dr=None):
        super().__init__(d, dr, dr is not None)
        self.results = results
        self.lens = [_in_len(r) for r in results]

    def constant(self, *maps):
        return all(r.constant(*maps) for r in self.results)

    def _bwd(self, sess, dU, grad):  # slices.go:47-56: sub-slice views of the upstream
        i = 0
        for r, l in zip(self.results, self.lens):
            if not r.constant(grad):
                _down(r, sess, dU.view(i, i + l), grad)
            i += l

    def _bwd_r(self, sess, dU, dUR, rgrad, grad):  # slices.go:108-119
        i = 0
        for r, l in zip(self.results, self.lens):
            if not r.constant(rgrad, grad):
                _down_r(r, sess, dU.view(i, i + l), dUR.view(i, i + l), rgrad, grad)
            i += l


def concat(*rs):
    total = sum(_in_len(r) for r in rs)
    out = context().empty(total)
    i = 0
    for r in rs:
        d = _in_d(r)
        _copy_into(out, i, d)
        i += d.n
    return _Joined(list(rs), out)


def concat_r(*rs):
    total = sum(_in_len(r) for r in rs)
    ctx = context()
    out, outr = ctx.empty(total), ctx.zeros(total)
    i = 0
    for r in rs:
        d, dr = _in_d(r), _in_dr(r)
        _copy_into(out, i, d)
        if dr is not None:
            _copy_into(outr, i, dr)
        i += d.n
    return _Joined(list(rs), out, outr)


class _Sliced(_DeviceResult):
    def __init__(self, inp, start, end, d, dr=None):
        super().__init__(d, dr, dr is not None)
        self.input, self.start, self.end = inp, start, end

    def constant(self, *maps):
        return self.input.constant(*maps)

    def _scatter(self, dU):
        full = context().zeros(_in_len(self.input))
        _copy_into(full, self.start, dU)
        return full

    def _bwd(self, sess, dU, grad):  # slices.go:152-166: zero-padded downstream
        if not self.input.constant(grad):
            _down(self.input, sess, self._scatter(dU), grad)

    def _bwd_r(self, sess, dU, dUR, rgrad, grad):  # slices.go:190-206
        if not self.input.constant(rgrad, grad):
            _down_r(self.input, sess, self._scatter(dU), self._scatter(dUR), rgrad, grad)


def slice_(inp, start, end):
    if not (0 <= start <= end <= _in_len(inp)):
        raise ShapeError(_lib.AF_ESHAPE, "slice bounds out of range")
    return _Sliced(inp, start, end, _in_d(inp).view(start, end).copy())


def slice_r(inp, start, end):
    if not (0 <= start <= end <= _in_len(inp)):
        raise ShapeError(_lib.AF_ESHAPE, "slice bounds out of range")
    dr = _in_dr(inp)
    r = dr.view(start, end).copy() if dr is not None else context().zeros(end - start)
    return _Sliced(inp, start, end, _in_d(inp).view(start, end).copy(), r)


class _Repeated(_DeviceResult):
    def __init__(self, inp, n, d, dr=None):
        super().__init__(d, dr, dr is not None)
        self.repeated, self.n = inp, n

    def constant(self, *maps):
        return self.repeated.constant(*maps)

    def _fold(self, dU):
        """Sum of the n parts = column sums of the n x len view (slices.go:244-250)."""
        ctx = context()
        l = _in_len(self.repeated)
        acc = ctx.zeros(l)
        check(ctx.lib.af_bias_grad(ctx.h, dU.ptr, acc.ptr, l, self.n))
        return acc

    def _bwd(self, sess, dU, grad):
        if self.repeated.constant(grad):
            return
        if dU.n != self.d.n:
            raise ShapeError(_lib.AF_ESHAPE, "invalid upstream size")
        _down(self.repeated, sess, self._fold(dU), grad)

    def _bwd_r(self, sess, dU, dUR, rgrad, grad):
        if self.repeated.constant(rgrad, grad):
            return
        if dU.n != self.d.n:
            raise ShapeError(_lib.AF_ESHAPE, "invalid upstream size")
        _down_r(self.repeated, sess, self._fold(dU), self._fold(dUR), rgrad, grad)


def _tile(d: DeviceVec, n: int) -> DeviceVec:
    out = context().empty(d.n * n)
    for i in range(n):  # n is small in the reference's uses (batch-size repeats of a bias vector)
        _copy_into(out, i * d.n, d)
    return out


def repeat(inp, n):  # slices.go:216-228
    return _Repeated(inp, n, _tile(_in_d(inp), n))


def repeat_r(inp, n):  # slices.go:260-275
    dr = _in_dr(inp)
    r = _tile(dr, n) if dr is not None else context().zeros(_in_len(inp) * n)
    return _Repeated(inp, n, _tile(_in_d(inp), n), r)


# ---- per-sample concatenation of packed batches --------------------------------------------------
class _JoinedBatched(_DeviceResult):
    def __init__(self, results, widths, n, d, dr=None):
        super().__init__(d, dr, dr is not None)
        self.results, self.widths, self.n = results, widths, n

    def constant(self, *maps):
        return all(r.constant(*maps) for r in self.results)

    def _part(self, dU, off, w):
        ctx = context()
        out = ctx.empty(self.n * w)
        check(ctx.lib.af_copy_2d(ctx.h, out.ptr, w, dU.ptr + 8 * off, sum(self.widths), self.n, w))
        return out

    def _bwd(self, sess, dU, grad):
        off = 0
        for r, w in zip(self.results, self.widths):
            if not r.constant(grad):
                _down(r, sess, self._part(dU, off, w), grad)
            off += w

    def _bwd_r(self, sess, dU, dUR, rgrad, grad):
        off = 0
        for r, w in zip(self.results, self.widths):
            if not r.constant(rgrad, grad):
                _down_r(r, sess, self._part(dU, off, w), self._part(dUR, off, w), rgrad, grad)
            off += w


def _concat_batched(rs, n, with_r):
    ctx = context()
    widths = []
    for r in rs:
        l = _in_len(r)
        if l % n != 0:
            raise ShapeError(_lib.AF_ESHAPE, "input count does not divide input length")
        widths.append(l // n)
    total = sum(widths)
    out = ctx.empty(n * total)
    outr = ctx.zeros(n * total) if with_r else None
    off = 0
    for r, w in zip(rs, widths):
        check(ctx.lib.af_copy_2d(ctx.h, out.ptr + 8 * off, total, _in_d(r).ptr, w, n, w))
        if with_r:
            dr = _in_dr(r)
            if dr is not None:
                check(ctx.lib.af_copy_2d(ctx.h, outr.ptr + 8 * off, total, dr.ptr, w, n, w))
        off += w
    return _JoinedBatched(list(rs), widths, n, out, outr)


def concat_batched(n, *rs):
    """Packed-batch Concat: sample b of the result is concat(sample b of each part)."""
    return _concat_batched(rs, n, False)


def concat_batched_r(n, *rs):
    return _concat_batched(rs, n, True)
