#!/usr/bin/env python
"""HBM roofline of the streaming kernels (elementwise.cu, mathops.cu): algorithmic bytes per element x elements /
device time (CUDA events, 3 warm-ups, operands far larger than the 126 MB L2 so every pass comes from HBM),
against the measured copy bandwidth in MEASURED_PEAKS.json.  One JSON line per kernel on stdout."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from autofunc_b200 import api


def main():
    n = int(os.environ.get("AF_STREAM_N", 1 << 26))  # 64 Mi doubles = 512 MiB per vector
    steps = int(os.environ.get('AF_STREAM_STEPS', 10))
    warm = int(os.environ.get('AF_STREAM_WARM', 30))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = api.Context(0, stream=stream.cuda_stream)
    api.set_context(ctx)
    L, h = ctx.lib, ctx.h
    peak = None
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    rng = np.random.default_rng(0)
    base = rng.uniform(0.1, 1.0, n)
    bufs = [ctx.upload(base) for _ in range(6)]
    pristine = ctx.upload(base)
    a, b, c, d, e, f = (v.ptr for v in bufs)
    s1 = ctx.zeros(2)
    rows = n // 1024
    cases = [
        # name, algorithmic bytes per element, call
        ("copy (af_copy)", 16, lambda: L.af_copy(h, a, b, n)),
        ("axpy (gradient.go:40-52)", 24, lambda: L.af_axpy(h, 1e-9, b, a, n)),
        ("add", 24, lambda: L.af_add(h, a, b, c, n)),
        ("mul_bwd R", 56, lambda: L.af_mul_bwd(h, a, b, c, d, e, f, n)),
        ("sigmoid_fwd_r", 32, lambda: L.af_sigmoid_fwd_r(h, a, b, c, d, n)),
        ("sigmoid_bwd_r", 48, lambda: L.af_sigmoid_bwd_r(h, a, b, c, d, n)),
        ("exp_fwd_r", 32, lambda: L.af_exp_fwd_r(h, a, b, c, d, n)),
        ("tanh_fwd_r", 32, lambda: L.af_tanh_fwd_r(h, a, b, c, d, n)),
        ("logsigmoid_bwd_r", 48, lambda: L.af_logsigmoid_bwd_r(h, a, b, c, d, n)),
        ("sum_all R", 16, lambda: L.af_sum_all(h, a, b, s1.ptr, s1.ptr + 8, n)),
        ("dot", 16, lambda: L.af_dot(h, a, b, None, None, s1.ptr, n)),
        ("sqerr_fwd_r", 32, lambda: L.af_sqerr_fwd_r(h, a, b, c, d, s1.ptr, s1.ptr + 8, n)),
        ("bias_grad_r (colsum 65536 x 1024)", 16, lambda: L.af_bias_grad_r(h, a, b, c, d, 1024, rows)),
        ("softmax_fwd_r rows of 1024", 32, lambda: L.af_softmax_fwd_r(h, a, b, 0.0, c, d, 1024, rows)),
        ("softmax_bwd_r rows of 1024", 48, lambda: L.af_softmax_bwd_r(h, a, b, 0.0, c, d, 1024, rows)),
        ("softmax_fwd_r rows of 10", 32, lambda: L.af_softmax_fwd_r(h, a, b, 0.0, c, d, 10, n // 10)),
        ("softmax_bwd_r rows of 10", 48, lambda: L.af_softmax_bwd_r(h, a, b, 0.0, c, d, 10, n // 10)),
        ("softmax_fwd_r rows of 16", 32, lambda: L.af_softmax_fwd_r(h, a, b, 0.0, c, d, 16, n // 16)),
        ("softmax_bwd_r rows of 16", 48, lambda: L.af_softmax_bwd_r(h, a, b, 0.0, c, d, 16, n // 16)),
        ("softmax_fwd_r rows of 32", 32, lambda: L.af_softmax_fwd_r(h, a, b, 0.0, c, d, 32, n // 32)),
        ("softmax_bwd_r rows of 32", 48, lambda: L.af_softmax_bwd_r(h, a, b, 0.0, c, d, 32, n // 32)),
        ("softmax_fwd_r rows of 4", 32, lambda: L.af_softmax_fwd_r(h, a, b, 0.0, c, d, 4, n // 4)),
        # A/B: the staged group kernels that served short rows before (af_set_option "short_rows" 0)
        ("softmax_fwd_r rows of 10 [short_rows=0]", 32, lambda: L.af_softmax_fwd_r(h, a, b, 0.0, c, d, 10, n // 10)),
        ("softmax_bwd_r rows of 10 [short_rows=0]", 48, lambda: L.af_softmax_bwd_r(h, a, b, 0.0, c, d, 10, n // 10)),
    ]
    only = os.environ.get("AF_STREAM_ONLY")  # substring filter, e.g. AF_STREAM_ONLY=softmax
    for name, bpe, call in cases:
        if only and only not in name:
            continue
        ctx.set_option("short_rows", 0 if "[short_rows=0]" in name else 1)
        for v in bufs:  # the backward kernels overwrite their operands in place: restore (device to device)
            api.check(L.af_copy(h, v.ptr, pristine.ptr, n))
        for _ in range(warm):  # ~15 ms of back-to-back launches: clocks are up before the timed region
            api.check(call())
        for v in bufs:
            api.check(L.af_copy(h, v.ptr, pristine.ptr, n))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            api.check(call())
        e1.record(stream)
        e1.synchronize()
        ms = e0.elapsed_time(e1) / steps
        gbs = bpe * n / ms / 1e6
        print(json.dumps({"kernel": name, "elements": n, "algorithmic_bytes_per_element": bpe, "ms": round(ms, 4),
                          "achieved_GBps": round(gbs, 1), "peak_GBps": peak,
                          "frac": round(gbs / peak, 3) if peak else None}), flush=True)


if __name__ == "__main__":
    main()
