#!/usr/bin/env python
"""Timing of the BASELINE.json configs other than the headline one (which bench.py owns):
  C1  bench/lintran.go  LinTran 1024x2048, n=10, fwd + Gradient + RGradient
  C2' bench/mlp.go      MLP 1000-2000-512-10 at n=1 (reference-exact) and n=64/1024
  C3  Hessian-vector step on a 4 x (2048->2048) MLP, n=4096
  C4  bench/lstm.go     LSTM H=512, T=128, B=256 (I=30, O=10), forward + BPTT + R
  C5  one rank's share of 4 x (8192->8192), n=8192 (the 8-GPU shard of n=65536)
Inputs resident, CUDA events, 3 warm-ups.  One JSON line per config on stdout.

Under torchrun (WORLD_SIZE > 1) every rank runs the same per-GPU problem on its own shard of the batch (weak scaling)
and each step ends with the NCCL sum all-reduce of the parameter-gradient accumulators ({grad | rgrad} slab); the
time is the max over ranks and the rates are whole-job aggregates."""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from autofunc_b200 import api, synth
from autofunc_b200.lstm import LSTMPlan


RANK, WORLD, LOCAL = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))


def reduce_slab(ptr, n):
    """Gradient.Add across ranks (gradient.go:28-30): the library's own collective (af_allreduce_sum: NCCL behind the C
    ABI, csrc/dp.cu) over a device range, in order on the step's stream."""
    if WORLD > 1:
        ctx = api.context()
        api.check(ctx.lib.af_allreduce_sum(ctx.h, ptr, n))


def timed(fn, steps, warmup=3, after=None):
    for _ in range(warmup):
        fn()
    if after:
        after()
    torch.cuda.synchronize()
    if WORLD > 1:
        import torch.distributed as dist
        dist.barrier()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    if after:
        after()
    e1.record()
    e1.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if WORLD > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def emit(d):
    if RANK == 0:
        d["n_gpus"] = WORLD
        print(json.dumps(d), flush=True)


def mlp_config(ctx, name, sizes, n, steps, flops):
    plan = api.MLPPlan(sizes, n, ctx)
    params, rvecs = synth.mlp_params(sizes, weight_scale=lambda k: 1 / np.sqrt(k))
    for i, (p, r) in enumerate(zip(params, rvecs)):
        plan.set_param(i, p)
        plan.set_rvector(i, r)
    x = synth.mlp_input(sizes, n, seed=124 + 7919 * RANK)
    ctx.lib.af_upload(ctx.h, plan.input().ptr, x.ctypes.data, x.size)
    n_out = n * sizes[-1]
    up, upr = plan.upstream(), plan.upstream_r()
    slab_ptr, slab_n = plan.grad(0).ptr, (plan.input().ptr - plan.grad(0).ptr) // 8

    # one bench iteration: fresh accumulators, outGrad = 1, outRGrad = 0, RunR(backProp=true); under torchrun the sum over
    # ranks is scheduled by the library (af_mlp_step_fresh_dp, deferred join: DESIGN.md 5); the last join is inside the timing
    mode = {"blocking": 1, "overlap": 2, "deferred": 3}[os.environ.get("AF_DP_MODE", "deferred")]
    plan.set_dp(mode if WORLD > 1 else 0)

    def step():
        plan.step_fresh_dp(True)
    ms = timed(step, steps, after=plan.dp_join)
    if os.environ.get("AF_TRACE"):
        cap = 256
        kinds, work, msa, cnt = (C.c_int * cap)(), (C.c_double * cap)(), (C.c_float * cap)(), C.c_int(0)
        api.check(ctx.lib.af_trace_begin(ctx.h, cap))
        step()
        api.check(ctx.lib.af_trace_end(ctx.h, cap, kinds, work, msa, C.byref(cnt)))
        for i in range(cnt.value):
            print(f"  [{i:2d}] kind={kinds[i]} work={work[i]:.4g} {msa[i] * 1e3:7.1f} us", file=sys.stderr)
    plan.close()
    emit({"config": name, "sizes": sizes, "n_per_gpu": n, "ms_per_step": ms, "samples_per_s": WORLD * n / ms * 1e3,
          "tflops_executed": WORLD * flops / ms / 1e9, "allreduce_MB": slab_n * 8 / 1e6 if WORLD > 1 else 0})


def executed_flops(sizes, n):
    f = 0
    for l in range(len(sizes) - 1):
        f += (8 if l == 0 else 18) * sizes[l] * sizes[l + 1]  # layer 1: zero-RX products are skipped
    return f * n


def lintran_config(ctx, steps):
    rows, cols, n = 1024, 2048, 10
    W, RW = synth.uniform(1, rows * cols), synth.uniform(2, rows * cols)
    X, RX = synth.uniform(3, n * cols), synth.uniform(4, n * cols)
    U, UR = synth.uniform(5, n * rows), synth.uniform(6, n * rows)
    d = {k: ctx.upload(v) for k, v in dict(W=W, RW=RW, X=X, RX=RX, U=U, UR=UR).items()}
    Y, RY, D, DR = ctx.empty(n * rows), ctx.empty(n * rows), ctx.empty(n * cols), ctx.empty(n * cols)
    gW, rgW, gX, rgX = ctx.zeros(rows * cols), ctx.zeros(rows * cols), ctx.zeros(n * cols), ctx.zeros(n * cols)
    L, h = ctx.lib, ctx.h

    split = os.environ.get("AF_C1_SPLIT") == "1"  # A/B: the round-1 call sequence (three passes + two accumulations)

    only = os.environ.get("AF_C1_ONLY", "")  # timing experiments: "fwd" or "bwd" alone

    def step():  # bench/lintran.go:83-89: BatchR + PropagateRGradient with W and X both in the maps
        if only != "bwd":
            api.check(L.af_lintran_fwd_r(h, d["W"].ptr, d["RW"].ptr, d["X"].ptr, d["RX"].ptr, Y.ptr, RY.ptr, rows, cols, n))
        if only == "fwd":
            return
        if split:
            api.check(L.af_lintran_wgrad_r(h, d["U"].ptr, d["UR"].ptr, d["X"].ptr, d["RX"].ptr, gW.ptr, rgW.ptr, rows, cols, n))
            api.check(L.af_lintran_igrad_r(h, d["U"].ptr, d["UR"].ptr, d["W"].ptr, d["RW"].ptr, D.ptr, DR.ptr, rows, cols, n))
            api.check(L.af_axpy(h, 1.0, D.ptr, gX.ptr, n * cols))
            api.check(L.af_axpy(h, 1.0, DR.ptr, rgX.ptr, n * cols))
        else:
            # linTranRResult.PropagateRGradient (lin_tran.go:393-425) as one call: both weight gradients, and the input
            # gradients added into the input variable's accumulators (variable.go:113-123)
            api.check(L.af_lintran_backward_r(h, d["U"].ptr, d["UR"].ptr, d["W"].ptr, d["RW"].ptr, d["X"].ptr, d["RX"].ptr,
                                              gW.ptr, rgW.ptr, gX.ptr, rgX.ptr, rows, cols, n, 1))
        reduce_slab(gW.ptr, rows * cols)   # the parameter's accumulators; X's gradients stay with their shard
        reduce_slab(rgW.ptr, rows * cols)
    ms_eager = timed(step, steps)
    ms = ms_eager
    if WORLD == 1 and os.environ.get("AF_C1_GRAPH", "1") != "0":
        # the seven launches of the step recorded once and replayed as one submission (af_graph_*)
        with ctx.record() as g:
            step()
        ms = timed(g.launch, steps)
    if os.environ.get("AF_TRACE"):
        cap = 256
        kinds, work, msa, cnt = (C.c_int * cap)(), (C.c_double * cap)(), (C.c_float * cap)(), C.c_int(0)
        api.check(ctx.lib.af_trace_begin(ctx.h, cap))
        step()
        api.check(ctx.lib.af_trace_end(ctx.h, cap, kinds, work, msa, C.byref(cnt)))
        for i in range(cnt.value):
            print(f"  [{i:2d}] kind={kinds[i]} work={work[i]:.4g} {msa[i] * 1e3:7.1f} us", file=sys.stderr)
    algo_bytes = 8 * 8 * rows * cols  # W, RW read twice + gW, rgW read-modify-write
    emit({"config": "C1 lintran 1024x2048 n=10 fwd+Grad+RGrad", "n_per_gpu": n, "ms_per_step": ms, "us_per_step": ms * 1e3,
          "us_per_step_eager": ms_eager * 1e3,
          "algorithmic_GBps": WORLD * algo_bytes / ms / 1e6, "samples_per_s": WORLD * n / ms * 1e3})


def lstm_config(ctx, I, H, O, T, B, steps):
    plan = LSTMPlan(I, H, O, T, B, ctx)
    for i in range(10):
        plan.set_param(i, synth.uniform(900 + i, plan.param_len(i)) * (0.05 if i % 2 == 0 else 1.0))
        plan.set_rvector(i, synth.uniform(950 + i, plan.param_len(i)))
    x = ctx.upload(synth.uniform(11, T * B * I, 0, 1))
    upv = ctx.upload(synth.uniform(12, T * B * O, 0, 1))
    plan.set_inputs(x)
    up, upr = plan.upstream(), plan.upstream_r()

    def step():
        plan.zero_grads()
        api.check(ctx.lib.af_copy(ctx.h, up.ptr, upv.ptr, up.n))
        api.check(ctx.lib.af_zero(ctx.h, upr.ptr, upr.n))
        plan.forward(True)
        plan.backward(True)
        reduce_slab(slab.ptr, slab.n)
    slab = plan.grad_slab()
    ms = timed(step, steps)
    fwd_ms = timed(lambda: plan.forward(True), steps)

    def bwd_only():
        api.check(ctx.lib.af_copy(ctx.h, up.ptr, upv.ptr, up.n))
        plan.backward(True)
    bwd_ms = timed(bwd_only, steps)
    if os.environ.get("AF_TRACE"):
        cap = 4096
        kinds, work, msa, cnt = (C.c_int * cap)(), (C.c_double * cap)(), (C.c_float * cap)(), C.c_int(0)
        api.check(ctx.lib.af_trace_begin(ctx.h, cap))
        step()
        api.check(ctx.lib.af_trace_end(ctx.h, cap, kinds, work, msa, C.byref(cnt)))
        agg = {}
        for i in range(cnt.value):
            key = (kinds[i], round(work[i]))
            a = agg.setdefault(key, [0, 0.0])
            a[0] += 1
            a[1] += msa[i]
        for (k, w), (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print(f"  kind={k} work={w:.4g} launches={c} total={t:.3f} ms avg={t / c * 1e3:.1f} us rate={w * c / t / 1e9 if t else 0:.2f} (TF/s or TB/s)", file=sys.stderr)
    flops = 18.0 * (4 * H + O) * (H + I) * T * B
    plan.close()
    emit({"config": f"C4 lstm I={I} H={H} O={O} T={T} B={B} fwd+BPTT+R", "lanes_per_gpu": B, "ms_per_step": ms,
          "forward_ms": fwd_ms, "backward_ms": bwd_ms,
          "sequences_per_s": WORLD * B / ms * 1e3, "tflops_reference_count": WORLD * flops / ms / 1e9})


def main():
    which = set(sys.argv[1:]) or {"c1", "c2", "c3", "c4"}
    torch.cuda.set_device(LOCAL)
    if WORLD > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", LOCAL))
    stream = torch.cuda.Stream()  # a real stream (the legacy default stream cannot be graph-captured)
    torch.cuda.set_stream(stream)
    ctx = api.Context(LOCAL, stream=stream.cuda_stream)
    api.set_context(ctx)
    if WORLD > 1:
        from autofunc_b200 import parallel
        parallel.dp_init_torch(ctx, LOCAL)  # torch.distributed only ships the communicator id
    for opt in ("pdl", "small_mlp", "small_mlp_head", "small_n", "lstm_split", "lstm_fused", "persist_tiles", "sk_bias", "few_samples", "few_samples_min", "few_samples_blocks", "few_samples_tma", "deterministic"):  # A/B switches: AF_OPT_PDL=0 etc.
        v = os.environ.get("AF_OPT_" + opt.upper())
        if v is not None:
            api.check(ctx.lib.af_set_option(ctx.h, opt.encode(), int(v)))
    if "c1" in which:
        lintran_config(ctx, 50)
    if "c2" in which:
        for n in (1, 4, 8, 10, 16, 64, 1024):
            mlp_config(ctx, f"C2 mlp n={n}", [1000, 2000, 512, 10], n, 20, executed_flops([1000, 2000, 512, 10], n))
    if "c3" in which:
        s = [2048] * 5
        mlp_config(ctx, "C3 hessian-vector 4x(2048->2048) n=4096", s, 4096, 5, executed_flops(s, 4096))
    if "c4" in which:
        lstm_config(ctx, 30, 512, 10, 128, 256, 3)
    if "c4a" in which:  # the aligned variant SURVEY 8 asks for next to the reference's I = 30, O = 10
        lstm_config(ctx, 512, 512, 512, 128, 256, 3)
    if "c5" in which:
        s = [8192] * 5
        mlp_config(ctx, "C5 shard 4x(8192->8192) n=8192 (1/8 of n=65536)", s, 8192, 2, executed_flops(s, 8192))
    if WORLD > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
