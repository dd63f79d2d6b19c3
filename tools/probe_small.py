#!/usr/bin/env python
"""Where the small-batch MLP step's time goes: forward only / full step / the zeroing around it."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from autofunc_b200 import api, synth

def timed(fn, steps=200, warmup=20):
    for _ in range(warmup): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / steps * 1e3

stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ctx = api.Context(0, stream=stream.cuda_stream); api.set_context(ctx)
import ast
sizes = ast.literal_eval(os.environ.get('AF_SIZES', '[1000, 2000, 512, 10]'))
for n in (1, 2, 4, 8):
    plan = api.MLPPlan(sizes, n, ctx)
    params, rvecs = synth.mlp_params(sizes, weight_scale=lambda k: 1 / np.sqrt(k))
    for i, (p, r) in enumerate(zip(params, rvecs)):
        plan.set_param(i, p); plan.set_rvector(i, r)
    x = synth.mlp_input(sizes, n)
    ctx.lib.af_upload(ctx.h, plan.input().ptr, x.ctypes.data, x.size)
    up, upr = plan.upstream(), plan.upstream_r()
    api.check(ctx.lib.af_fill(ctx.h, up.ptr, 1.0, n * 10)); api.check(ctx.lib.af_zero(ctx.h, upr.ptr, n * 10))
    res = {"n": n}
    for opt in (1, 0):
        api.check(ctx.lib.af_set_option(ctx.h, b"small_mlp", opt))
        tag = "persistent" if opt else "graph"
        res[f"{tag}_fwd_us"] = round(timed(lambda: plan.step(True, False)), 1)
        res[f"{tag}_step_us"] = round(timed(lambda: plan.step(True, True)), 1)
        res[f"{tag}_nonR_step_us"] = round(timed(lambda: plan.step(False, True)), 1)
    api.check(ctx.lib.af_set_option(ctx.h, b"small_mlp", 1))
    res["persistent_fresh_step_us"] = round(timed(lambda: plan.step_fresh(True)), 1)
    res["zero_grads_us"] = round(timed(lambda: plan.zero_grads()), 1)
    res["fill_zero_us"] = round(timed(lambda: (ctx.lib.af_fill(ctx.h, up.ptr, 1.0, n * 10), ctx.lib.af_zero(ctx.h, upr.ptr, n * 10))), 1)
    print(json.dumps(res), flush=True)
    plan.close()
