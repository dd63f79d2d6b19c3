#!/usr/bin/env python
"""One eager forward + backward of the C4 LSTM plan (the first call of a plan never replays a graph): the command to put
under ncu for a look at the fused step kernels.
    ncu --set full --clock-control none --import-source on -k regex:gemm_dmma -s 10 -c 1 -o fwd python tools/ncu_lstm.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from autofunc_b200 import api, synth
from autofunc_b200.lstm import LSTMPlan

I, H, O, T, B = 30, 512, 10, 128, 256
ctx = api.context()
plan = LSTMPlan(I, H, O, T, B, ctx)
for i in range(10):
    plan.set_param(i, synth.uniform(900 + i, plan.param_len(i)) * (0.05 if i % 2 == 0 else 1.0))
    plan.set_rvector(i, synth.uniform(950 + i, plan.param_len(i)))
plan.set_inputs(ctx.upload(synth.uniform(11, T * B * I, 0, 1)))
up = ctx.upload(synth.uniform(12, T * B * O, 0, 1))
api.check(ctx.lib.af_copy(ctx.h, plan.upstream().ptr, up.ptr, len(up)))
plan.zero_grads()
plan.forward(True)
plan.backward(True)
ctx.sync()
print("ok", ctx.launch_count())
