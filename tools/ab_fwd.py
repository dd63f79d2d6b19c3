"""A/B timing of single contraction launches (used while tuning): forward layer shapes of the bench."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from autofunc_b200 import api, synth
ctx = api.Context(0, stream=torch.cuda.current_stream().cuda_stream); api.set_context(ctx)
L, h = ctx.lib, ctx.h
def t(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); e1.synchronize(); return e0.elapsed_time(e1) / it * 1e3
for (rows, cols, n, rx) in [(2000, 1000, 4096, False), (512, 2000, 4096, True), (2048, 2048, 4096, True), (2048, 542, 256, True)]:
    W, RW, b, Rb = (ctx.upload(synth.uniform(i, k)) for i, k in ((1, rows * cols), (2, rows * cols), (3, rows), (4, rows)))
    X, RX = ctx.upload(synth.uniform(5, n * cols)), ctx.upload(synth.uniform(6, n * cols))
    S, RS = ctx.empty(n * rows), ctx.empty(n * rows)
    U, UR = ctx.upload(synth.uniform(7, n * rows)), ctx.upload(synth.uniform(8, n * rows))
    D, DR = ctx.empty(n * cols), ctx.empty(n * cols)
    gW, rgW = ctx.zeros(rows * cols), ctx.zeros(rows * cols)
    f = t(lambda: api.check(L.af_dense_fwd(h, W.ptr, RW.ptr, b.ptr, Rb.ptr, X.ptr, RX.ptr if rx else None, S.ptr, RS.ptr, rows, cols, n, 1)))
    f0 = t(lambda: api.check(L.af_dense_fwd(h, W.ptr, RW.ptr, b.ptr, Rb.ptr, X.ptr, RX.ptr if rx else None, S.ptr, RS.ptr, rows, cols, n, 0)))
    ig = t(lambda: api.check(L.af_dense_igrad(h, U.ptr, UR.ptr, W.ptr, RW.ptr, None, None, D.ptr, DR.ptr, rows, cols, n)))
    wg = t(lambda: api.check(L.af_lintran_wgrad_r(h, U.ptr, UR.ptr, X.ptr, RX.ptr if rx else None, gW.ptr, rgW.ptr, rows, cols, n)))
    print(f"rows={rows} cols={cols} n={n}: fwd+sigmoid {f:8.1f} us  fwd+bias {f0:8.1f} us  igrad {ig:8.1f} us  wgrad {wg:8.1f} us")
