#!/usr/bin/env python
"""A few launches of few_rows_wgrad_kernel at the bench head's shape (for ncu: -k regex:few_rows_wgrad)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from autofunc_b200 import api

ctx = api.context()
ctx.set_option("few_rows_wgrad", 2)
rows, cols, n = 10, 512, 4096
rng = np.random.default_rng(0)
X, RX, U, UR = (ctx.upload(rng.uniform(-1, 1, k)) for k in (n * cols, n * cols, n * rows, n * rows))
G, RG = ctx.zeros(rows * cols), ctx.zeros(rows * cols)
for _ in range(5):
    api.check(ctx.lib.af_lintran_wgrad_r(ctx.h, U.ptr, UR.ptr, X.ptr, RX.ptr, G.ptr, RG.ptr, rows, cols, n))
print(float(np.abs(G.numpy()).sum()))
