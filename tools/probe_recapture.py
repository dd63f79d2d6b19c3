#!/usr/bin/env python
"""A plan whose step graph is re-captured right after an option change (no eager run in between): the fix-up workspace must
exist already (it is allocated with the context).  Prints the worst deviation between the two modes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from autofunc_b200 import api
from oracle import autofunc_ref as ref

ctx = api.context()
sizes, n = [1000, 2000, 512, 10], 256
onet = ref.MLP(sizes, weight_scale=lambda k: 1 / np.sqrt(k))
plan = api.MLPPlan(sizes, n)
for i, v in enumerate(onet.variables):
    plan.set_param(i, v.vector)
    plan.set_rvector(i, onet.rvec[v])
x = onet.make_input(n).vector
for _ in range(3):
    a = plan.step_host(x)
for det in (1, 0, 1):
    ctx.set_option("deterministic", det)
    for _ in range(3):
        b = plan.step_host(x)
    worst = max(float(np.abs(p - q).max()) for p, q in zip(a[2] + a[3], b[2] + b[3]))
    print("deterministic", det, "ok, worst |diff| vs the first steps", worst)
ctx.set_option("deterministic", 0)
plan.close()
