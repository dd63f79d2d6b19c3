"""Hessian row-magnitude estimation with the R-operator (the use the reference's README names:
"approximating ... the Hessian's rows' magnitudes", README.md:5, after Dauphin et al., arXiv:1502.04390).

For a scalar objective f(theta) = <out_grad, net(x; theta)> summed over the batch, one R-step with the
R-vector v returns the exact Hessian-vector product Hv in the R-gradient (bench/mlp.go:52-54 is exactly
one such step).  With probes v ~ N(0, I), E[(Hv)_i^2] = ||H_i||^2, so

    D_i = sqrt( mean_p (H v_p)_i^2 )

estimates the magnitude of row i of the Hessian.  Everything stays in HBM: per probe the R-vectors are
uploaded, the plan steps, and the squared R-gradients are accumulated with af_mul / af_axpy."""
from __future__ import annotations

import numpy as np

from .api import MLPPlan, check


def probe_vectors(plan: MLPPlan, seed: int, probe: int):
    """Deterministic N(0,1) probe for every parameter (host side; the same stream feeds the oracle in tests)."""
    rng = np.random.default_rng([seed, probe])
    return [rng.standard_normal(plan.param_len(i)) for i in range(2 * plan.L)]


def hessian_row_magnitudes(plan: MLPPlan, x, probes: int, seed: int = 0, out_grad=None):
    """Returns one array per parameter (bench/mlp.go order W0, b0, W1, b1, ...) with D = sqrt(mean (Hv)^2)."""
    ctx = plan.ctx
    L, h = ctx.lib, ctx.h
    P = 2 * plan.L
    x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
    check(L.af_upload(h, plan.input().ptr, x.ctypes.data, x.size))
    n_out = plan.batch * plan.sizes[-1]
    og = None if out_grad is None else ctx.upload(out_grad)
    acc = [ctx.zeros(plan.param_len(i)) for i in range(P)]
    tmp = [ctx.empty(plan.param_len(i)) for i in range(P)]
    for p in range(probes):
        for i, v in enumerate(probe_vectors(plan, seed, p)):
            plan.set_rvector(i, v)
        plan.zero_grads()
        if og is None:
            check(L.af_fill(h, plan.upstream().ptr, 1.0, n_out))
        else:
            check(L.af_copy(h, plan.upstream().ptr, og.ptr, n_out))
        check(L.af_zero(h, plan.upstream_r().ptr, n_out))
        plan.step(True, True)
        for i in range(P):
            rg = plan.rgrad(i)
            check(L.af_mul(h, rg.ptr, rg.ptr, tmp[i].ptr, rg.n))
            check(L.af_axpy(h, 1.0 / probes, tmp[i].ptr, acc[i].ptr, rg.n))
    return [np.sqrt(a.numpy()) for a in acc]
