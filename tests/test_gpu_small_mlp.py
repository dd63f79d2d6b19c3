"""The persistent whole-step kernel for n <= 4 samples (csrc/small_mlp.cu: all layers, both directions, grid barriers
between phases -- the reference bench exactly as written is n = 1, bench/mlp.go:36,52) against the oracle, and against
the per-kernel path it replaces (af_set_option "small_mlp" 0)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import autofunc_ref as ref

RTOL, ATOL = 1e-10, 1e-12


@pytest.fixture(scope="module")
def api():
    import torch
    assert torch.cuda.is_available()
    from autofunc_b200 import api as a
    a.context()
    return a


def close(gpu, want, what):
    gpu, want = np.asarray(gpu), np.asarray(want)
    assert gpu.shape == want.shape, what
    err = np.abs(gpu - want)
    assert not (err > ATOL + RTOL * np.abs(want)).any(), f"{what}: worst {err.max():.3e}"


def _plan(api, onet, sizes, n, absent=()):
    plan = api.MLPPlan(sizes, n)
    for i, v in enumerate(onet.variables):
        plan.set_param(i, v.vector)
        plan.set_rvector(i, None if i in absent else onet.rvec[v])
    return plan


@pytest.mark.parametrize("sizes", [[1000, 2000, 512, 10], [6, 8, 4], [64, 2], [34, 50, 18, 22, 6], [512, 600]])
@pytest.mark.parametrize("n", [1, 2, 3, 5, 8])
@pytest.mark.parametrize("with_r", [True, False])
def test_small_mlp_step_matches_oracle(api, sizes, n, with_r):
    onet = ref.MLP(sizes, weight_scale=lambda k: 1 / np.sqrt(k))
    x = onet.make_input(n)
    top = n * sizes[-1]
    up, upr = ref.splitmix_uniform(71, top), ref.splitmix_uniform(72, top)
    if with_r:
        ores, ograd, orgrad = onet.step_r(x, n, out_grad=up, out_rgrad=upr)
    else:
        ores, ograd = onet.step(x, n, out_grad=up)
    plan = _plan(api, onet, sizes, n)
    before = plan.ctx.launch_count()
    out, rout, grads, rgrads = plan.step_host(x.vector, with_r=with_r, upstream=up, upstream_r=upr if with_r else None)
    if n <= 4:
        assert plan.ctx.launch_count() - before == 1, "n <= 4 must run as ONE persistent kernel"
    close(out, ores.output(), "Output")
    if with_r:
        close(rout, ores.routput(), "ROutput")
    for i, v in enumerate(onet.variables):
        close(grads[i], ograd[v], f"grad[{i}]")
        if with_r:
            close(rgrads[i], orgrad[v], f"rgrad[{i}]")
    # repeated resident steps accumulate (bench/mlp.go:33-40); the barrier counter carries over between launches
    ctx = plan.ctx
    for k in (2, 3, 4):
        api.check(ctx.lib.af_upload(ctx.h, plan.upstream().ptr, up.ctypes.data, top))
        api.check(ctx.lib.af_upload(ctx.h, plan.upstream_r().ptr, upr.ctypes.data, top))
        plan.step(with_r=with_r)
        for i in (0, len(onet.variables) - 1):
            close(plan.grad(i).numpy(), k * ograd[onet.variables[i]], f"accumulated grad[{i}] x{k}")
    # one bench iteration in one call: fresh accumulators, outGrad = 1 / outRGrad = 0 (bench/mlp.go:33-35,118-126)
    if with_r:
        wres, wgrad, wrgrad = onet.step_r(x, n)
    else:
        wres, wgrad = onet.step(x, n)
    plan.step_fresh(with_r=with_r)
    close(plan.output().numpy(), wres.output(), "fresh Output")
    for i, v in enumerate(onet.variables):
        close(plan.grad(i).numpy(), wgrad[v], f"fresh grad[{i}]")
        if with_r:
            close(plan.rgrad(i).numpy(), wrgrad[v], f"fresh rgrad[{i}]")
    plan.close()


def test_small_mlp_absent_rvectors_forward_only_and_ab(api):
    sizes, n = [40, 64, 24, 10], 4
    onet = ref.MLP(sizes, weight_scale=lambda k: 1 / np.sqrt(k))
    x = onet.make_input(n)
    absent = (0, 3)  # W1 and b2 are not in the RVector (variable.go:85-91): their R terms vanish
    rv = {v: r for i, (v, r) in enumerate(onet.rvec.items()) if i not in absent}
    res = onet.layers.batch_r(rv, ref.RVariable(x, rv), n)
    g, rg = ref.new_gradient(onet.variables), ref.new_rgradient(onet.variables)
    res.propagate_rgradient(np.ones(n * 10), np.zeros(n * 10), rg, g)
    plan = _plan(api, onet, sizes, n, absent)
    out, rout, grads, rgrads = plan.step_host(x.vector)
    close(out, res.output(), "Output")
    close(rout, res.routput(), "ROutput")
    for i, v in enumerate(onet.variables):
        close(grads[i], g[v], f"grad[{i}]")
        close(rgrads[i], rg[v], f"rgrad[{i}]")
    # forward only (Run(b, false), bench/mlp.go:29-32)
    plan.zero_grads()
    plan.step(with_r=True, backprop=False)
    close(plan.output().numpy(), res.output(), "forward-only Output")
    assert not plan.grad(0).numpy().any()
    # the per-kernel path gives the same numbers
    ctx = plan.ctx
    api.check(ctx.lib.af_set_option(ctx.h, b"small_mlp", 0))
    try:
        before = ctx.launch_count()
        out2, rout2, grads2, rgrads2 = plan.step_host(x.vector)
        assert ctx.launch_count() - before > 1
    finally:
        api.check(ctx.lib.af_set_option(ctx.h, b"small_mlp", 1))
    close(out2, out, "A/B Output")
    for i in range(len(grads)):
        close(grads2[i], grads[i], f"A/B grad[{i}]")
        close(rgrads2[i], rgrads[i], f"A/B rgrad[{i}]")
    plan.close()


def test_odd_widths_fall_back_to_the_per_kernel_path(api):
    """Odd contraction lengths cannot use 16-byte row accesses: the plan takes the generic per-kernel path, same numbers."""
    sizes, n = [7, 9, 5], 2
    onet = ref.MLP(sizes, weight_scale=lambda k: 1 / np.sqrt(k))
    x = onet.make_input(n)
    ores, ograd, orgrad = onet.step_r(x, n)
    plan = _plan(api, onet, sizes, n)
    before = plan.ctx.launch_count()
    out, rout, grads, rgrads = plan.step_host(x.vector)
    assert plan.ctx.launch_count() - before > 1
    close(out, ores.output(), "Output")
    close(rout, ores.routput(), "ROutput")
    for i, v in enumerate(onet.variables):
        close(grads[i], ograd[v], f"grad[{i}]")
        close(rgrads[i], orgrad[v], f"rgrad[{i}]")
    plan.close()
