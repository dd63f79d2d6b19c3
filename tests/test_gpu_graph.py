"""Recorded launch sequences (af_graph_begin / af_graph_end / af_graph_launch): the reference's own benchmark shapes
are a handful of samples (bench/lintran.go:13-17 n = 10, bench/mlp.go:36 n = 1), where a chain of ABI calls is bound by
launch gaps.  A replayed recording must give the eager results, against the oracle at 1e-12 + 1e-10|ref|."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import autofunc_ref as ref

RTOL, ATOL = 1e-10, 1e-12


@pytest.fixture(scope="module")
def api():
    import torch
    assert torch.cuda.is_available()
    import autofunc_b200  # noqa: F401
    from autofunc_b200 import api as a
    a.context()
    return a


def close(gpu, want, what):
    err = np.abs(np.asarray(gpu) - np.asarray(want))
    assert not (err > ATOL + RTOL * np.abs(want)).any(), f"{what}: worst {err.max():.3e}"


@pytest.mark.parametrize("rows,cols,n", [(1024, 2048, 10), (96, 160, 3), (64, 128, 48)])
def test_recorded_lintran_step_replays_to_the_oracle(api, rows, cols, n):
    """bench/lintran.go:83-89 (BatchR + PropagateRGradient, W and X both in the maps) recorded once, replayed twice:
    the accumulators hold twice the oracle's gradients, outputs are the oracle's."""
    ctx = api.context()
    L, h = ctx.lib, ctx.h
    w, x = ref.Variable(ref.splitmix_uniform(1, rows * cols)), ref.Variable(ref.splitmix_uniform(2, cols * n))
    rv = {w: ref.splitmix_uniform(3, rows * cols), x: ref.splitmix_uniform(4, cols * n)}
    up, upr = ref.splitmix_uniform(5, rows * n), ref.splitmix_uniform(6, rows * n)
    res = ref.LinTran(w, rows, cols).batch_r(rv, ref.RVariable(x, rv), n)
    og, org = ref.new_gradient([w, x]), ref.new_rgradient([w, x])
    oy, ory = res.output().copy(), res.routput().copy()
    res.propagate_rgradient(up.copy(), upr.copy(), org, og)

    d = {k: ctx.upload(v) for k, v in dict(W=w.vector, RW=rv[w], X=x.vector, RX=rv[x], U=up, UR=upr).items()}
    Y, RY, D, DR = (ctx.empty(n * rows), ctx.empty(n * rows), ctx.empty(n * cols), ctx.empty(n * cols))
    gW, rgW, gX, rgX = (ctx.zeros(rows * cols), ctx.zeros(rows * cols), ctx.zeros(n * cols), ctx.zeros(n * cols))

    def step():
        api.check(L.af_lintran_fwd_r(h, d["W"].ptr, d["RW"].ptr, d["X"].ptr, d["RX"].ptr, Y.ptr, RY.ptr, rows, cols, n))
        api.check(L.af_lintran_wgrad_r(h, d["U"].ptr, d["UR"].ptr, d["X"].ptr, d["RX"].ptr, gW.ptr, rgW.ptr, rows, cols, n))
        api.check(L.af_lintran_igrad_r(h, d["U"].ptr, d["UR"].ptr, d["W"].ptr, d["RW"].ptr, D.ptr, DR.ptr, rows, cols, n))
        api.check(L.af_axpy(h, 1.0, D.ptr, gX.ptr, n * cols))
        api.check(L.af_axpy(h, 1.0, DR.ptr, rgX.ptr, n * cols))

    step()  # eager warm-up: sizes scratch, sets kernel attributes
    for v in (gW, rgW, gX, rgX, Y, RY):
        api.check(L.af_zero(h, v.ptr, v.n))
    before = ctx.launch_count()
    with ctx.record() as g:
        step()
    assert ctx.launch_count() == before, "recording must not count (or run) launches"
    assert g.launches >= 5
    assert not Y.numpy().any() and not gW.numpy().any(), "recording executed the kernels"
    g.launch()
    g.launch()
    assert ctx.launch_count() == before + 2 * g.launches
    close(Y.numpy(), oy, "Y")
    close(RY.numpy(), ory, "RY")
    for got, want, what in ((gW, og[w], "gW"), (rgW, org[w], "rgW"), (gX, og[x], "gX"), (rgX, org[x], "rgX")):
        close(got.numpy(), 2 * want, what)
    g.close()


def test_recording_refuses_synchronising_calls_and_recovers(api):
    ctx = api.context()
    L, h = ctx.lib, ctx.h
    a, b, o = ctx.upload(np.arange(8.0)), ctx.upload(np.ones(8)), ctx.zeros(8)
    api.check(L.af_add(h, a.ptr, b.ptr, o.ptr, 8))
    with pytest.raises(api.AutofuncError, match="af_graph_begin and af_graph_end"):
        with ctx.record():
            api.check(L.af_add(h, a.ptr, b.ptr, o.ptr, 8))
            ctx.sync()
    # the abandoned recording left the context usable
    with ctx.record() as g:
        api.check(L.af_axpy(h, 2.0, a.ptr, o.ptr, 8))
    g.launch()
    assert o.numpy().tolist() == (np.arange(8.0) + 1 + 2 * np.arange(8.0)).tolist()
    with pytest.raises(api.AutofuncError):
        api.check(L.af_graph_end(h, C.byref(C.c_void_p())))  # not recording
    plan = api.MLPPlan([8, 4], 2)
    with pytest.raises(api.AutofuncError, match="graphs of their own"):
        with ctx.record():
            plan.step(True, True)
    plan.close()
    g.close()
