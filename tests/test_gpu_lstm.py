"""GPU parity of the LSTM fast path (af_lstm_*) against the oracle's restatement of bench/lstm.go,
lane by lane (a batch of B lanes = B independent reference runs whose parameter gradients add)."""
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
    bad = err > ATOL + RTOL * np.abs(want)
    assert not bad.any(), f"{what}: {bad.sum()} of {bad.size} beyond tolerance; worst {err.max():.3e}"


def _oracle(net, x, up, upr, with_r):
    T, B = x.shape[0], x.shape[1]
    outs = np.zeros((T, B, net.O))
    routs = np.zeros((T, B, net.O))
    grad = {v: np.zeros_like(v.vector) for v in net.variables}
    rgrad = {v: np.zeros_like(v.vector) for v in net.variables}
    for b in range(B):
        xs, us = [x[t, b] for t in range(T)], [up[t, b] for t in range(T)]
        if with_r:
            urs = None if upr is None else [upr[t, b] for t in range(T)]
            o, ro, g, rg = net.run_r(xs, us, urs)
            routs[:, b] = np.array(ro)
            for v in net.variables:
                rgrad[v] += rg[v]
        else:
            o, g = net.run(xs, us)
        outs[:, b] = np.array(o)
        for v in net.variables:
            grad[v] += g[v]
    return outs, routs, grad, rgrad


@pytest.mark.parametrize("I,H,O,T,B,scale", [(30, 100, 10, 5, 3, 1.0),    # bench/lstm.go:13-18 sizes, short
                                            (6, 8, 4, 4, 2, 0.5), (3, 5, 2, 3, 2, 0.5), (16, 64, 10, 6, 70, 0.2)])
@pytest.mark.parametrize("with_r", [True, False])
def test_lstm_matches_oracle(api, I, H, O, T, B, scale, with_r):
    from autofunc_b200.lstm import LSTMPlan
    net = ref.LSTMNet(I, H, O, weight_scale=scale)
    x = ref.splitmix_uniform(501, T * B * I, 0, 1).reshape(T, B, I)      # lstm.go:56 inputs U[0,1)
    up = ref.splitmix_uniform(502, T * B * O, 0, 1).reshape(T, B, O)     # lstm.go:67 out-grads U[0,1)
    upr = ref.splitmix_uniform(503, T * B * O).reshape(T, B, O) if with_r else None
    outs, routs, grad, rgrad = _oracle(net, x, up, upr, with_r)
    plan = LSTMPlan(I, H, O, T, B)
    for i, v in enumerate(net.variables):
        plan.set_param(i, v.vector)
        plan.set_rvector(i, net.rvec[v])
    out, rout, grads, rgrads = plan.run_host(x, up, upr, with_r=with_r)
    close(out.reshape(T, B, O), outs, "outputs")
    if with_r:
        close(rout.reshape(T, B, O), routs, "R-outputs")
    for i, v in enumerate(net.variables):
        close(grads[i], grad[v], f"grad[{i}]")
        if with_r:
            close(rgrads[i], rgrad[v], f"rgrad[{i}]")
    # second run captures the time loops into CUDA graphs, third replays them: same results
    for rep in (2, 3):
        out2, rout2, grads2, rgrads2 = plan.run_host(x, up, upr, with_r=with_r)
        close(out2.reshape(T, B, O), outs, f"outputs (run {rep})")
        for i, v in enumerate(net.variables):
            close(grads2[i], grad[v], f"grad[{i}] (run {rep})")
            if with_r:
                close(rgrads2[i], rgrad[v], f"rgrad[{i}] (run {rep})")
    plan.close()


def test_lstm_full_size_properties(api):
    """BASELINE config 4 (H=512, T=128, B=256; I=30, O=10): R-linearity and Apply/ApplyR consistency."""
    from autofunc_b200.lstm import LSTMPlan
    I, H, O, T, B = 30, 512, 10, 128, 256
    plan = LSTMPlan(I, H, O, T, B)
    rng = [ref.splitmix_uniform(900 + i, plan.param_len(i)) * (0.05 if i % 2 == 0 else 1.0) for i in range(10)]
    rv = [ref.splitmix_uniform(950 + i, plan.param_len(i)) for i in range(10)]
    for i in range(10):
        plan.set_param(i, rng[i])
        plan.set_rvector(i, rv[i])
    x = ref.splitmix_uniform(11, T * B * I, 0, 1)
    up = ref.splitmix_uniform(12, T * B * O, 0, 1)
    out, rout, g, rg = plan.run_host(x, up)
    out0, _, g0, _ = plan.run_host(x, up, with_r=False)
    close(out0, out, "Apply vs ApplyR outputs")
    for i in range(10):
        close(g0[i], g[i], f"Apply vs ApplyR grad[{i}]")
    for i in range(10):
        plan.set_rvector(i, 2 * rv[i])
    out2, rout2, g2, rg2 = plan.run_host(x, up)
    close(rout2, 2 * rout, "R-linearity outputs")
    for i in range(10):
        close(rg2[i], 2 * rg[i], f"R-linearity rgrad[{i}]")
    assert np.isfinite(rout).all() and all(np.isfinite(a).all() for a in rg)
    plan.close()
