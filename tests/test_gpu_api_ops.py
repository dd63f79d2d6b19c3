"""GPU tests of the operator-API rows that sit either side of the contraction kernels: Concat / Slice /
Repeat (slices.go), the LSTM cell assembled from operators (bench/lstm.go:139-151 through the RFunc
interface), and seqfunc.MapBatcher / MapRBatcher (seqfunc/map_batcher.go)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import autofunc_ref as ref
from oracle.functest import RFuncChecker

RTOL, ATOL = 1e-10, 1e-12


@pytest.fixture(scope="module")
def api():
    import torch
    assert torch.cuda.is_available()
    import autofunc_b200  # noqa: F401  (installs slices into api)
    from autofunc_b200 import api as a
    a.context()
    return a


def close(gpu, want, what):
    gpu, want = np.asarray(gpu), np.asarray(want)
    assert gpu.shape == want.shape, what
    err = np.abs(gpu - want)
    assert not (err > ATOL + RTOL * np.abs(want)).any(), f"{what}: worst {err.max():.3e}"


def _slices_func(ns):
    class F:
        def apply(self, x):
            a = ns.slice_(x, 0, 2)
            b = ns.repeat(ns.slice_(x, 2, 3), 2)
            return ns.mul(ns.concat(a, b, ns.slice_(x, 3, 5)), ns.concat(b, ns.slice_(x, 1, 5)))

        def apply_r(self, rv, x):
            a = ns.slice_r(x, 0, 2)
            b = ns.repeat_r(ns.slice_r(x, 2, 3), 2)
            return ns.mul_r(ns.concat_r(a, b, ns.slice_r(x, 3, 5)), ns.concat_r(b, ns.slice_r(x, 1, 5)))
    return F()


def test_slices_fullcheck_oracle_and_product(api):
    """The reference's acceptance procedure on a function built from Concat, Slice and Repeat."""
    for ns in (ref, api):
        x = ns.Variable([0.5, -1.25, 2.0, 0.75, -0.3])
        rv = {x: np.array([1.0, -2.0, 0.5, 3.0, -1.5])}
        assert RFuncChecker(ns, _slices_func(ns), [x], x, rv).full_check() == []


def _lstm_step_ops(ns, p, state, x, n, rv, concat_r):
    """bench/lstm.go:139-151,187-194 from operators; p = [W_i,b_i,W_g,b_g,W_f,b_f,W_o,b_o,W_out,b_out]."""
    H = p[1].vector.size
    C = p[0].vector.size // H
    O = p[9].vector.size
    s = ns.Sigmoid()
    inp = concat_r(n, state, x)
    gate = lambda k: s.batch_r(rv, ns.LinAdd(p[2 * k + 1]).batch_r(rv, ns.LinTran(p[2 * k], H if k < 4 else O, C).batch_r(rv, inp, n), n), n)
    new, mask, forget, og = gate(0), gate(1), gate(2), gate(3)
    out_state = ns.add_r(ns.mul_r(forget, state), ns.mul_r(mask, new))
    masked = ns.mul_r(out_state, og)
    aug = concat_r(n, masked, x)
    res = s.batch_r(rv, ns.LinAdd(p[9]).batch_r(rv, ns.LinTran(p[8], O, C).batch_r(rv, aug, n), n), n)
    return out_state, res


def test_lstm_step_from_operators_matches_plan(api):
    """One LSTM time step (two steps chained) built from the operator API equals the fused LSTM plan."""
    from autofunc_b200.lstm import LSTMPlan
    I, H, O, T, B = 6, 8, 4, 2, 3
    net = ref.LSTMNet(I, H, O, weight_scale=0.5)
    x = ref.splitmix_uniform(501, T * B * I, 0, 1).reshape(T, B, I)
    up = ref.splitmix_uniform(502, T * B * O, 0, 1).reshape(T, B, O)
    plan = LSTMPlan(I, H, O, T, B)
    for i, v in enumerate(net.variables):
        plan.set_param(i, v.vector)
        plan.set_rvector(i, net.rvec[v])
    out, rout, grads, rgrads = plan.run_host(x, up)
    pv = [api.Variable(v.vector) for v in net.variables]
    prv = {a: net.rvec[b] for a, b in zip(pv, net.variables)}
    s0 = api.Variable(np.zeros(B * H))
    state = api.RVariable(s0, prv)
    results = []
    for t in range(T):
        xv = api.RVariable(api.Variable(x[t].reshape(-1)), prv)
        state, res = _lstm_step_ops(api, pv, state, xv, B, prv, api.concat_batched_r)
        results.append(res)
    close(np.concatenate([r.output() for r in results]), out, "outputs")
    close(np.concatenate([r.routput() for r in results]), rout, "R-outputs")
    # full BPTT through the operator graph: sum_t <up_t, res_t> has the plan's gradients
    grad, rgrad = api.new_gradient(pv), api.new_rgradient(pv)
    for t in range(T):
        results[t].propagate_rgradient(up[t].reshape(-1).copy(), np.zeros(B * O), rgrad, grad)
    for i, v in enumerate(pv):
        close(grad[v], grads[i], f"grad[{i}]")
        close(rgrad[v], rgrads[i], f"rgrad[{i}]")
    plan.close()


def test_map_rbatcher_ragged_sequences(api):
    """seqfunc/test/map_test.go:88-120 (TestMapBatcherOutput): mapping a Batcher over ragged sequences equals
    applying it vector by vector; gradients equal the oracle's per-vector sums."""
    from autofunc_b200 import seq
    sizes = [4, 6, 3]
    onet = ref.MLP(sizes, weight_scale=lambda k: 1 / np.sqrt(k))
    pvars = [api.Variable(v.vector) for v in onet.variables]
    prv = {pv: onet.rvec[ov] for pv, ov in zip(pvars, onet.variables)}
    layers = api.fuse_dense([api.LinTran(pvars[0], 6, 4), api.LinAdd(pvars[1]), api.Sigmoid(),
                             api.LinTran(pvars[2], 3, 6), api.LinAdd(pvars[3]), api.Sigmoid()])
    lens = [3, 1, 0, 2]  # ragged lanes, one empty (seqfunc/test/map_test.go:28-34)
    seqs = [[ref.splitmix_uniform(700 + 10 * l + t, 4) for t in range(n)] for l, n in enumerate(lens)]
    res = seq.MapRBatcher(layers).apply_seqs_r(prv, seq.ConstRResult(seqs))
    per = ref.ComposedRFunc(list(onet.layers))
    ograd, orgrad = ref.new_gradient(onet.variables), ref.new_rgradient(onet.variables)
    ups = [[ref.splitmix_uniform(800 + 10 * l + t, 3) for t in range(n)] for l, n in enumerate(lens)]
    uprs = [[ref.splitmix_uniform(850 + 10 * l + t, 3) for t in range(n)] for l, n in enumerate(lens)]
    for l, s in enumerate(seqs):
        assert len(res.output_seqs()[l]) == len(s)
        for t, v in enumerate(s):
            xv = ref.Variable(v)
            r = per.apply_r(onet.rvec, ref.RVariable(xv, onet.rvec))
            close(res.output_seqs()[l][t], r.output(), f"lane {l} t {t}")
            close(res.routput_seqs()[l][t], r.routput(), f"lane {l} t {t} R")
            r.propagate_rgradient(ups[l][t].copy(), uprs[l][t].copy(), orgrad, ograd)
    grad, rgrad = api.new_gradient(pvars), api.new_rgradient(pvars)
    res.propagate_rgradient(ups, uprs, rgrad, grad)
    for i, (pv, ov) in enumerate(zip(pvars, onet.variables)):
        close(grad[pv], ograd[ov], f"grad[{i}]")
        close(rgrad[pv], orgrad[ov], f"rgrad[{i}]")
    assert set(grad) == set(pvars)  # temporaries were deleted again (map_batcher.go:136-137)
