"""GPU parity at the sizes BASELINE.json names and round 1 never compared with the oracle:

  C4  bench/lstm.go's LSTM at H = 512, T = 128 (I = 30, O = 10 as the reference has them, and the aligned
      I = O = 512 variant): the 2048 x 542 stacked gate matrix on the split / stream-K contraction path, 128 dependent
      steps in each direction -- every lane against the oracle (bench/lstm.go:139-234 restated one lane at a time), and
      the full B = 256 lane batch against the oracle by lane-independence (outputs) and lane-additivity (parameter
      gradients: the reference's gradient is a sum over lanes, seqfunc/map_batcher.go:26-34).
  C5  an 8192 -> 8192 dense layer R-step (the longest contraction on the path, K = 8192: largest reorder error) and
      the 4 x (8192 -> 8192) plan, at batches the oracle finishes in seconds.

Tolerance: |gpu - oracle| <= 1e-12 + 1e-10 |oracle| (BASELINE.json north_star)."""
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


def close(gpu, want, what, scale=1.0):
    gpu, want = np.asarray(gpu), np.asarray(want) * scale
    assert gpu.shape == want.shape, what
    err = np.abs(gpu - want)
    bad = err > scale * ATOL + RTOL * np.abs(want)
    assert not bad.any(), f"{what}: {bad.sum()} of {bad.size} beyond tolerance; worst {err.max():.3e}"


def _oracle_lanes(net, x, up, upr):
    """The reference runs ONE sequence at a time (bench/lstm.go:41-48); B lanes = B runs whose gradients add."""
    T, B = x.shape[0], x.shape[1]
    outs, routs = np.zeros((T, B, net.O)), np.zeros((T, B, net.O))
    grad = {v: np.zeros_like(v.vector) for v in net.variables}
    rgrad = {v: np.zeros_like(v.vector) for v in net.variables}
    for b in range(B):
        o, ro, g, rg = net.run_r([x[t, b] for t in range(T)], [up[t, b] for t in range(T)],
                                 None if upr is None else [upr[t, b] for t in range(T)])
        outs[:, b], routs[:, b] = np.array(o), np.array(ro)
        for v in net.variables:
            grad[v] += g[v]
            rgrad[v] += rg[v]
    return outs, routs, grad, rgrad


def _plan_for(net, I, H, O, T, B):
    from autofunc_b200.lstm import LSTMPlan
    plan = LSTMPlan(I, H, O, T, B)
    for i, v in enumerate(net.variables):
        plan.set_param(i, v.vector)
        plan.set_rvector(i, net.rvec[v])
    return plan


# 0.05 / 0.03 keep the gates in their active range, so every term of the R-backward matters and the map is well conditioned
@pytest.mark.parametrize("I,O,scale,lanes", [(30, 10, 0.05, 8), (512, 512, 0.03, 4)])
def test_c4_lstm_h512_t128_matches_oracle(api, I, O, scale, lanes):
    H, T, B = 512, 128, lanes
    net = ref.LSTMNet(I, H, O, weight_scale=scale)
    x = ref.splitmix_uniform(601, T * B * I, 0, 1).reshape(T, B, I)      # lstm.go:56 inputs U[0,1)
    up = ref.splitmix_uniform(602, T * B * O, 0, 1).reshape(T, B, O)     # lstm.go:67 out-grads U[0,1)
    upr = ref.splitmix_uniform(603, T * B * O).reshape(T, B, O)
    outs, routs, grad, rgrad = _oracle_lanes(net, x, up, upr)
    plan = _plan_for(net, I, H, O, T, B)
    for rep in range(3):  # eager, graph capture, graph replay
        out, rout, grads, rgrads = plan.run_host(x, up, upr)
        close(out.reshape(T, B, O), outs, f"outputs (run {rep})")
        close(rout.reshape(T, B, O), routs, f"R-outputs (run {rep})")
        for i, v in enumerate(net.variables):
            close(grads[i], grad[v], f"grad[{i}] (run {rep})")
            close(rgrads[i], rgrad[v], f"rgrad[{i}] (run {rep})")
    plan.close()


def test_c4_lstm_reference_init_is_chaotic_and_matches_within_its_conditioning(api):
    """bench/lstm.go:108,113 exactly as written -- U[-1,1) weights -- at H = 512: the pre-activations have a standard
    deviation of ~8, most gates saturate and the few that do not make the recurrence CHAOTIC: perturbing the oracle's own
    weights by one unit in the last place moves the oracle's outputs by up to 4e-3 and its R-outputs by 1e10 after 128
    steps (measured; the first step at which the 1-ulp twin leaves the 1e-10 band is t ~ 35).  No two correct float64
    implementations with different summation orders (gonum's blocked Dgemm vs OpenBLAS vs DMMA tiles) agree to 1e-10
    there, the reference with itself included.  So this configuration is pinned two ways:
      (a) strictly, at 1e-12 + 1e-10 |ref|, on the steps before the oracle's own 1-ulp twin leaves that band;
      (b) on all 128 steps within the map's conditioning: |gpu - ref| <= tol + 1024 max|ref_twin - ref|, the maximum taken
          over the steps so far (outputs; the divergence grows along the sequence) or over the vector (gradients)."""
    I, H, O, T, B = 30, 512, 10, 128, 4
    x = ref.splitmix_uniform(601, T * B * I, 0, 1).reshape(T, B, I)
    up = ref.splitmix_uniform(602, T * B * O, 0, 1).reshape(T, B, O)
    net = ref.LSTMNet(I, H, O, weight_scale=1.0)
    twin = ref.LSTMNet(I, H, O, weight_scale=1.0)
    for v in twin.variables:  # one unit in the last place, random sign
        v.vector *= 1 + 2.2e-16 * np.sign(ref.splitmix_uniform(9, len(v.vector)))
    with np.errstate(over="ignore"):  # exp(-z) overflows to inf for saturated gates: sigmoid = 0, as in Go
        outs, routs, grad, rgrad = _oracle_lanes(net, x, up, None)
        outs2, routs2, grad2, rgrad2 = _oracle_lanes(twin, x, up, None)
    plan = _plan_for(net, I, H, O, T, B)
    out, rout, grads, rgrads = plan.run_host(x, up, None)
    out, rout = out.reshape(T, B, O), rout.reshape(T, B, O)
    plan.close()
    tol = lambda want: ATOL + RTOL * np.abs(want)
    sens = np.abs(outs2 - outs) > tol(outs)
    rsens = np.abs(routs2 - routs) > tol(routs)
    t_ok = int(min(np.argmax(sens.any(axis=(1, 2))) if sens.any() else T, np.argmax(rsens.any(axis=(1, 2))) if rsens.any() else T))
    assert t_ok >= 8, f"the oracle's 1-ulp twin already diverges at step {t_ok}"
    close(out[:t_ok], outs[:t_ok], f"outputs, steps < {t_ok}")
    close(rout[:t_ok], routs[:t_ok], f"R-outputs, steps < {t_ok}")

    def within_conditioning(got, want, twin_val, what, by_step=False):
        err = np.abs(np.asarray(got) - want)
        d = np.abs(twin_val - want)
        spread = np.maximum.accumulate(d.max(axis=(1, 2)))[:, None, None] if by_step else d.max()
        bad = err > tol(want) + 1024 * spread
        assert not bad.any(), f"{what}: {bad.sum()} of {bad.size} beyond the map's conditioning; worst {err.max():.3e}"
    within_conditioning(out, outs, outs2, "outputs, all steps", by_step=True)
    within_conditioning(rout, routs, routs2, "R-outputs, all steps", by_step=True)
    for i, v in enumerate(net.variables):
        within_conditioning(grads[i], grad[v], grad2[twin.variables[i]], f"grad[{i}]")
        within_conditioning(rgrads[i], rgrad[v], rgrad2[twin.variables[i]], f"rgrad[{i}]")
    assert np.isfinite(out).all() and np.isfinite(rout).all()


def test_c4_lstm_full_batch_256_lanes_against_oracle_lanes(api):
    """BASELINE config 4 exactly (H = 512, T = 128, B = 256, I = 30, O = 10).  The oracle runs 8 distinct lanes; the
    product runs all 256, lane b carrying the inputs of oracle lane b mod 8.  Lanes are independent through the whole
    recurrence, so every product lane must reproduce its oracle lane, and the parameter gradients -- sums over lanes --
    must be 32 x the oracle's 8-lane gradients."""
    I, H, O, T, B, K = 30, 512, 10, 128, 256, 8
    net = ref.LSTMNet(I, H, O, weight_scale=0.05)
    x8 = ref.splitmix_uniform(611, T * K * I, 0, 1).reshape(T, K, I)
    up8 = ref.splitmix_uniform(612, T * K * O, 0, 1).reshape(T, K, O)
    upr8 = ref.splitmix_uniform(613, T * K * O).reshape(T, K, O)
    outs, routs, grad, rgrad = _oracle_lanes(net, x8, up8, upr8)
    rep = B // K
    tile = lambda a: np.ascontiguousarray(np.tile(a, (1, rep, 1)))  # lane b <- lane b mod 8
    plan = _plan_for(net, I, H, O, T, B)
    out, rout, grads, rgrads = plan.run_host(tile(x8), tile(up8), tile(upr8))
    close(out.reshape(T, B, O), tile(outs), "outputs, 256 lanes")
    close(rout.reshape(T, B, O), tile(routs), "R-outputs, 256 lanes")
    for i, v in enumerate(net.variables):
        close(grads[i], grad[v], f"grad[{i}] = 32 x oracle", scale=rep)
        close(rgrads[i], rgrad[v], f"rgrad[{i}] = 32 x oracle", scale=rep)
    plan.close()


def test_c5_dense_layer_8192_matches_oracle(api):
    """One 8192 -> 8192 dense layer (LinTran -> LinAdd -> Sigmoid), ApplyR + PropagateRGradient with the input in both
    maps, so all six LinTran kernels G1-G6 run at K = 8192 together with the bias and sigmoid forms: through the
    operator API (unfused: lin_tran.go / lin_add.go / math_funcs.go one kernel each) and fused (DenseSigmoid)."""
    M = K = 8192
    n = 512
    sc = 1 / np.sqrt(K)
    ow, ob = ref.Variable(ref.splitmix_uniform(7001, M * K) * sc), ref.Variable(ref.splitmix_uniform(7002, M))
    ox = ref.Variable(ref.splitmix_uniform(7003, n * K))
    orv = {ow: ref.splitmix_uniform(7004, M * K), ob: ref.splitmix_uniform(7005, M), ox: ref.splitmix_uniform(7006, n * K)}
    up, upr = ref.splitmix_uniform(7007, n * M), ref.splitmix_uniform(7008, n * M)
    olayers = ref.ComposedRBatcher([ref.LinTran(ow, M, K), ref.LinAdd(ob), ref.Sigmoid()])
    ores = olayers.batch_r(orv, ref.RVariable(ox, orv), n)
    ograd, orgrad = ref.new_gradient([ow, ob, ox]), ref.new_rgradient([ow, ob, ox])
    ores.propagate_rgradient(up.copy(), upr.copy(), orgrad, ograd)

    pw, pb, px = api.Variable(ow.vector), api.Variable(ob.vector), api.Variable(ox.vector)
    prv = {pw: orv[ow], pb: orv[ob], px: orv[ox]}
    for fused in (False, True):
        layers = api.ComposedRBatcher([api.LinTran(pw, M, K), api.LinAdd(pb), api.Sigmoid()])
        if fused:
            layers = api.fuse_dense(layers)
        res = layers.batch_r(prv, api.RVariable(px, prv), n)
        grad, rgrad = api.new_gradient([pw, pb, px]), api.new_rgradient([pw, pb, px])
        res.propagate_rgradient(up.copy(), upr.copy(), rgrad, grad)
        tag = "fused" if fused else "unfused"
        close(res.output(), ores.output(), f"Output ({tag})")
        close(res.routput(), ores.routput(), f"ROutput ({tag})")
        for name, pv, ov in (("W", pw, ow), ("b", pb, ob), ("x", px, ox)):
            close(grad[pv], ograd[ov], f"grad[{name}] ({tag})")
            close(rgrad[pv], orgrad[ov], f"rgrad[{name}] ({tag})")
        del res, grad, rgrad
    api.context().trim()


def test_c5_plan_4x8192_matches_oracle(api):
    """BASELINE config 5's network, 4 x (8192 -> 8192), as the af_mlp plan runs it (one data-parallel rank's step):
    n = 256 rows against the oracle -- Output, ROutput and all 16 gradient / R-gradient vectors (2.1 GB each way)."""
    sizes, n = [8192] * 5, 256
    onet = ref.MLP(sizes, weight_scale=lambda k: 1 / np.sqrt(k))
    x = onet.make_input(n)
    ores, ograd, orgrad = onet.step_r(x, n)
    plan = api.MLPPlan(sizes, n)
    for i, v in enumerate(onet.variables):
        plan.set_param(i, v.vector)
        plan.set_rvector(i, onet.rvec[v])
    out, rout, grads, rgrads = plan.step_host(x.vector)
    close(out, ores.output(), "Output")
    close(rout, ores.routput(), "ROutput")
    for i, v in enumerate(onet.variables):
        close(grads[i], ograd[v], f"grad[{i}]")
        close(rgrads[i], orgrad[v], f"rgrad[{i}]")
    plan.close()
