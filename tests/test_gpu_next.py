"""GPU parity tests of the SURVEY 8(f) rows: the reference's own tests for the functions, cost heads and graph
combinators either side of the dense path (test/arithmetic_test.go, test/math_funcs_test.go, test/pool_test.go,
test/fold_test.go, test/lin_alg_test.go), run against the product through the C ABI, plus direct comparisons with
the oracle at 1e-12 + 1e-10|ref| on larger seeded inputs and the device-resident Gradient / SGD step."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import autofunc_ref as ref
from oracle.functest import RFuncChecker
from tests import fixtures as fx

RTOL, ATOL = 1e-10, 1e-12


@pytest.fixture(scope="module")
def api():
    import torch
    assert torch.cuda.is_available()
    import autofunc_b200  # noqa: F401
    from autofunc_b200 import api as a
    a.context()
    return a


def close(gpu, want, what, rtol=RTOL, atol=ATOL):
    gpu, want = np.asarray(gpu), np.asarray(want)
    assert gpu.shape == want.shape, f"{what}: shape {gpu.shape} vs {want.shape}"
    err = np.abs(gpu - want)
    bad = err > atol + rtol * np.abs(want)
    assert not bad.any(), f"{what}: worst {err.max():.3e} at {int(np.argmax(err))}"


# ---- the reference's acceptance tests, against the product -----------------------------------------
def test_arithmetic_full_check(api):  # test/arithmetic_test.go:69-77
    f = fx.ArithmeticFixture(api)
    assert RFuncChecker(api, f.func, f.variables, f.input, f.rv).full_check() == []


def test_log_domain_outputs(api):  # test/arithmetic_test.go:79-107
    a, b = api.Variable(fx.ARITH_VEC1), api.Variable(fx.ARITH_VEC2)
    np.testing.assert_allclose(api.add_log_domain(a, b).output(), np.log(np.exp(a.vector) + np.exp(b.vector)), atol=1e-5)
    v = api.Variable([3, 3.5, -1, 2])
    assert abs(api.sum_all_log_domain(v).output()[0] - np.log(np.sum(np.exp(v.vector)))) < 1e-8


@pytest.mark.parametrize("name,positive", [("Log", True), ("LogSigmoid3", False), ("Softmax", False),
                                           ("Softmax3", False), ("Sin", False), ("Norm", False), ("Tanh", False)])
def test_next_math_func_full_check(api, name, positive):  # test/math_funcs_test.go:32-130
    make = {
        "Log": lambda ns: ns.Log(), "Sin": lambda ns: ns.Sin(), "Norm": lambda ns: ns.Norm(), "Tanh": lambda ns: ns.Tanh(),
        "LogSigmoid3": lambda ns: ns.ComposedRFunc([ns.LogSigmoid(), ns.LogSigmoid(), ns.LogSigmoid()]),
        "Softmax": lambda ns: ns.Softmax(), "Softmax3": lambda ns: ns.Softmax(3),
    }[name]
    v = api.Variable(fx.MATH_VEC)
    inp = api.Variable(fx.MATH_VEC_POS) if positive else v
    rv = {v: np.array(fx.MATH_RV)}
    assert RFuncChecker(api, make(api), [v], inp, rv).full_check() == []


def test_cos_and_norm_outputs(api):  # test/math_funcs_test.go:102-120
    for arg in ref.splitmix_uniform(5, 10, -10, 10):
        assert abs(api.Cos().apply(api.Variable([arg])).output()[0] - np.cos(arg)) < 1e-5
    assert abs(api.Norm().apply(api.Variable([3, 4])).output()[0] - 5.0) < 1e-5


def test_pool_full_check(api):  # test/pool_test.go:29-38
    v = api.Variable(fx.POOL_VEC)
    assert RFuncChecker(api, fx.pool_test_func(api), [v], v, {v: np.array(fx.POOL_RV, dtype=float)}).full_check() == []


def test_fold_output_and_full_check(api):  # test/fold_test.go:32-65
    v0, v1 = api.Variable(fx.FOLD_VEC), api.Variable(fx.FOLD_STATE)
    f = fx.fold_test_func(api, v1)
    np.testing.assert_allclose(f.apply(v0).output(), fx.FOLD_KAT_OUT, atol=1e-5)
    rng = np.random.default_rng(0)
    rv = {v0: rng.standard_normal(6), v1: rng.standard_normal(2)}
    assert RFuncChecker(api, f, [v0, v1], v0, rv).full_check() == []


def test_matmulvec_output_and_full_check(api):  # test/lin_alg_test.go:55-73
    f = fx.LinTranFixture(api)
    m = fx.matmulvec_test_func(api, f.mat1.data)
    assert np.array_equal(m.apply(f.vec).output(), fx.LINTRAN_KAT_OUT)
    assert RFuncChecker(api, m, f.variables, f.vec, f.rv).full_check() == []


# ---- direct parity with the oracle on larger seeded inputs -------------------------------------------
def _pair(api, seed, n, lo=-1.0, hi=1.0, with_r=True):
    x = ref.splitmix_uniform(seed, n, lo, hi)
    xr = ref.splitmix_uniform(seed + 1000, n)
    ov, pv = ref.Variable(x), api.Variable(x)
    return (ov, {ov: xr} if with_r else {}), (pv, {pv: xr.copy()} if with_r else {})


def _run_r(ns, f, v, rv, up, up_r):
    res = f.apply_r(rv, ns.RVariable(v, rv))
    g, rg = ns.new_gradient([v]), ns.new_rgradient([v])
    res.propagate_rgradient(up.copy(), up_r.copy(), rg, g)
    return np.array(res.output()), np.array(res.routput()), np.array(g[v]), np.array(rg[v])


def _run(ns, f, v, up):
    res = f.apply(v)
    g = ns.new_gradient([v])
    res.propagate_gradient(up.copy(), g)
    return np.array(res.output()), np.array(g[v])


UNARY = {
    "Log": (lambda ns: ns.Log(), 0.1, 3.0),
    "LogSigmoid": (lambda ns: ns.LogSigmoid(), -30.0, 30.0),
    "Sin": (lambda ns: ns.Sin(), -10.0, 10.0),
    "Cos": (lambda ns: ns.Cos(), -10.0, 10.0),
    "Tanh": (lambda ns: ns.Tanh(), -4.0, 4.0),
    "Inverse": (lambda ns: _Wrap(ns, "inverse"), 0.2, 3.0),
    "Pow2.5": (lambda ns: _Wrap(ns, "pow_", 2.5), 0.2, 3.0),
    "Pow1": (lambda ns: _Wrap(ns, "pow_", 1.0), 0.2, 3.0),
    "Pow0": (lambda ns: _Wrap(ns, "pow_", 0.0), 0.2, 3.0),
    "AddScaler": (lambda ns: _Wrap(ns, "add_scaler", -1.75), -1.0, 1.0),
}


class _Wrap:
    def __init__(self, ns, name, *args):
        self.ns, self.name, self.args = ns, name, args

    def apply(self, r):
        return getattr(self.ns, self.name)(r, *self.args)

    def apply_r(self, rv, r):
        return getattr(self.ns, self.name + ("r" if self.name.endswith("_") else "_r"))(r, *self.args)


@pytest.mark.parametrize("name", sorted(UNARY))
@pytest.mark.parametrize("n", [1, 1037, 70001])
def test_unary_functions_match_oracle(api, name, n):
    make, lo, hi = UNARY[name]
    (ov, orv), (pv, prv) = _pair(api, 11, n, lo, hi)
    up, up_r = ref.splitmix_uniform(5, n), ref.splitmix_uniform(6, n)
    want = _run_r(ref, make(ref), ov, orv, up, up_r)
    got = _run_r(api, make(api), pv, prv, up, up_r)
    for w, g, what in zip(want, got, ("output", "routput", "grad", "rgrad")):
        close(g, w, f"{name} n={n} {what}")
    want = _run(ref, make(ref), ov, up)
    got = _run(api, make(api), pv, up)
    for w, g, what in zip(want, got, ("output", "grad")):
        close(g, w, f"{name} n={n} non-R {what}")
    # variable absent from the RVector (variable.go:85-91): zero R input, kernels get a NULL pointer
    want = _run_r(ref, make(ref), ov, {}, up, up_r)
    got = _run_r(api, make(api), pv, {}, up, up_r)
    for w, g, what in zip(want, got, ("output", "routput", "grad", "rgrad")):
        close(g, w, f"{name} n={n} absent-R {what}")


@pytest.mark.parametrize("rows,length,temp", [(1, 7, 0.0), (37, 10, 0.0), (4096, 10, 3.0), (5, 1024, 1.0), (3, 3001, 0.5),
                                              (300, 257, 2.0), (50, 24, 0.0), (45, 32, 1.5), (20, 200, 0.0),
                                              (9, 256, 0.0), (33, 16, 0.0),
                                              # rows of <= 16 entries: one thread per row (softmax_*_short_kernel)
                                              (1000, 2, 0.0), (300, 1, 2.0), (777, 13, 0.5), (5000, 16, 0.0), (187, 10, 0.0)])
def test_batched_softmax_matches_oracle(api, rows, length, temp):
    """One fused launch per direction vs the reference's chain under RFuncBatcher (batcher.go:57-84)."""
    n = rows * length
    (ov, orv), (pv, prv) = _pair(api, 21, n, -3.0, 3.0)
    up, up_r = ref.splitmix_uniform(7, n), ref.splitmix_uniform(8, n)

    class B:
        def __init__(self, ns):
            self.ns = ns

        def apply(self, r):
            return self.ns.Softmax(temp).batch(r, rows)

        def apply_r(self, rv, r):
            return self.ns.Softmax(temp).batch_r(rv, r, rows)
    want = _run_r(ref, B(ref), ov, orv, up, up_r)
    got = _run_r(api, B(api), pv, prv, up, up_r)
    for w, g, what in zip(want, got, ("output", "routput", "grad", "rgrad")):
        close(g, w, f"softmax {rows}x{length} T={temp} {what}")
    np.testing.assert_allclose(got[0].reshape(rows, length).sum(axis=1), 1.0, rtol=1e-12)  # size-independent property
    want = _run(ref, B(ref), ov, up)
    got = _run(api, B(api), pv, up)
    for w, g, what in zip(want, got, ("output", "grad")):
        close(g, w, f"softmax {rows}x{length} non-R {what}")


def test_softmax_shape_error(api):
    with pytest.raises(ValueError, match="input count does not divide input length"):
        api.Softmax().batch(api.Variable(np.ones(10)), 3)


@pytest.mark.parametrize("n", [1, 10, 40960, 1 << 20])
def test_squared_error_head_matches_oracle(api, n):
    a, ar = ref.splitmix_uniform(31, n), ref.splitmix_uniform(32, n)
    y, yr = ref.splitmix_uniform(33, n), ref.splitmix_uniform(34, n)
    out = {}
    for ns in (ref, api):
        va, vy = ns.Variable(a), ns.Variable(y)
        rv = {va: ar.copy(), vy: yr.copy()}
        res = ns.squared_error_r(ns.RVariable(va, rv), ns.RVariable(vy, rv))
        g, rg = ns.new_gradient([va, vy]), ns.new_rgradient([va, vy])
        res.propagate_rgradient(np.array([0.7]), np.array([-1.3]), rg, g)
        res2 = ns.squared_error(va, vy)
        g2 = ns.new_gradient([va])
        res2.propagate_gradient(np.array([2.0]), g2)
        out[ns] = [np.array(res.output()), np.array(res.routput()), g[va], g[vy], rg[va], rg[vy],
                   np.array(res2.output()), g2[va]]
    # the cost is a length-n sum: relative to the oracle's sequential sum the tree order moves it by ~n*eps
    for w, g, what in zip(out[ref], out[api], ("cost", "r-cost", "g_a", "g_y", "rg_a", "rg_y", "cost non-R", "g_a non-R")):
        close(g, w, f"squared_error n={n} {what}", atol=1e-12 + 1e-16 * n)


def test_mlp_with_softmax_and_cost_heads_matches_oracle(api):
    """Dense layers -> batched softmax -> squared error, R-step: the callers' side of the path end to end."""
    sizes, n = [24, 40, 10], 33
    onet = ref.MLP(sizes, weight_scale=lambda k: 1 / np.sqrt(k))
    x = onet.make_input(n)
    target = ref.splitmix_uniform(77, n * sizes[-1], 0.0, 1.0)
    res = {}
    for ns in (ref, api):
        variables = [ns.Variable(v.vector) for v in onet.variables]
        rv = {nv: onet.rvec[ov].copy() for nv, ov in zip(variables, onet.variables)}
        layers = []
        for l in range(len(sizes) - 1):
            layers += [ns.LinTran(variables[2 * l], sizes[l + 1], sizes[l]), ns.LinAdd(variables[2 * l + 1]), ns.Sigmoid()]
        net = ns.ComposedRBatcher(layers[:-1] + [ns.Softmax(2.0)])
        xv, tv = ns.Variable(x.vector), ns.Variable(target)
        out = net.batch_r(rv, ns.RVariable(xv, rv), n)
        cost = ns.squared_error_r(out, ns.RVariable(tv, rv))
        g, rg = ns.new_gradient(variables), ns.new_rgradient(variables)
        cost.propagate_rgradient(np.array([1.0]), np.array([0.0]), rg, g)
        res[ns] = ([np.array(cost.output()), np.array(cost.routput())] + [g[v] for v in variables] + [rg[v] for v in variables])
    for i, (w, g) in enumerate(zip(res[ref], res[api])):
        close(g, w, f"softmax+cost head item {i}")


def test_fold_long_recurrence_matches_oracle(api):
    """Fold over 64 timesteps of a dense recurrent cell: state stays in HBM, BPTT through the temp-variable protocol."""
    H, I, T = 48, 16, 64
    w = ref.splitmix_uniform(41, H * (H + I)) / np.sqrt(H + I)
    wr = ref.splitmix_uniform(42, H * (H + I))
    xs = ref.splitmix_uniform(43, T * I)
    s0 = ref.splitmix_uniform(44, H)
    out = {}
    for ns in (ref, api):
        W, X, S = ns.Variable(w), ns.Variable(xs), ns.Variable(s0)
        rv = {W: wr.copy(), S: ref.splitmix_uniform(45, H)}
        lt = ns.LinTran(W, H, H + I)
        sig = ns.Sigmoid()
        ins = ns.split_r(T, ns.RVariable(X, rv))
        res = ns.fold_r(ns.RVariable(S, rv), ins, lambda s, i: sig.apply_r(rv, lt.apply_r(rv, ns.concat_r(s, i))))
        g, rg = ns.new_gradient([W, X, S]), ns.new_rgradient([W, X, S])
        res.propagate_rgradient(np.ones(H), np.zeros(H), rg, g)
        out[ns] = [np.array(res.output()), np.array(res.routput())] + [g[v] for v in (W, X, S)] + [rg[v] for v in (W, X, S)]
    for i, (w_, g_) in enumerate(zip(out[ref], out[api])):
        close(g_, w_, f"fold item {i}")


def test_matmulvecs_computed_matrix_matches_oracle(api):
    """MatMulVecs with a matrix that is itself a function of variables (lin_alg.go:28-49): the gradient flows
    into the matrix through the temp-variable protocol; sizes large enough for the TMA + DMMA kernels."""
    rows, cols, n = 96, 64, 40
    a = ref.splitmix_uniform(51, rows * cols) / 8
    b = ref.splitmix_uniform(52, rows * cols) / 8
    x = ref.splitmix_uniform(53, cols * n)
    out = {}
    for ns in (ref, api):
        A, B, X = ns.Variable(a), ns.Variable(b), ns.Variable(x)
        rv = {A: ref.splitmix_uniform(54, rows * cols), X: ref.splitmix_uniform(55, cols * n)}
        ra, rb, rx = (ns.RVariable(v, rv) for v in (A, B, X))
        res = ns.mat_mul_vecs_r(ns.mul_r(ra, ns.add_r(ra, rb)), rows, cols, rx)
        g, rg = ns.new_gradient([A, B, X]), ns.new_rgradient([A, B, X])
        up, up_r = ref.splitmix_uniform(56, rows * n), ref.splitmix_uniform(57, rows * n)
        res.propagate_rgradient(up, up_r, rg, g)
        res2 = ns.mat_mul_vecs(ns.mul(A, ns.add(A, B)), rows, cols, X)
        g2 = ns.new_gradient([A, X])
        res2.propagate_gradient(ref.splitmix_uniform(56, rows * n), g2)
        out[ns] = ([np.array(res.output()), np.array(res.routput())] + [g[v] for v in (A, B, X)]
                   + [rg[v] for v in (A, B, X)] + [np.array(res2.output()), g2[A], g2[X]])
    for i, (w_, g_) in enumerate(zip(out[ref], out[api])):
        close(g_, w_, f"matmulvecs item {i}")
    with pytest.raises(ValueError, match="invalid matrix data size"):
        api.mat_mul_vecs(api.Variable(np.ones(5)), 2, 3, api.Variable(np.ones(3)))
    with pytest.raises(ValueError, match="invalid vecs size"):
        api.mat_mul_vecs(api.Variable(np.ones(6)), 2, 3, api.Variable(np.ones(4)))


# ---- Gradient.AddToVars / Zero / Scale / Add / Copy on the device ------------------
def test_device_gradient_sgd_loop_matches_oracle(api):
    """Five SGD steps (gradient.go:40-52) with parameters, gradients and the update resident in HBM; nothing is
    downloaded until the end.  The oracle runs the same loop on the host."""
    sizes, n, lr, steps = [20, 32, 6], 16, -0.05, 5
    onet = ref.MLP(sizes, weight_scale=lambda k: 1 / np.sqrt(k))
    x = onet.make_input(n)
    target = ref.splitmix_uniform(91, n * sizes[-1], 0.0, 1.0)
    finals = {}
    for ns in (ref, api):
        variables = [ns.Variable(v.vector) for v in onet.variables]
        layers = []
        for l in range(len(sizes) - 1):
            layers += [ns.LinTran(variables[2 * l], sizes[l + 1], sizes[l]), ns.LinAdd(variables[2 * l + 1]), ns.Sigmoid()]
        net = ns.ComposedRBatcher(layers)
        xv, tv = ns.Variable(x.vector), ns.Variable(target)
        grad = api.DeviceGradient(variables) if ns is api else ref.new_gradient(variables)
        costs = []
        for _ in range(steps):
            cost = ns.squared_error(net.batch(xv, n), tv)
            cost.propagate_gradient(np.array([1.0]), grad)
            if ns is api:
                grad.scale(0.5)
                grad.add_to_vars(2 * lr)
                grad.zero()
            else:
                for gv in grad.values():
                    gv *= 0.5
                ref.gradient_add_to_vars(grad, 2 * lr)
                for gv in grad.values():
                    gv[:] = 0
            costs.append(cost)
        finals[ns] = [np.array(c.output()) for c in costs] + [np.array(v.vector) for v in variables]
    assert finals[ref][0][0] > finals[ref][steps - 1][0]  # the loop does descend
    for i, (w, g) in enumerate(zip(finals[ref], finals[api])):
        close(g, w, f"sgd item {i}")


def test_device_gradient_map_semantics(api):
    a, b = api.Variable([1.0, 2.0, 3.0]), api.Variable([0.5, -0.5])
    g = api.DeviceGradient([a])
    res = api.mul(api.add(a, a), a)  # 2a^2 -> gradient 4a
    res.propagate_gradient(np.ones(3), g)
    res.propagate_gradient(np.ones(3), g)  # accumulates (variable.go:45)
    assert b not in g and np.allclose(g.host()[a], 8 * a.vector)
    c = g.copy()
    g.add(c)
    assert np.allclose(g.host()[a], 16 * a.vector) and np.allclose(c.host()[a], 8 * a.vector)
    g.add_to_vars(0.25)
    assert np.allclose(a.vector, [5.0, 10.0, 15.0])  # lazily downloaded after the device-side update
    a.set([1.0, 1.0, 1.0])  # host write wins again
    assert np.allclose(api.add(a, a).output(), 2.0)


def test_mlp_plan_training_loop_in_hbm_matches_oracle(api):
    """af_mlp_step + af_mlp_add_to_vars: K gradient-ascent steps on the bench objective (outGrad = 1,
    bench/mlp.go:118-126) without the gradient ever leaving the device, vs the oracle's host loop."""
    sizes, n, lr, steps = [40, 64, 24, 10], 96, 0.01, 4
    onet = ref.MLP(sizes, weight_scale=lambda k: 1 / np.sqrt(k))
    x = onet.make_input(n)
    plan = api.MLPPlan(sizes, n)
    for i, v in enumerate(onet.variables):
        plan.set_param(i, v.vector)
        plan.set_rvector(i, None)
    ctx = api.context()
    api.check(ctx.lib.af_upload(ctx.h, plan.input().ptr, x.vector.ctypes.data, x.vector.size))
    for _ in range(steps):
        plan.zero_grads()
        api.check(ctx.lib.af_fill(ctx.h, plan.upstream().ptr, 1.0, n * sizes[-1]))
        plan.step(with_r=False, backprop=True)
        plan.add_to_vars(lr)
        _, grad = onet.step(x, n)
        ref.gradient_add_to_vars(grad, lr)
    for i, v in enumerate(onet.variables):
        close(plan.get_param(i), v.vector, f"param {i} after {steps} steps")
    plan.close()


def test_lstm_plan_add_to_vars(api):
    from autofunc_b200.lstm import LSTMPlan
    I, H, O, T, B = 6, 16, 4, 5, 3
    onet = ref.LSTMNet(I, H, O, weight_scale=0.3)
    plan = LSTMPlan(I, H, O, T, B)
    for i, v in enumerate(onet.variables):
        plan.set_param(i, v.vector)
    xs = ref.splitmix_uniform(3, T * B * I, 0.0, 1.0)
    ups = ref.splitmix_uniform(4, T * B * O, 0.0, 1.0)
    _, _, grads, _ = plan.run_host(xs, ups, with_r=False)
    plan.add_to_vars(-0.125)
    for i, v in enumerate(onet.variables):
        close(plan.param(i).numpy(), v.vector - 0.125 * grads[i], f"lstm param {i}")
    plan.close()


def test_packed_map_rbatcher_matches_host_path_and_oracle(api):
    """seqfunc.MapRBatcher over a list packed time-major in HBM (no per-timestep host copies) equals the host
    joinTime/appendSplit path and the oracle's vector-by-vector application (seqfunc/test/map_test.go:88-120);
    the sequence gradient arrives packed in one Variable."""
    from autofunc_b200 import seq
    sizes = [4, 6, 2]
    onet = ref.MLP(sizes, weight_scale=lambda k: 1 / np.sqrt(k))
    pvars = [api.Variable(v.vector) for v in onet.variables]
    prv = {pv: onet.rvec[ov] for pv, ov in zip(pvars, onet.variables)}
    layers = api.fuse_dense([api.LinTran(pvars[0], 6, 4), api.LinAdd(pvars[1]), api.Sigmoid(),
                             api.LinTran(pvars[2], 2, 6), api.LinAdd(pvars[3]), api.Sigmoid()])
    lens = [3, 1, 0, 2, 5]
    seqs = [[ref.splitmix_uniform(700 + 10 * l + t, 4) for t in range(n)] for l, n in enumerate(lens)]
    rseqs = [[ref.splitmix_uniform(750 + 10 * l + t, 4) for t in range(n)] for l, n in enumerate(lens)]
    ups = [[ref.splitmix_uniform(800 + 10 * l + t, 2) for t in range(n)] for l, n in enumerate(lens)]
    uprs = [[ref.splitmix_uniform(850 + 10 * l + t, 2) for t in range(n)] for l, n in enumerate(lens)]
    plens, widths, flat = seq.PackedSeqs.pack_host(seqs)
    _, _, rflat = seq.PackedSeqs.pack_host(rseqs)
    xin = api.Variable(flat)
    rv = dict(prv)
    rv[xin] = rflat
    res = seq.MapRBatcher(layers).apply_seqs_r(rv, seq.PackedVarResult(xin, plens, widths, rv))
    # oracle: every vector on its own
    per = ref.ComposedRFunc(list(onet.layers))
    ograd, orgrad = ref.new_gradient(onet.variables), ref.new_rgradient(onet.variables)
    xg = [[None] * n for n in lens]
    xrg = [[None] * n for n in lens]
    outs, routs = res.output_seqs(), res.routput_seqs()
    for l, s in enumerate(seqs):
        assert len(outs[l]) == len(s)
        for t, v in enumerate(s):
            xv = ref.Variable(v)
            orv = dict(onet.rvec)
            orv[xv] = rseqs[l][t]
            r = per.apply_r(orv, ref.RVariable(xv, orv))
            close(outs[l][t], r.output(), f"lane {l} t {t}")
            close(routs[l][t], r.routput(), f"lane {l} t {t} R")
            g1, rg1 = {xv: np.zeros(4)}, {xv: np.zeros(4)}
            g1.update(ograd)
            rg1.update(orgrad)
            r.propagate_rgradient(ups[l][t].copy(), uprs[l][t].copy(), rg1, g1)
            xg[l][t], xrg[l][t] = g1[xv], rg1[xv]
    grad, rgrad = api.new_gradient(pvars + [xin]), api.new_rgradient(pvars + [xin])
    res.propagate_rgradient(ups, uprs, rgrad, grad)
    for i, (pv, ov) in enumerate(zip(pvars, onet.variables)):
        close(grad[pv], ograd[ov], f"grad[{i}]")
        close(rgrad[pv], orgrad[ov], f"rgrad[{i}]")
    shape = seq.PackedSeqs(plens, widths, api.context().upload(flat))
    for got, want, what in ((shape.unpack_host(grad[xin]), xg, "input grad"), (shape.unpack_host(rgrad[xin]), xrg, "input rgrad")):
        for l in range(len(lens)):
            for t in range(lens[l]):
                close(got[l][t], want[l][t], f"{what} lane {l} t {t}")
    assert set(grad) == set(pvars + [xin])  # temporaries deleted again (map_batcher.go:136-137)
    # the host joinTime path gives the same sequences
    res_h = seq.MapRBatcher(layers).apply_seqs_r(prv, seq.ConstRResult(seqs))
    for l in range(len(lens)):
        for t in range(lens[l]):
            close(outs[l][t], res_h.output_seqs()[l][t], f"host path lane {l} t {t}")
    # non-R mapping over a constant packed list, device-resident gradient map
    res2 = seq.MapBatcher(layers).apply_seqs(seq.PackedConstResult.from_host(seqs))
    dg = api.DeviceGradient(pvars)
    res2.propagate_gradient(ups, dg)
    for i, (pv, ov) in enumerate(zip(pvars, onet.variables)):
        close(dg.host()[pv], ograd[ov], f"device grad[{i}]")


def test_map_batcher_output_on_reference_seq_fixtures(api):
    """seqfunc/test/map_test.go:88-120 (TestMapBatcherOutput) on the reference's own TestSeqs / TranMatrix fixtures
    (tests/golden/reference_fixtures.json): mapping the Batcher over the lanes -- through the host joinTime path AND
    the packed device path -- equals applying the Func vector by vector; gradients/R-gradients of TranMatrix equal the
    oracle's per-vector sums."""
    import json
    import os
    from autofunc_b200 import seq
    G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_fixtures.json")))["seqfunc/test/map_test.go"]
    tv = [np.array(v) for v in G["TestVars"]]
    lanes = [[0, 1, 2], [2, 0, 3], [1], [3], [0, 3]]  # map_test.go:28-34
    seqs = [[tv[i] for i in lane] for lane in lanes]
    M, RM = np.array(G["TranMatrix"]), np.array(G["rv_TranMatrix"])
    pm = api.Variable(M)
    plt = api.LinTran(pm, 4, 4)
    om = ref.Variable(M)
    olt = ref.LinTran(om, 4, 4)
    expected = [[olt.apply(ref.Variable(v)).output() for v in s] for s in seqs]
    for res in (seq.MapBatcher(plt).apply_seqs(seq.ConstResult(seqs)),
                seq.MapBatcher(plt).apply_seqs(seq.PackedConstResult.from_host(seqs))):
        got = res.output_seqs()
        assert [len(s) for s in got] == [len(s) for s in expected]
        for l, s in enumerate(expected):
            for t, v in enumerate(s):
                assert np.max(np.abs(got[l][t] - v)) <= 1e-5  # the reference's tolerance
                close(got[l][t], v, f"lane {l} t {t}")
    ups = [[ref.splitmix_uniform(60 + 7 * l + t, 4) for t in range(len(s))] for l, s in enumerate(seqs)]
    uprs = [[ref.splitmix_uniform(90 + 7 * l + t, 4) for t in range(len(s))] for l, s in enumerate(seqs)]
    og, org = {om: np.zeros(16)}, {om: np.zeros(16)}
    for l, s in enumerate(seqs):
        for t, v in enumerate(s):
            r = olt.apply_r({om: RM}, ref.RVariable(ref.Variable(v), {}))
            r.propagate_rgradient(ups[l][t].copy(), uprs[l][t].copy(), org, og)
    for inp in (seq.ConstRResult(seqs), seq.PackedConstResult.from_host(seqs)):
        res = seq.MapRBatcher(plt).apply_seqs_r({pm: RM}, inp)
        g, rg = api.new_gradient([pm]), api.new_rgradient([pm])
        res.propagate_rgradient(ups, uprs, rg, g)
        close(g[pm], og[om], "TranMatrix gradient")
        close(rg[pm], org[om], "TranMatrix R-gradient")


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_networks_pass_the_reference_acceptance_procedure(api, seed):
    """functest.RFuncChecker.FullCheck on the product for the same random networks the oracle is checked on."""
    f = fx.RandomNetFixture(api, seed)
    assert RFuncChecker(api, f.func, f.variables, f.input, f.rv).full_check() == []
