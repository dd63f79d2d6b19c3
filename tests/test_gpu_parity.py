"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI,
against the CPU oracle on identical seeded inputs.

Tolerance (BASELINE.json north_star): |gpu - oracle| <= 1e-12 + 1e-10 * |oracle| element-wise,
float64.  The reference's own KATs and its FullCheck acceptance procedure run through the
product API as well."""
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
    assert torch.cuda.is_available(), "GPU tests need a B200"
    from autofunc_b200 import api as a
    a.context()  # raises loudly if the extension is missing or the device is not sm_100
    return a


def close(gpu, want, what=""):
    gpu, want = np.asarray(gpu), np.asarray(want)
    assert gpu.shape == want.shape, f"{what}: shape {gpu.shape} vs {want.shape}"
    err = np.abs(gpu - want)
    lim = ATOL + RTOL * np.abs(want)
    bad = err > lim
    assert not bad.any(), f"{what}: {bad.sum()} of {bad.size} beyond 1e-12+1e-10|ref|; worst err {err.max():.3e} (ref {want.flat[err.argmax()]:.6e})"


# ---- the reference's golden vectors through the product ------------------------------------
def test_lintran_kats(api):
    f = fx.LinTranFixture(api)
    assert np.array_equal(f.mat1.apply(f.vec).output(), fx.LINTRAN_KAT_OUT)           # lin_tran_test.go:65-73
    assert np.array_equal(f.batch_test.apply(f.vec2).output(), fx.LINTRAN_KAT_BATCH_OUT)  # :85-97


def test_lintran_fullcheck(api):  # test/lin_tran_test.go:75-83
    f = fx.LinTranFixture(api)
    assert RFuncChecker(api, api.ComposedRFunc([f.mat1, f.mat2]), f.variables, f.vec, f.rv).full_check() == []


def test_lintran_batch_fullcheck(api):  # test/lin_tran_test.go:99-107
    f = fx.LinTranFixture(api)
    assert RFuncChecker(api, f.batch_test, f.variables, f.vec2, f.rv).full_check() == []


def test_linadd_fullcheck(api):  # test/lin_add_test.go:21-30
    v1, v2 = api.Variable(fx.LINADD_VEC1), api.Variable(fx.LINADD_VEC2)
    rv = {v1: np.array(fx.LINADD_RV1), v2: np.array(fx.LINADD_RV2)}
    f = api.ComposedRFunc([api.LinAdd(v1), api.LinAdd(v2)])
    assert RFuncChecker(api, f, [v1, v2], v2, rv).full_check() == []


@pytest.mark.parametrize("which", ["exp", "sqnorm", "sigmoid3"])
def test_math_func_fullcheck(api, which):  # test/math_funcs_test.go:22-60
    make = {"exp": lambda: api.Exp(), "sqnorm": lambda: api.SquaredNorm(),
            "sigmoid3": lambda: api.ComposedRFunc([api.Sigmoid(), api.Sigmoid(), api.Sigmoid()])}[which]
    v = api.Variable(fx.MATH_VEC)
    rv = {v: np.array(fx.MATH_RV)}
    assert RFuncChecker(api, make(), [v], v, rv).full_check() == []


def test_arithmetic_subset_fullcheck(api):
    a, b = api.Variable([1, -2, 3, 0.5]), api.Variable([0.3, 2, -1, 4])
    rv = {a: np.array([1, 2, -3, 0.25])}  # b deliberately absent (test/arithmetic_test.go:33-34)

    class F:
        def apply(self, x):
            d = api.sub(x, b)
            return api.scale(api.sum_all(api.square(api.add(api.mul(d, x), d))), 0.5)

        def apply_r(self, r, x):
            d = api.sub_r(x, api.RVariable(b, r))
            return api.scale_r(api.sum_all_r(api.square_r(api.add_r(api.mul_r(d, x), d))), 0.5)
    assert RFuncChecker(api, F(), [a, b], a, rv, prec=1e-4).full_check() == []


def test_fused_dense_fullcheck(api):
    """The fused LinTran+LinAdd+Sigmoid layer passes the reference's acceptance procedure too."""
    rng = np.random.default_rng(5)
    w1, b1 = api.Variable(rng.uniform(-1, 1, 12)), api.Variable(rng.uniform(-1, 1, 3))
    w2, b2 = api.Variable(rng.uniform(-1, 1, 6)), api.Variable(rng.uniform(-1, 1, 2))
    x = api.Variable(rng.uniform(-1, 1, 8))
    allv = [w1, b1, w2, b2, x]
    rv = {v: rng.uniform(-1, 1, v.vector.size) for v in allv[:-2]}  # b2, x absent from the RVector... x present below
    rv[x] = rng.uniform(-1, 1, 8)
    l1, l2 = api.DenseSigmoid(w1, b1, 3, 4), api.DenseSigmoid(w2, b2, 2, 3)

    class F:
        def apply(self, v):
            return l2.batch(l1.batch(v, 2), 2)

        def apply_r(self, r, v):
            return l2.batch_r(r, l1.batch_r(r, v, 2), 2)
    assert RFuncChecker(api, F(), allv, x, rv).full_check() == []


# ---- per-kernel parity on shapes that take the TMA + DMMA path -------------------------------
SHAPES = [  # rows(M), cols(K), n
    (1024, 2048, 10),   # bench/lintran.go:13-17 (config C1)
    (2000, 1000, 64),   # MLP layer 1 (bench/mlp.go:17-20)
    (512, 2000, 33),
    (10, 512, 300),     # last MLP layer: 10 outputs
    (130, 66, 257),     # ragged in every dimension, all tails
    (64, 16, 128),
    (8, 2, 1),          # single sample, one k-tile mostly zero fill
]


def _rand(seed, n):
    return ref.splitmix_uniform(seed, n)


@pytest.mark.parametrize("rows,cols,n", SHAPES)
# path 2 = no generic fallback (tensor-core kernel in both tile configs; n <= 16 takes the streaming Gemv/Ger
# kernels unless small_n = 0 forces those shapes through the tensor-core kernel too), path 1 = generic kernel
@pytest.mark.parametrize("path,big_cfg,small_n", [(2, 0, 1), (2, 2, 1), (2, 2, 0), (1, 0, 1)])
def test_lintran_kernels_match_oracle(api, rows, cols, n, path, big_cfg, small_n):
    ctx = api.context()
    ctx.set_gemm_path(path)
    api.check(ctx.lib.af_set_option(ctx.h, b"big_cfg", big_cfg))
    api.check(ctx.lib.af_set_option(ctx.h, b"small_n", small_n))
    try:
        W, RW = _rand(1, rows * cols), _rand(2, rows * cols)
        X, RX = _rand(3, n * cols), _rand(4, n * cols)
        U, UR = _rand(5, n * rows), _rand(6, n * rows)
        wv, xv = ref.Variable(W), ref.Variable(X)
        lt = ref.LinTran(wv, rows, cols)
        rdata = ref.RVariable(wv, {wv: RW})

        class RIn:
            def output(self): return X
            def routput(self): return RX
        # oracle
        y = lt.multiply(X)
        ry = lt.multiply_r(rdata, RIn())
        g0, rg0 = _rand(7, rows * cols), _rand(8, rows * cols)  # non-zero accumulators: += semantics
        g, rg = {wv: g0.copy()}, g0 * 0 + rg0
        lt.data_gradient(U, X, g)
        lt.data_gradient_r(U, UR, X, RX, rg)
        d = lt.input_gradient(U)
        dr = lt.input_gradient_r(rdata, U, UR)
        # product, through the C ABI
        L, h = ctx.lib, ctx.h
        dW, dRW, dX, dRX, dU, dUR = (ctx.upload(a) for a in (W, RW, X, RX, U, UR))
        dY, dRY = ctx.empty(n * rows), ctx.empty(n * rows)
        api.check(L.af_lintran_fwd(h, dW.ptr, dX.ptr, dY.ptr, rows, cols, n))
        close(dY.numpy(), y, "G1 multiply")
        dY2 = ctx.empty(n * rows)
        api.check(L.af_lintran_fwd_r(h, dW.ptr, dRW.ptr, dX.ptr, dRX.ptr, dY2.ptr, dRY.ptr, rows, cols, n))
        close(dY2.numpy(), y, "G1 (dual)")
        close(dRY.numpy(), ry, "G2 multiplyR")
        # RX absent == zero R-vector (variable.go:85-91)
        api.check(L.af_lintran_fwd_r(h, dW.ptr, dRW.ptr, dX.ptr, None, None, dRY.ptr, rows, cols, n))
        close(dRY.numpy(), (X.reshape(n, cols) @ RW.reshape(rows, cols).T).reshape(-1), "G2 with RX = 0")
        dG, dRG = ctx.upload(g0), ctx.upload(rg0)
        api.check(L.af_lintran_wgrad(h, dU.ptr, dX.ptr, dG.ptr, rows, cols, n))
        close(dG.numpy(), g[wv], "G3 dataGradient")
        dG2 = ctx.upload(g0)
        api.check(L.af_lintran_wgrad_r(h, dU.ptr, dUR.ptr, dX.ptr, dRX.ptr, dG2.ptr, dRG.ptr, rows, cols, n))
        close(dG2.numpy(), g[wv], "G3 (dual)")
        close(dRG.numpy(), rg, "G4 dataGradientR")
        dD, dDR = ctx.empty(n * cols), ctx.empty(n * cols)
        api.check(L.af_lintran_igrad(h, dU.ptr, dW.ptr, dD.ptr, rows, cols, n))
        close(dD.numpy(), d, "G5 inputGradient")
        dD2 = ctx.empty(n * cols)
        api.check(L.af_lintran_igrad_r(h, dU.ptr, dUR.ptr, dW.ptr, dRW.ptr, dD2.ptr, dDR.ptr, rows, cols, n))
        close(dD2.numpy(), d, "G5 (dual)")
        close(dDR.numpy(), dr, "G6 inputGradientR")
    finally:
        ctx.set_gemm_path(0)
        api.check(ctx.lib.af_set_option(ctx.h, b"big_cfg", 2))
        api.check(ctx.lib.af_set_option(ctx.h, b"small_n", 1))


def test_odd_pitch_falls_back_or_refuses(api):
    """Odd row pitch violates TMA's 16-byte rule: auto mode must still be right (generic kernel),
    and 'require DMMA' mode must refuse instead of silently falling back."""
    ctx = api.context()
    rows, cols, n = 5, 7, 3
    W, X = _rand(11, rows * cols), _rand(12, n * cols)
    dW, dX, dY = ctx.upload(W), ctx.upload(X), ctx.empty(n * rows)
    api.check(ctx.lib.af_lintran_fwd(ctx.h, dW.ptr, dX.ptr, dY.ptr, rows, cols, n))
    close(dY.numpy(), (X.reshape(n, cols) @ W.reshape(rows, cols).T).reshape(-1), "odd pitch")
    ctx.set_gemm_path(2)
    try:
        assert ctx.lib.af_lintran_fwd(ctx.h, dW.ptr, dX.ptr, dY.ptr, rows, cols, n) == -5  # AF_ENOTSUP
    finally:
        ctx.set_gemm_path(0)


def test_elementwise_kernels_match_oracle(api):
    ctx = api.context()
    L, h = ctx.lib, ctx.h
    n, length = 37, 1001
    N = n * length
    X, RX, U, UR = _rand(21, N) * 4, _rand(22, N), _rand(23, N), _rand(24, N)
    b, Rb = _rand(25, length), _rand(26, length)
    dX, dRX, db, dRb = ctx.upload(X), ctx.upload(RX), ctx.upload(b), ctx.upload(Rb)
    dS, dRS = ctx.empty(N), ctx.empty(N)
    api.check(L.af_bias_fwd_r(h, dX.ptr, dRX.ptr, db.ptr, dRb.ptr, dS.ptr, dRS.ptr, length, n))
    close(dS.numpy(), (X.reshape(n, length) + b).reshape(-1), "bias fwd")
    close(dRS.numpy(), (RX.reshape(n, length) + Rb).reshape(-1), "bias fwd R")
    api.check(L.af_sigmoid_fwd_r(h, dX.ptr, dRX.ptr, dS.ptr, dRS.ptr, N))
    s = 1 / (1 + np.exp(-X))
    rs = s * (1 - s) * RX
    close(dS.numpy(), s, "sigmoid")
    close(dRS.numpy(), rs, "sigmoid R")
    dU, dUR = ctx.upload(U), ctx.upload(UR)
    api.check(L.af_sigmoid_bwd_r(h, dS.ptr, dRS.ptr, dU.ptr, dUR.ptr, N))
    close(dU.numpy(), s * (1 - s) * U, "sigmoid bwd")
    close(dUR.numpy(), rs * (1 - s) * U - s * rs * U + s * (1 - s) * UR, "sigmoid bwd R")
    gb0, rgb0 = _rand(27, length), _rand(28, length)
    dgb, drgb = ctx.upload(gb0), ctx.upload(rgb0)
    dU, dUR = ctx.upload(U), ctx.upload(UR)
    api.check(L.af_bias_grad_r(h, dU.ptr, dUR.ptr, dgb.ptr, drgb.ptr, length, n))
    close(dgb.numpy(), gb0 + U.reshape(n, length).sum(0), "bias grad")
    close(drgb.numpy(), rgb0 + UR.reshape(n, length).sum(0), "bias grad R")
    out, outr = ctx.empty(1), ctx.empty(1)
    api.check(L.af_sum_all(h, dX.ptr, dRX.ptr, out.ptr, outr.ptr, N))
    close(out.numpy(), [X.sum()], "sum_all")
    close(outr.numpy(), [RX.sum()], "sum_all R")


# ---- composed networks -------------------------------------------------------------------------
def _oracle_and_product_nets(api, sizes, weight_scale=None):
    onet = ref.MLP(sizes, weight_scale=weight_scale)
    pvars = [api.Variable(v.vector) for v in onet.variables]
    prv = {pv: onet.rvec[ov] for pv, ov in zip(pvars, onet.variables)}
    layers = api.ComposedRBatcher()
    for k in range(len(sizes) - 1):
        layers += [api.LinTran(pvars[2 * k], sizes[k + 1], sizes[k]), api.LinAdd(pvars[2 * k + 1]), api.Sigmoid()]
    return onet, pvars, prv, layers


@pytest.mark.parametrize("sizes,n,scaled", [([1000, 2000, 512, 10], 16, False), ([1000, 2000, 512, 10], 16, True),
                                            ([30, 20, 6], 5, True), ([64, 48, 32, 16], 130, True)])
@pytest.mark.parametrize("fused", [False, True])
def test_mlp_step_through_operator_api(api, sizes, n, scaled, fused):
    """bench/mlp.go:43-57 (RunR with backProp) and :29-41 (Run) driven through the reference-shaped
    API: ComposedRBatcher of LinTran/LinAdd/Sigmoid, or the same net rewritten by fuse_dense."""
    ws = (lambda k: 1 / np.sqrt(k)) if scaled else None
    onet, pvars, prv, layers = _oracle_and_product_nets(api, sizes, ws)
    if fused:
        layers = api.fuse_dense(layers)
    x = onet.make_input(n)
    up = ref.splitmix_uniform(900, n * sizes[-1])
    upr = ref.splitmix_uniform(901, n * sizes[-1])
    ores, ograd, orgrad = onet.step_r(x, n, up, upr)
    px = api.Variable(x.vector)
    res = layers.batch_r(prv, api.RVariable(px, prv), n)
    grad, rgrad = api.new_gradient(pvars), api.new_rgradient(pvars)
    res.propagate_rgradient(up.copy(), upr.copy(), rgrad, grad)
    close(res.output(), ores.output(), "Output")
    close(res.routput(), ores.routput(), "ROutput")
    for i, (pv, ov) in enumerate(zip(pvars, onet.variables)):
        close(grad[pv], ograd[ov], f"grad[{i}]")
        close(rgrad[pv], orgrad[ov], f"rgrad[{i}]")
    # non-R twin
    ores2, ograd2 = onet.step(x, n, up)
    res2 = layers.batch(px, n)
    grad2 = api.new_gradient(pvars)
    res2.propagate_gradient(up.copy(), grad2)
    close(res2.output(), ores2.output(), "Output (Apply)")
    for i, (pv, ov) in enumerate(zip(pvars, onet.variables)):
        close(grad2[pv], ograd2[ov], f"grad[{i}] (Apply)")


def test_gradient_pruning_and_nil_grad(api):
    """Only map keys receive gradient (lin_tran.go:374-386,410-420); grad may be nil (result.go:62-65)."""
    sizes, n = [16, 12, 8], 6
    onet, pvars, prv, layers = _oracle_and_product_nets(api, sizes, lambda k: 1 / np.sqrt(k))
    x = onet.make_input(n)
    px = api.Variable(x.vector)
    res = layers.batch_r(prv, api.RVariable(px, prv), n)
    ores = onet.layers.batch_r(onet.rvec, ref.RVariable(x, onet.rvec), n)
    only = [2]  # W of layer 2 only
    rgrad = api.new_rgradient([pvars[i] for i in only])
    orgrad = ref.new_rgradient([onet.variables[i] for i in only])
    res.propagate_rgradient(np.ones(n * 8), np.zeros(n * 8), rgrad, None)
    ores.propagate_rgradient(np.ones(n * 8), np.zeros(n * 8), orgrad, None)
    assert set(rgrad) == {pvars[2]}
    close(rgrad[pvars[2]], orgrad[onet.variables[2]], "rgrad pruned")


@pytest.mark.parametrize("scaled", [False, True])
@pytest.mark.parametrize("with_r", [True, False])
@pytest.mark.parametrize("n", [96, 1, 10])  # n = 1 is the reference bench exactly as written (bench/mlp.go:36,52)
def test_mlp_plan_matches_oracle(api, scaled, with_r, n):
    """af_mlp_step_host (the whole-network fast path, host buffers in and out) vs the oracle on the
    bench network 1000-2000-512-10 at a batch the oracle finishes in seconds."""
    sizes = [1000, 2000, 512, 10]
    ws = (lambda k: 1 / np.sqrt(k)) if scaled else None
    onet = ref.MLP(sizes, weight_scale=ws)
    plan = api.MLPPlan(sizes, n)
    for i, v in enumerate(onet.variables):
        plan.set_param(i, v.vector)
        plan.set_rvector(i, onet.rvec[v])
    x = onet.make_input(n)
    if with_r:
        ores, ograd, orgrad = onet.step_r(x, n)
    else:
        ores, ograd = onet.step(x, n)
    out, rout, grads, rgrads = plan.step_host(x.vector, with_r=with_r)
    close(out, ores.output(), "Output")
    if with_r:
        close(rout, ores.routput(), "ROutput")
    for i, v in enumerate(onet.variables):
        close(grads[i], ograd[v], f"grad[{i}]")
        if with_r:
            close(rgrads[i], orgrad[v], f"rgrad[{i}]")
    # accumulation: a second resident step without zeroing doubles the accumulators (bench/mlp.go:33-40)
    api.check(plan.ctx.lib.af_fill(plan.ctx.h, plan.upstream().ptr, 1.0, n * 10))
    api.check(plan.ctx.lib.af_zero(plan.ctx.h, plan.upstream_r().ptr, n * 10))
    plan.step(with_r=with_r)
    close(plan.grad(0).numpy(), 2 * ograd[onet.variables[0]], "accumulated grad[0]")
    # third and fourth steps: for batches <= 256 these replay the captured CUDA graph of the step
    for k in (3, 4):
        api.check(plan.ctx.lib.af_fill(plan.ctx.h, plan.upstream().ptr, 1.0, n * 10))
        api.check(plan.ctx.lib.af_zero(plan.ctx.h, plan.upstream_r().ptr, n * 10))
        plan.step(with_r=with_r)
        close(plan.grad(0).numpy(), k * ograd[onet.variables[0]], f"accumulated grad[0] x{k}")
        close(plan.grad(5).numpy(), k * ograd[onet.variables[5]], f"accumulated grad[5] x{k}")
    close(plan.output().numpy(), ores.output(), "Output after graph replay")
    plan.close()


def test_mlp_pipelined_host_steps_match_blocking_steps(api):
    """af_mlp_submit_host / af_mlp_wait (two steps in flight, double-buffered slots) give the same results
    as af_mlp_step_host, step by step, for different inputs per step."""
    sizes, n, steps = [64, 48, 10], 200, 5
    onet = ref.MLP(sizes, weight_scale=lambda k: 1 / np.sqrt(k))
    plan = api.MLPPlan(sizes, n)
    for i, v in enumerate(onet.variables):
        plan.set_param(i, v.vector)
        plan.set_rvector(i, onet.rvec[v])
    xs = [ref.splitmix_uniform(300 + k, n * sizes[0]) for k in range(steps)]
    want = [plan.step_host(x) for x in xs]
    P = 2 * plan.L
    bufs = [dict(out=np.empty(n * 10), rout=np.empty(n * 10), grads=[np.empty(plan.param_len(i)) for i in range(P)],
                 rgrads=[np.empty(plan.param_len(i)) for i in range(P)]) for _ in range(steps)]
    for k in range(steps):
        plan.submit_host(xs[k], **bufs[k])
    for k in range(steps):
        plan.wait()
    for k in range(steps):
        out, rout, grads, rgrads = want[k]
        # (tolerance, not bit equality: contractions that split their k range combine partial tiles with floating-point
        # atomics, so two runs of the same step may differ in the last bits -- DESIGN.md 3.1 "reproducibility")
        close(bufs[k]["out"], out, f"step {k} Output")
        close(bufs[k]["rout"], rout, f"step {k} ROutput")
        for i in range(P):
            close(bufs[k]["grads"][i], grads[i], f"step {k} grad[{i}]")
            close(bufs[k]["rgrads"][i], rgrads[i], f"step {k} rgrad[{i}]")
    # and the blocking call still works afterwards (rebinds slot 0)
    out, rout, grads, rgrads = plan.step_host(xs[0])
    close(out, want[0][0], "Output of the blocking call after the pipelined ones")
    plan.close()


def test_mlp_grad_ready_hook_covers_accumulators_once(api):
    """af_mlp_set_grad_ready: the ranges reported during a step tile the accumulator block exactly once (so a
    data-parallel caller all-reduces every gradient once), layer 1's weight gradient arrives in row chunks,
    and chunking does not change the results."""
    import ctypes as C
    from autofunc_b200 import _lib
    sizes, n = [128, 320, 64, 10], 300
    onet = ref.MLP(sizes, weight_scale=lambda k: 1 / np.sqrt(k))
    plan = api.MLPPlan(sizes, n)
    for i, v in enumerate(onet.variables):
        plan.set_param(i, v.vector)
        plan.set_rvector(i, onet.rvec[v])
    x = onet.make_input(n)
    want = plan.step_host(x.vector)
    ranges = []
    cb = _lib.GRAD_READY_FN(lambda user, ptr, count: ranges.append((ptr, count)))
    api.check(plan.ctx.lib.af_mlp_set_grad_ready(plan.h, C.cast(cb, C.c_void_p), None, 4))
    got = plan.step_host(x.vector)
    api.check(plan.ctx.lib.af_mlp_set_grad_ready(plan.h, None, None, 1))
    for a, b in zip(want[2] + want[3], got[2] + got[3]):
        close(b, a, "gradients with chunked layer-1 weight gradient")
    base = plan.grad(0).ptr
    end = plan.input().ptr
    covered = np.zeros((end - base) // 8, dtype=np.int32)
    for ptr, count in ranges:
        assert base <= ptr and ptr + 8 * count <= end
        covered[(ptr - base) // 8:(ptr - base) // 8 + count] += 1
    assert covered.max() == 1
    for i in range(2 * plan.L):  # every gradient and R-gradient element is inside exactly one reported range
        for vec in (plan.grad(i), plan.rgrad(i)):
            o = (vec.ptr - base) // 8
            assert (covered[o:o + vec.n] == 1).all()
    assert len(ranges) > 2 + 2 * 3  # layer-1 weight gradient came in several chunks
    plan.close()


def test_mlp_plan_missing_rvectors(api):
    """Variables absent from the RVector have zero R (variable.go:85-91; arithmetic_test.go:33-34)."""
    sizes, n = [32, 24, 8], 40
    onet = ref.MLP(sizes, weight_scale=lambda k: 1 / np.sqrt(k))
    del onet.rvec[onet.variables[1]], onet.rvec[onet.variables[2]]
    plan = api.MLPPlan(sizes, n)
    for i, v in enumerate(onet.variables):
        plan.set_param(i, v.vector)
        if v in onet.rvec:
            plan.set_rvector(i, onet.rvec[v])
    x = onet.make_input(n)
    ores, ograd, orgrad = onet.step_r(x, n)
    out, rout, grads, rgrads = plan.step_host(x.vector)
    close(rout, ores.routput(), "ROutput")
    for i, v in enumerate(onet.variables):
        close(grads[i], ograd[v], f"grad[{i}]")
        close(rgrads[i], orgrad[v], f"rgrad[{i}]")
    plan.close()


# ---- full-size, size-independent properties (BASELINE sizes; no oracle run needed) ---------------
def test_full_size_properties(api):
    sizes, n = [1000, 2000, 512, 10], 4096
    from autofunc_b200 import synth
    params, rvecs = synth.mlp_params(sizes, weight_scale=lambda k: 1 / np.sqrt(k))
    x = synth.mlp_input(sizes, n)
    plan = api.MLPPlan(sizes, n)
    for i, (p, r) in enumerate(zip(params, rvecs)):
        plan.set_param(i, p)
        plan.set_rvector(i, r)
    out, rout, grads, rgrads = plan.step_host(x)
    # (1) Apply and ApplyR agree (functest/rfunc_checker.go:208-243)
    out0, _, grads0, _ = plan.step_host(x, with_r=False)
    close(out0, out, "Apply vs ApplyR output")
    for i in range(6):
        close(grads0[i], grads[i], f"Apply vs ApplyR grad[{i}]")
    # (2) the R-operator is linear in the R-vector: doubling rv doubles ROutput and the R-gradient
    for i, r in enumerate(rvecs):
        plan.set_rvector(i, 2 * r)
    out2, rout2, grads2, rgrads2 = plan.step_host(x)
    close(rout2, 2 * rout, "R-linearity ROutput")
    for i in range(6):
        close(rgrads2[i], 2 * rgrads[i], f"R-linearity rgrad[{i}]")
        close(grads2[i], grads[i], f"grad independent of rv [{i}]")
    plan.close()
    # (3) gradients are sums over samples: two half batches add up to the full batch
    half = api.MLPPlan(sizes, n // 2)
    for i, (p, r) in enumerate(zip(params, rvecs)):
        half.set_param(i, p)
        half.set_rvector(i, r)
    xa = x.reshape(n, -1)
    _, _, ga, rga = half.step_host(np.ascontiguousarray(xa[: n // 2]).reshape(-1))
    _, _, gb, rgb = half.step_host(np.ascontiguousarray(xa[n // 2:]).reshape(-1))
    for i in range(6):
        close(ga[i] + gb[i], grads[i], f"batch additivity grad[{i}]")
        close(rga[i] + rgb[i], rgrads[i], f"batch additivity rgrad[{i}]")
    half.close()
    assert np.isfinite(out).all() and np.isfinite(rout).all()


def test_hessian_row_magnitudes_match_oracle(api):
    """README.md:5 use case (BASELINE config 3 at a size the oracle handles): D = sqrt(mean_p (H v_p)^2) from
    R-gradients, product vs oracle with identical probes; and (Hv) agrees with a finite difference of gradients."""
    from autofunc_b200 import hessian
    sizes, n, probes = [24, 32, 16, 6], 20, 3
    onet = ref.MLP(sizes, weight_scale=lambda k: 1 / np.sqrt(k))
    plan = api.MLPPlan(sizes, n)
    for i, v in enumerate(onet.variables):
        plan.set_param(i, v.vector)
    x = onet.make_input(n)
    got = hessian.hessian_row_magnitudes(plan, x.vector, probes, seed=7)
    acc = [np.zeros(v.vector.size) for v in onet.variables]
    for p in range(probes):
        vs = hessian.probe_vectors(plan, 7, p)
        onet.rvec = {v: vs[i] for i, v in enumerate(onet.variables)}
        _, _, rgrad = onet.step_r(x, n)
        for i, v in enumerate(onet.variables):
            acc[i] += rgrad[v] ** 2 / probes
    for i in range(len(acc)):
        close(got[i], np.sqrt(acc[i]), f"row magnitudes of parameter {i}")
    # Hv against a central difference of the (exact) gradient along v: the R-operator's defining property
    d = 1e-6
    base = [v.vector.copy() for v in onet.variables]
    gs = []
    for sgn in (1, -1):
        for v, b0 in zip(onet.variables, base):
            v.vector[:] = b0 + sgn * d * onet.rvec[v]
        gs.append(onet.step(x, n)[1])
    for v, b0 in zip(onet.variables, base):
        v.vector[:] = b0
    _, _, rgrad = onet.step_r(x, n)
    for v in onet.variables:
        fd = (gs[0][v] - gs[1][v]) / (2 * d)
        assert np.abs(fd - rgrad[v]).max() < 1e-6 * max(1.0, np.abs(rgrad[v]).max())
    plan.close()


# ---- error behaviour (panics with fixed strings -> ShapeError with the same text) -------
def test_shape_errors(api):
    w = api.Variable(np.zeros(12))
    lt = api.LinTran(w, 3, 4)
    with pytest.raises(api.ShapeError, match="input length should be 4 but got 5"):
        lt.apply(api.Variable(np.zeros(5)))
    with pytest.raises(api.ShapeError, match="input length should be 8 but got 4"):
        lt.batch(api.Variable(np.zeros(4)), 2)
    with pytest.raises(api.ShapeError, match="vector sizes do not match"):
        api.mul(api.Variable(np.zeros(3)), api.Variable(np.zeros(4)))
    with pytest.raises(api.ShapeError, match="input count does not divide input length"):
        api.Sigmoid().batch(api.Variable(np.zeros(5)), 2)


@pytest.mark.parametrize("sizes,scaled", [([1000, 2000, 512, 10], False), ([1000, 2000, 512, 10], True),
                                          ([2048, 2048, 2048, 2048, 2048], True)])
def test_bench_configuration_matches_oracle_at_full_size(api, sizes, scaled):
    """BASELINE configs[1] exactly as bench.py runs it -- 1000-2000-512-10, n = 4096, RunR with backProp -- and
    configs[2] (the Hessian-vector step on 4 x (2048 -> 2048), n = 4096) against the oracle directly (numpy finishes
    these sizes in seconds): Output, ROutput and every gradient / R-gradient vector."""
    n = 4096
    ws = (lambda k: 1 / np.sqrt(k)) if scaled else None   # None = bench/mlp.go's U[-1,1) init (saturated sigmoids)
    onet = ref.MLP(sizes, weight_scale=ws)
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


# ---- the streaming kernels for very narrow layers (small_n.cu) ------------------------------------------------------
@pytest.mark.parametrize("rows,cols,n", [(10, 512, 4096), (1, 2, 17), (16, 66, 130), (3, 6400, 19), (7, 30, 1000)])
@pytest.mark.parametrize("with_r", [True, False])
@pytest.mark.parametrize("activation", [1, 0])
def test_few_rows_forward_matches_oracle(api, rows, cols, n, with_r, activation):
    """af_dense_fwd on a layer with <= 16 outputs and a large batch takes few_rows_fwd_kernel (one streaming pass with the
    weights staged in shared memory; wider-than-shared-memory weights fall back to the tensor-core tiles): against
    LinTran -> LinAdd (-> Sigmoid) of the oracle, with the R operands present, absent one at a time, and the non-R form."""
    ctx = api.context()
    L, h = ctx.lib, ctx.h
    W, RW, b, Rb = _rand(41, rows * cols), _rand(42, rows * cols), _rand(43, rows), _rand(44, rows)
    X, RX = _rand(45, n * cols), _rand(46, n * cols)
    z = X.reshape(n, cols) @ W.reshape(rows, cols).T + b
    dW, dRW, db, dRb, dX, dRX = (ctx.upload(a) for a in (W, RW, b, Rb, X, RX))
    for use_rw, use_rx in ([(True, True), (False, True), (True, False)] if with_r else [(False, False)]):
        rz = np.zeros((n, rows)) + (Rb if with_r else 0.0)
        if use_rw:
            rz = rz + X.reshape(n, cols) @ RW.reshape(rows, cols).T
        if use_rx:
            rz = rz + RX.reshape(n, cols) @ W.reshape(rows, cols).T
        s = 1 / (1 + np.exp(-z)) if activation else z
        rs = s * (1 - s) * rz if activation else rz
        dS, dRS = ctx.empty(n * rows), (ctx.empty(n * rows) if with_r else None)
        api.check(L.af_dense_fwd(h, dW.ptr, dRW.ptr if use_rw else None, db.ptr, dRb.ptr if with_r else None, dX.ptr,
                                 dRX.ptr if use_rx else None, dS.ptr, dRS.ptr if with_r else None, rows, cols, n, activation))
        close(dS.numpy(), s.reshape(-1), f"S rw={use_rw} rx={use_rx}")
        if with_r:
            close(dRS.numpy(), rs.reshape(-1), f"RS rw={use_rw} rx={use_rx}")


@pytest.mark.parametrize("n,unit", [(4096, True), (300, False), (77, True)])
@pytest.mark.parametrize("with_r", [True, False])
def test_mlp_head_fusion_matches_unfused_steps_and_oracle(api, n, unit, with_r):
    """The 10-class head of the bench network runs its forward, the top sigmoid backward (on a given upstream, or on
    outGrad = 1 / outRGrad = 0 generated in the kernel) and its bias gradient in ONE kernel.  Against the oracle, and
    against the same plan with the streaming kernels switched off (tensor-core tiles + separate launches)."""
    sizes = [64, 96, 10]
    onet = ref.MLP(sizes, weight_scale=lambda k: 1 / np.sqrt(k))
    x = onet.make_input(n)
    up = None if unit else ref.splitmix_uniform(77, n * 10)
    upr = None if (unit or not with_r) else ref.splitmix_uniform(78, n * 10)
    if with_r:
        ores, ograd, orgrad = onet.step_r(x, n, up, upr)
    else:
        ores, ograd = onet.step(x, n, up)
    ctx = api.context()
    results = []
    for small_n in (1, 0):
        api.check(ctx.lib.af_set_option(ctx.h, b"small_n", small_n))
        try:
            plan = api.MLPPlan(sizes, n)
            for i, v in enumerate(onet.variables):
                plan.set_param(i, v.vector)
                plan.set_rvector(i, onet.rvec[v])
            for rep in range(2):
                out, rout, grads, rgrads = plan.step_host(x.vector, with_r=with_r, upstream=up, upstream_r=upr)
                close(out, ores.output(), f"Output small_n={small_n} rep={rep}")
                for i, v in enumerate(onet.variables):
                    close(grads[i], ograd[v], f"grad[{i}] small_n={small_n} rep={rep}")
                    if with_r:
                        close(rgrads[i], orgrad[v], f"rgrad[{i}] small_n={small_n} rep={rep}")
                if with_r:
                    close(rout, ores.routput(), f"ROutput small_n={small_n} rep={rep}")
            results.append(grads)
            plan.close()
        finally:
            api.check(ctx.lib.af_set_option(ctx.h, b"small_n", 1))
    for a, b_ in zip(*results):
        close(a, b_, "fused vs unfused head")


@pytest.mark.parametrize("sizes,n", [([130, 200, 96, 10], 300),     # ragged tiles: 200 = 3.1 x 64 columns, 300 = 4.7 x 64 rows
                                     ([64, 512, 10], 1000),          # the head's input gradient sums the 512 columns below it
                                     ([40, 66, 130, 34, 10], 257),   # three tensor-core input gradients in a row
                                     ([1000, 2000, 512, 10], 512)])  # the bench network
@pytest.mark.parametrize("with_r", [True, False])
def test_mlp_bias_gradient_fused_into_input_gradient_matches_separate_pass_and_oracle(api, sizes, n, with_r):
    """lin_add.go:94-104 batched is a column sum of the layer's upstream; the input-gradient launch that PRODUCES that
    upstream (sigmoid backward fused, math_funcs.go:365-371) adds the sums from its epilogue.  Against the oracle, and
    against the same plan with the fusion off (af_set_option "fused_colsum" 0: colsum_kernel passes), which must need
    more launches."""
    onet = ref.MLP(sizes, weight_scale=lambda k: 1 / np.sqrt(k))
    x = onet.make_input(n)
    if with_r:
        ores, ograd, orgrad = onet.step_r(x, n)
    else:
        ores, ograd = onet.step(x, n)
    ctx = api.context()
    results, launches = [], []
    for fused in (1, 0):
        ctx.set_option("fused_colsum", fused)
        try:
            plan = api.MLPPlan(sizes, n)
            for i, v in enumerate(onet.variables):
                plan.set_param(i, v.vector)
                plan.set_rvector(i, onet.rvec[v])
            for rep in range(2):  # accumulators are zeroed per step: the second step must give the same sums
                before = ctx.launch_count()
                out, rout, grads, rgrads = plan.step_host(x.vector, with_r=with_r)
                count = ctx.launch_count() - before
                close(out, ores.output(), f"Output fused={fused} rep={rep}")
                for i, v in enumerate(onet.variables):
                    close(grads[i], ograd[v], f"grad[{i}] fused={fused} rep={rep}")
                    if with_r:
                        close(rgrads[i], orgrad[v], f"rgrad[{i}] fused={fused} rep={rep}")
            results.append(grads)
            launches.append(count)
            plan.close()
        finally:
            ctx.set_option("fused_colsum", 1)
    for a, b_ in zip(*results):
        close(a, b_, "fused vs separate bias gradient")
    if n > 256:  # (batches <= 256 replay a recorded graph: the launch counter then counts the graph's kernels all the same)
        assert launches[0] < launches[1], f"fusion saved no launch: {launches}"


@pytest.mark.parametrize("rows,cols,n", [(2048, 30, 700), (50, 7, 130), (64, 32, 64), (3, 1, 17)])
@pytest.mark.parametrize("with_r,use_rw", [(True, True), (True, False), (False, False)])
@pytest.mark.parametrize("bias", [True, False])
def test_few_cols_forward_matches_oracle(api, rows, cols, n, with_r, use_rw, bias):
    """af_dense_fwd / af_lintran_fwd_r on a layer with <= 32 inputs, a large batch and no R-input take
    few_cols_fwd_kernel (the LSTM's x projection: I = 30): Y = X W^T (+ b), RY = X RW^T (+ Rb)."""
    ctx = api.context()
    L, h = ctx.lib, ctx.h
    W, RW, b, Rb, X = _rand(51, rows * cols), _rand(52, rows * cols), _rand(53, rows), _rand(54, rows), _rand(55, n * cols)
    y = X.reshape(n, cols) @ W.reshape(rows, cols).T + (b if bias else 0.0)
    ry = (X.reshape(n, cols) @ RW.reshape(rows, cols).T if use_rw else np.zeros((n, rows))) + (Rb if bias else 0.0)
    dW, dRW, db, dRb, dX = (ctx.upload(a) for a in (W, RW, b, Rb, X))
    dY, dRY = ctx.empty(n * rows), (ctx.empty(n * rows) if with_r else None)
    if bias:
        api.check(L.af_dense_fwd(h, dW.ptr, dRW.ptr if use_rw else None, db.ptr, dRb.ptr if with_r else None, dX.ptr, None,
                                 dY.ptr, dRY.ptr if with_r else None, rows, cols, n, 0))
    elif with_r:
        api.check(L.af_lintran_fwd_r(h, dW.ptr, dRW.ptr if use_rw else None, dX.ptr, None, dY.ptr, dRY.ptr, rows, cols, n))
    else:
        api.check(L.af_lintran_fwd(h, dW.ptr, dX.ptr, dY.ptr, rows, cols, n))
    close(dY.numpy(), y.reshape(-1), "Y")
    if with_r:
        close(dRY.numpy(), ry.reshape(-1), "RY")
