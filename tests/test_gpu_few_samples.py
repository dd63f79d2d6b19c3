"""LinTran with a few samples (9..16; the reference's LinTranBenchmark runs n = 10, bench/lintran.go:12-16):
the streaming tensor-core kernels of csrc/few_samples.cu and the one-call backward af_lintran_backward(_r)
(linTranRResult.PropagateRGradient, lin_tran.go:393-425) against the oracle's LinTran, entry point by entry point."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import autofunc_ref as ref

RTOL, ATOL = 1e-10, 1e-12


@pytest.fixture(scope="module")
def api():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a B200"
    from autofunc_b200 import api as a
    a.context()
    return a


def close(gpu, want, what=""):
    gpu, want = np.asarray(gpu), np.asarray(want)
    assert gpu.shape == want.shape, f"{what}: shape {gpu.shape} vs {want.shape}"
    err = np.abs(gpu - want)
    bad = err > ATOL + RTOL * np.abs(want)
    assert not bad.any(), f"{what}: {bad.sum()} of {bad.size} beyond 1e-12+1e-10|ref|; worst err {err.max():.3e}"


def _oracle(rows, cols, n, W, RW, X, RX, U, UR):
    """The six products by the oracle's LinTran (lin_tran.go restated: oracle/autofunc_ref.py)."""
    wv = ref.Variable(W)
    lt = ref.LinTran(wv, rows, cols)
    rdata = ref.RVariable(wv, {wv: RW})

    class RIn:
        def output(self): return X
        def routput(self): return RX
    y, ry = lt.multiply(X), lt.multiply_r(rdata, RIn())
    g, rg = {wv: np.zeros(rows * cols)}, np.zeros(rows * cols)
    lt.data_gradient(U, X, g)
    lt.data_gradient_r(U, UR, X, RX, rg)
    return y, ry, g[wv], rg, lt.input_gradient(U), lt.input_gradient_r(rdata, U, UR)


SHAPES = [(1024, 2048, 10),  # C1
          (2000, 1000, 16), (512, 2000, 9), (10, 512, 12), (130, 66, 13), (37, 258, 11), (8, 2, 16), (2048, 8192, 10)]


@pytest.mark.parametrize("rows,cols,n", SHAPES)
@pytest.mark.parametrize("few_samples", [1, 0])
def test_backward_entry_points_match_oracle(api, rows, cols, n, few_samples):
    ctx = api.context()
    ctx.set_option("few_samples", few_samples)
    try:
        W, RW = ref.splitmix_uniform(1, rows * cols), ref.splitmix_uniform(2, rows * cols)
        X, RX = ref.splitmix_uniform(3, n * cols), ref.splitmix_uniform(4, n * cols)
        U, UR = ref.splitmix_uniform(5, n * rows), ref.splitmix_uniform(6, n * rows)
        y, ry, g, rg, d, dr = _oracle(rows, cols, n, W, RW, X, RX, U, UR)
        g0, rg0 = ref.splitmix_uniform(7, rows * cols), ref.splitmix_uniform(8, rows * cols)
        d0, dr0 = ref.splitmix_uniform(9, n * cols), ref.splitmix_uniform(10, n * cols)
        L, h = ctx.lib, ctx.h
        dW, dRW, dX, dRX, dU, dUR = (ctx.upload(a) for a in (W, RW, X, RX, U, UR))

        # forward (G1 + G2)
        dY, dRY = ctx.empty(n * rows), ctx.empty(n * rows)
        api.check(L.af_lintran_fwd_r(h, dW.ptr, dRW.ptr, dX.ptr, dRX.ptr, dY.ptr, dRY.ptr, rows, cols, n))
        close(dY.numpy(), y, "Y")
        close(dRY.numpy(), ry, "RY")
        dY1 = ctx.empty(n * rows)
        api.check(L.af_lintran_fwd(h, dW.ptr, dX.ptr, dY1.ptr, rows, cols, n))
        close(dY1.numpy(), y, "Y (plain)")

        # all four gradients in one call, store semantics for the input gradient; accumulators start non-zero
        dG, dRG, dD, dDR = ctx.upload(g0), ctx.upload(rg0), ctx.upload(d0), ctx.upload(dr0)
        api.check(L.af_lintran_backward_r(h, dU.ptr, dUR.ptr, dW.ptr, dRW.ptr, dX.ptr, dRX.ptr, dG.ptr, dRG.ptr, dD.ptr, dDR.ptr,
                                          rows, cols, n, 0))
        close(dG.numpy(), g0 + g, "gW")
        close(dRG.numpy(), rg0 + rg, "rgW")
        close(dD.numpy(), d, "D (store)")
        close(dDR.numpy(), dr, "DR (store)")
        # ... and adding into the input's accumulators (variable.go:113-123)
        dG, dRG, dD, dDR = ctx.upload(g0), ctx.upload(rg0), ctx.upload(d0), ctx.upload(dr0)
        api.check(L.af_lintran_backward_r(h, dU.ptr, dUR.ptr, dW.ptr, dRW.ptr, dX.ptr, dRX.ptr, dG.ptr, dRG.ptr, dD.ptr, dDR.ptr,
                                          rows, cols, n, 1))
        close(dG.numpy(), g0 + g, "gW (2)")
        close(dRG.numpy(), rg0 + rg, "rgW (2)")
        close(dD.numpy(), d0 + d, "D (accumulate)")
        close(dDR.numpy(), dr0 + dr, "DR (accumulate)")
        # W not in grad (gW NULL), input constant (D, DR NULL), R-vectors absent (RW, RX NULL == zero, variable.go:85-91)
        dRG, dD, dDR = ctx.upload(rg0), ctx.empty(n * cols), ctx.empty(n * cols)
        api.check(L.af_lintran_backward_r(h, dU.ptr, dUR.ptr, dW.ptr, None, dX.ptr, None, None, dRG.ptr, dD.ptr, dDR.ptr,
                                          rows, cols, n, 0))
        Um, URm, Xm, Wm = U.reshape(n, rows), UR.reshape(n, rows), X.reshape(n, cols), W.reshape(rows, cols)
        close(dRG.numpy(), rg0 + (URm.T @ Xm).ravel(), "rgW with RX = 0")
        close(dD.numpy(), d, "D with RW = 0")
        close(dDR.numpy(), (URm @ Wm).ravel(), "DR with RW = 0")
        dG = ctx.upload(g0)
        api.check(L.af_lintran_backward_r(h, dU.ptr, dUR.ptr, None, None, dX.ptr, dRX.ptr, dG.ptr, None, None, None, rows, cols, n, 0))
        close(dG.numpy(), g0 + g, "gW alone")
        # plain (non-R) form
        dG, dD = ctx.upload(g0), ctx.upload(d0)
        api.check(L.af_lintran_backward(h, dU.ptr, dW.ptr, dX.ptr, dG.ptr, dD.ptr, rows, cols, n, 1))
        close(dG.numpy(), g0 + g, "gW (plain)")
        close(dD.numpy(), d0 + d, "D (plain, accumulate)")
        # the separate entry points route through the same kernels
        dG, dRG, dD, dDR = ctx.upload(g0), ctx.upload(rg0), ctx.empty(n * cols), ctx.empty(n * cols)
        api.check(L.af_lintran_wgrad_r(h, dU.ptr, dUR.ptr, dX.ptr, dRX.ptr, dG.ptr, dRG.ptr, rows, cols, n))
        api.check(L.af_lintran_igrad_r(h, dU.ptr, dUR.ptr, dW.ptr, dRW.ptr, dD.ptr, dDR.ptr, rows, cols, n))
        close(dG.numpy(), g0 + g, "gW (wgrad_r)")
        close(dRG.numpy(), rg0 + rg, "rgW (wgrad_r)")
        close(dD.numpy(), d, "D (igrad_r)")
        close(dDR.numpy(), dr, "DR (igrad_r)")
    finally:
        ctx.set_option("few_samples", 1)


@pytest.mark.parametrize("n", [1, 3, 8, 10, 16, 17, 64])
def test_backward_entry_point_at_every_batch_size_class(api, n):
    """The one-call backward takes the few-samples kernel only for 9..16 samples; every other size runs the two passes."""
    rows, cols = 96, 130
    ctx = api.context()
    W, RW = ref.splitmix_uniform(1, rows * cols), ref.splitmix_uniform(2, rows * cols)
    X, RX = ref.splitmix_uniform(3, n * cols), ref.splitmix_uniform(4, n * cols)
    U, UR = ref.splitmix_uniform(5, n * rows), ref.splitmix_uniform(6, n * rows)
    _, _, g, rg, d, dr = _oracle(rows, cols, n, W, RW, X, RX, U, UR)
    d0 = ref.splitmix_uniform(9, n * cols)
    dW, dRW, dX, dRX, dU, dUR = (ctx.upload(a) for a in (W, RW, X, RX, U, UR))
    for acc in (0, 1):
        dG, dRG, dD, dDR = ctx.zeros(rows * cols), ctx.zeros(rows * cols), ctx.upload(d0), ctx.upload(d0)
        api.check(ctx.lib.af_lintran_backward_r(ctx.h, dU.ptr, dUR.ptr, dW.ptr, dRW.ptr, dX.ptr, dRX.ptr, dG.ptr, dRG.ptr,
                                                dD.ptr, dDR.ptr, rows, cols, n, acc))
        close(dG.numpy(), g, f"n={n} gW")
        close(dRG.numpy(), rg, f"n={n} rgW")
        close(dD.numpy(), acc * d0 + d, f"n={n} D acc={acc}")
        close(dDR.numpy(), acc * d0 + dr, f"n={n} DR acc={acc}")


@pytest.mark.parametrize("n", [9, 10, 16])
@pytest.mark.parametrize("activation", [0, 1])
def test_dense_layer_with_few_samples_matches_oracle(api, n, activation):
    """af_dense_fwd (bias and sigmoid applied by the block that folds the K splits) and af_dense_igrad (sigmoid backward
    of the layer below in the fold) on the bench network's first layer."""
    rows, cols = 2000, 1000
    ctx = api.context()
    W, RW = ref.splitmix_uniform(1, rows * cols) / np.sqrt(cols), ref.splitmix_uniform(2, rows * cols)
    b, Rb = ref.splitmix_uniform(11, rows), ref.splitmix_uniform(12, rows)
    X, RX = ref.splitmix_uniform(3, n * cols), ref.splitmix_uniform(4, n * cols)
    Xm, RXm, Wm, RWm = X.reshape(n, cols), RX.reshape(n, cols), W.reshape(rows, cols), RW.reshape(rows, cols)
    z = Xm @ Wm.T + b
    rz = RXm @ Wm.T + Xm @ RWm.T + Rb
    if activation:
        s = 1 / (1 + np.exp(-z))
        want, rwant = s, s * (1 - s) * rz
    else:
        want, rwant = z, rz
    dS, dRS = ctx.empty(n * rows), ctx.empty(n * rows)
    up = [ctx.upload(a) for a in (W, RW, b, Rb, X, RX)]
    api.check(ctx.lib.af_dense_fwd(ctx.h, up[0].ptr, up[1].ptr, up[2].ptr, up[3].ptr, up[4].ptr, up[5].ptr, dS.ptr, dRS.ptr,
                                   rows, cols, n, activation))
    close(dS.numpy(), want.ravel(), "S")
    close(dRS.numpy(), rwant.ravel(), "RS")
    # input gradient with the sigmoid backward of the layer below (activations A, RA of width cols) fused
    U, UR = ref.splitmix_uniform(5, n * rows), ref.splitmix_uniform(6, n * rows)
    A = 1 / (1 + np.exp(-ref.splitmix_uniform(13, n * cols)))
    RA = ref.splitmix_uniform(14, n * cols)
    Um, URm = U.reshape(n, rows), UR.reshape(n, rows)
    d, dr = Um @ Wm, URm @ Wm + Um @ RWm
    Am, RAm = A.reshape(n, cols), RA.reshape(n, cols)
    want_d = Am * (1 - Am) * d
    want_dr = RAm * (1 - Am) * d - Am * RAm * d + Am * (1 - Am) * dr
    dD, dDR = ctx.empty(n * cols), ctx.empty(n * cols)
    dU, dUR, dA, dRA = (ctx.upload(a) for a in (U, UR, A, RA))
    api.check(ctx.lib.af_dense_igrad(ctx.h, dU.ptr, dUR.ptr, up[0].ptr, up[1].ptr, dA.ptr, dRA.ptr, dD.ptr, dDR.ptr, rows, cols, n))
    close(dD.numpy(), want_d.ravel(), "D through sigmoid backward")
    close(dDR.numpy(), want_dr.ravel(), "DR through sigmoid backward")


def test_lintran_benchmark_step_through_the_operator_api(api):
    """bench/lintran.go:52-91 (RunR with backProp) through the mirror of the reference's operators: BatchR, then
    PropagateRGradient with W and the input both in the maps -- the path that issues af_lintran_backward_r with
    accumulate_input = 1 -- against the oracle, twice (the maps accumulate)."""
    rows, cols, n = 256, 384, 10
    W, X = ref.splitmix_uniform(1, rows * cols), ref.splitmix_uniform(3, n * cols)
    RW, RX = ref.splitmix_uniform(2, rows * cols), ref.splitmix_uniform(4, n * cols)
    U, UR = ref.splitmix_uniform(5, n * rows), ref.splitmix_uniform(6, n * rows)
    results = []
    for mod in (ref, api):
        wv, xv = mod.Variable(W.copy()), mod.Variable(X.copy())
        lt = mod.LinTran(wv, rows, cols)
        rv = {wv: RW.copy(), xv: RX.copy()}
        grad, rgrad = mod.new_gradient([wv, xv]), mod.new_rgradient([wv, xv])
        for _ in range(2):
            res = lt.batch_r(rv, mod.RVariable(xv, rv), n)
            res.propagate_rgradient(U.copy(), UR.copy(), rgrad, grad)
        plain = mod.new_gradient([wv, xv])
        lt.batch(xv, n).propagate_gradient(U.copy(), plain)
        results.append((np.asarray(res.output()), np.asarray(res.routput()), grad[wv], grad[xv], rgrad[wv], rgrad[xv],
                        plain[wv], plain[xv]))
    for name, want, got in zip(("Output", "ROutput", "grad[W]", "grad[X]", "rgrad[W]", "rgrad[X]", "plain grad[W]", "plain grad[X]"),
                               results[0], results[1]):
        close(got, want, name)


# ---- the <= 16-row weight gradient over a large batch (few_rows_wgrad_kernel: the 10-class head of bench/mlp.go) ----
@pytest.mark.parametrize("rows,cols,n", [(10, 512, 4096),   # the bench network's head at the headline batch
                                        (1, 2, 64), (8, 66, 257), (12, 130, 1000), (16, 1000, 333), (9, 4098, 70)])
@pytest.mark.parametrize("variant", ["dual", "no_rx", "single"])
def test_few_rows_weight_gradient_matches_oracle_and_repeats_bit_for_bit(api, rows, cols, n, variant):
    """dataGradient / dataGradientR (lin_tran.go:179-270) with few output rows, `+=` into non-zero accumulators; the
    kernel's sums are ordered (warp order, then chunk order), so two runs give the same bits."""
    ctx = api.context()
    ctx.set_option("few_rows_wgrad", 2)
    try:
        X, RX = ref.splitmix_uniform(3, n * cols), ref.splitmix_uniform(4, n * cols)
        U, UR = ref.splitmix_uniform(5, n * rows), ref.splitmix_uniform(6, n * rows)
        G0, RG0 = ref.splitmix_uniform(7, rows * cols), ref.splitmix_uniform(8, rows * cols)
        wv = ref.Variable(np.zeros(rows * cols))
        lt = ref.LinTran(wv, rows, cols)
        g, rg = {wv: G0.copy()}, RG0.copy()
        lt.data_gradient(U, X, g)
        lt.data_gradient_r(U, UR, X, RX if variant == "dual" else np.zeros_like(X), rg)
        L, h = ctx.lib, ctx.h
        dX, dRX, dU, dUR = (ctx.upload(a) for a in (X, RX, U, UR))
        runs = []
        for _ in range(3):
            dG, dRG = ctx.upload(G0), ctx.upload(RG0)
            before = ctx.launch_count()
            if variant == "single":
                api.check(L.af_lintran_wgrad(h, dU.ptr, dX.ptr, dG.ptr, rows, cols, n))
            else:
                api.check(L.af_lintran_wgrad_r(h, dU.ptr, dUR.ptr, dX.ptr, dRX.ptr if variant == "dual" else None,
                                               dG.ptr, dRG.ptr, rows, cols, n))
            assert ctx.launch_count() - before == 1, "one launch: no zeroing, no separate fold kernel"
            runs.append((dG.numpy(), dRG.numpy()))
        close(runs[0][0], g[wv], "G3 dataGradient")
        if variant != "single":
            close(runs[0][1], rg, "G4 dataGradientR")
        for a, b in runs[1:]:
            assert np.array_equal(a.view(np.uint64), runs[0][0].view(np.uint64))
            assert np.array_equal(b.view(np.uint64), runs[0][1].view(np.uint64))
    finally:
        ctx.set_option("few_rows_wgrad", 1)
