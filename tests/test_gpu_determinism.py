"""Run-to-run reproducibility (af_set_option "deterministic" 1).

The reference is a sequential Go program: the same inputs give the same bits every time.  The library's default
schedule combines split-K / stream-K partial tiles, column-sum partials and a few small-batch partial sums with
float64 atomics, so its last bits depend on arrival order (always inside 1e-12 + 1e-10 |ref|).  In deterministic mode
every sum's order is fixed by the shapes alone; these tests run the same call several times and demand BIT equality,
at the reference's own sizes (n = 1, n = 10), at the bench size (n = 4096) and for the LSTM -- and that the mode still
agrees with the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import autofunc_ref as ref

RTOL, ATOL = 1e-10, 1e-12
REPEATS = 4


@pytest.fixture(scope="module")
def api():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a B200"
    from autofunc_b200 import api as a
    a.context()
    return a


@pytest.fixture
def det(api):
    ctx = api.context()
    ctx.set_option("deterministic", 1)
    yield ctx
    ctx.set_option("deterministic", 0)


def close(gpu, want, what=""):
    gpu, want = np.asarray(gpu), np.asarray(want)
    assert gpu.shape == want.shape, what
    err = np.abs(gpu - want)
    bad = err > ATOL + RTOL * np.abs(want)
    assert not bad.any(), f"{what}: {bad.sum()} of {bad.size} beyond tolerance; worst {err.max():.3e}"


def same_bits(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, what
    diff = a.view(np.uint64) != b.view(np.uint64)
    assert not diff.any(), f"{what}: {diff.sum()} of {diff.size} elements differ between two runs (max {np.abs(a - b).max():.3e})"


def _flat(result):
    out, rout, grads, rgrads = result
    return [("out", out), ("rout", rout)] + [(f"grad[{i}]", g) for i, g in enumerate(grads)] + \
           [(f"rgrad[{i}]", g) for i, g in enumerate(rgrads)]


@pytest.mark.parametrize("n", [1, 4, 10, 96, 1024, 4096])
def test_mlp_step_is_bit_reproducible(api, det, n):
    """bench/mlp.go:43-57 on the bench network: n = 1 is the reference bench as written, 4096 the headline batch."""
    sizes = [1000, 2000, 512, 10]
    onet = ref.MLP(sizes, weight_scale=lambda k: 1 / np.sqrt(k))
    plan = api.MLPPlan(sizes, n)
    for i, v in enumerate(onet.variables):
        plan.set_param(i, v.vector)
        plan.set_rvector(i, onet.rvec[v])
    x = onet.make_input(n)
    first = plan.step_host(x.vector)
    for rep in range(1, REPEATS):  # (from the second call on, batches <= 256 replay the step's CUDA graph)
        again = plan.step_host(x.vector)
        for (name, a), (_, b) in zip(_flat(first), _flat(again)):
            same_bits(a, b, f"n={n} run {rep} {name}")
    if n <= 96:  # sizes the oracle finishes in seconds
        ores, ograd, orgrad = onet.step_r(x, n)
        close(first[0], ores.output(), "Output")
        close(first[1], ores.routput(), "ROutput")
        for i, v in enumerate(onet.variables):
            close(first[2][i], ograd[v], f"grad[{i}]")
            close(first[3][i], orgrad[v], f"rgrad[{i}]")
    plan.close()
    # the default schedule on the same plan shape agrees with the deterministic one inside the tolerance
    det.set_option("deterministic", 0)
    plan = api.MLPPlan(sizes, n)
    for i, v in enumerate(onet.variables):
        plan.set_param(i, v.vector)
        plan.set_rvector(i, onet.rvec[v])
    fast = plan.step_host(x.vector)
    for (name, a), (_, b) in zip(_flat(first), _flat(fast)):
        close(b, a, f"n={n} default vs deterministic {name}")
    plan.close()


def test_mlp_resident_bench_iteration_is_bit_reproducible(api, det):
    """The bench's own step (af_mlp_step_fresh_dp, unit upstream generated on the device) at n = 4096."""
    sizes, n = [1000, 2000, 512, 10], 4096
    onet = ref.MLP(sizes)
    plan = api.MLPPlan(sizes, n)
    for i, v in enumerate(onet.variables):
        plan.set_param(i, v.vector)
        plan.set_rvector(i, onet.rvec[v])
    xin = onet.make_input(n).vector
    api.check(det.lib.af_upload(det.h, plan.input().ptr, np.ascontiguousarray(xin).ctypes.data, xin.size))
    runs = []
    for _ in range(REPEATS):
        plan.step_fresh_dp(True)
        plan.dp_join()
        runs.append([plan.output().numpy(), plan.routput().numpy()] + [plan.grad(i).numpy() for i in range(6)] +
                    [plan.rgrad(i).numpy() for i in range(6)])
    for rep in range(1, REPEATS):
        for k, (a, b) in enumerate(zip(runs[0], runs[rep])):
            same_bits(a, b, f"run {rep} buffer {k}")
    plan.close()


@pytest.mark.parametrize("rows,cols,n", [(1024, 2048, 10),   # bench/lintran.go:12-16 (C1)
                                        (2000, 1000, 4096), (10, 512, 4096), (130, 66, 257)])
def test_lintran_entry_points_are_bit_reproducible(api, det, rows, cols, n):
    """af_lintran_{fwd,wgrad,igrad}_r one by one (G1-G6), accumulating into non-zero gradient buffers."""
    ctx, lib = det, det.lib
    W, RW = ref.splitmix_uniform(1, rows * cols), ref.splitmix_uniform(2, rows * cols)
    X, RX = ref.splitmix_uniform(3, n * cols), ref.splitmix_uniform(4, n * cols)
    U, UR = ref.splitmix_uniform(5, n * rows), ref.splitmix_uniform(6, n * rows)
    G0 = ref.splitmix_uniform(7, rows * cols)
    dev = {k: ctx.upload(v) for k, v in dict(W=W, RW=RW, X=X, RX=RX, U=U, UR=UR).items()}

    def run():
        Y, RY = ctx.zeros(n * rows), ctx.zeros(n * rows)
        gW, rgW = ctx.upload(G0), ctx.upload(G0)
        D, DR = ctx.zeros(n * cols), ctx.zeros(n * cols)
        api.check(lib.af_lintran_fwd_r(ctx.h, dev["W"].ptr, dev["RW"].ptr, dev["X"].ptr, dev["RX"].ptr, Y.ptr, RY.ptr, rows, cols, n))
        api.check(lib.af_lintran_wgrad_r(ctx.h, dev["U"].ptr, dev["UR"].ptr, dev["X"].ptr, dev["RX"].ptr, gW.ptr, rgW.ptr, rows, cols, n))
        api.check(lib.af_lintran_igrad_r(ctx.h, dev["U"].ptr, dev["UR"].ptr, dev["W"].ptr, dev["RW"].ptr, D.ptr, DR.ptr, rows, cols, n))
        return [v.numpy() for v in (Y, RY, gW, rgW, D, DR)]

    first = run()
    for rep in range(1, REPEATS):
        for name, a, b in zip(("Y", "RY", "gW", "rgW", "D", "DR"), first, run()):
            same_bits(a, b, f"{rows}x{cols} n={n} run {rep} {name}")
    if rows * cols * n <= 1024 * 2048 * 10:  # against the oracle's LinTran at sizes numpy finishes at once
        Wm, RWm, Xm, RXm, Um, URm = (a.reshape(s) for a, s in ((W, (rows, cols)), (RW, (rows, cols)), (X, (n, cols)),
                                                                (RX, (n, cols)), (U, (n, rows)), (UR, (n, rows))))
        close(first[0], (Xm @ Wm.T).ravel(), "Y")
        close(first[1], (RXm @ Wm.T + Xm @ RWm.T).ravel(), "RY")
        close(first[2], G0 + (Um.T @ Xm).ravel(), "gW")
        close(first[3], G0 + (URm.T @ Xm + Um.T @ RXm).ravel(), "rgW")
        close(first[4], (Um @ Wm).ravel(), "D")
        close(first[5], (URm @ Wm + Um @ RWm).ravel(), "DR")


@pytest.mark.parametrize("rows,cols,n", [(512, 2000, 4096),    # the bench network's layer 2: stream-K forward and weight gradient
                                        (2000, 1000, 4096),   # its layer 1: stream-K weight gradient (ragged 64-tiles both ways)
                                        (2000, 1000, 1024),   # n = 1024: stream-K forward and input gradient too
                                        (700, 530, 900)])     # ragged everywhere
@pytest.mark.parametrize("mode", ["deterministic", "sk_fixup_everywhere"])
def test_streamk_ordered_fixup_matches_oracle_and_repeats_bit_for_bit(api, rows, cols, n, mode):
    """Stream-K WITHOUT atomics (gemm_dmma_kernel SPECIAL = 3): the CTA that holds a split tile's first k-tiles fetches the
    other CTA's parked partial sums, adds them in one fixed order and runs the ordinary fused epilogue.  In deterministic
    mode it replaces the whole-tile launches wherever stream-K balances better; with af_set_option "sk_fixup" 2 it also
    replaces the default schedule's red.global.add.f64 launches.  Every LinTran entry point (G1-G6, `+=` into non-zero
    accumulators) and the fused dense layer (bias + sigmoid forward, sigmoid backward in the input gradient) against
    numpy, and bit equality between runs."""
    ctx = api.context()
    lib = ctx.lib
    if mode == "deterministic":
        ctx.set_option("deterministic", 1)
    else:
        ctx.set_option("sk_fixup", 2)
    try:
        W, RW = ref.splitmix_uniform(1, rows * cols) * 0.05, ref.splitmix_uniform(2, rows * cols)
        X, RX = ref.splitmix_uniform(3, n * cols), ref.splitmix_uniform(4, n * cols)
        U, UR = ref.splitmix_uniform(5, n * rows), ref.splitmix_uniform(6, n * rows)
        b, Rb = ref.splitmix_uniform(9, rows), ref.splitmix_uniform(10, rows)
        Sp, RSp = ref.splitmix_uniform(11, n * cols, 0.05, 0.95), ref.splitmix_uniform(12, n * cols)
        G0 = ref.splitmix_uniform(7, rows * cols)
        dev = {k: ctx.upload(v) for k, v in dict(W=W, RW=RW, X=X, RX=RX, U=U, UR=UR, b=b, Rb=Rb, Sp=Sp, RSp=RSp).items()}

        def run():
            Y, RY = ctx.zeros(n * rows), ctx.zeros(n * rows)
            S, RS = ctx.zeros(n * rows), ctx.zeros(n * rows)
            gW, rgW = ctx.upload(G0), ctx.upload(G0)
            D, DR = ctx.zeros(n * cols), ctx.zeros(n * cols)
            E, ER = ctx.zeros(n * cols), ctx.zeros(n * cols)
            api.check(lib.af_lintran_fwd_r(ctx.h, dev["W"].ptr, dev["RW"].ptr, dev["X"].ptr, dev["RX"].ptr, Y.ptr, RY.ptr, rows, cols, n))
            api.check(lib.af_dense_fwd(ctx.h, dev["W"].ptr, dev["RW"].ptr, dev["b"].ptr, dev["Rb"].ptr, dev["X"].ptr, dev["RX"].ptr,
                                       S.ptr, RS.ptr, rows, cols, n, 1))
            api.check(lib.af_lintran_wgrad_r(ctx.h, dev["U"].ptr, dev["UR"].ptr, dev["X"].ptr, dev["RX"].ptr, gW.ptr, rgW.ptr, rows, cols, n))
            api.check(lib.af_lintran_igrad_r(ctx.h, dev["U"].ptr, dev["UR"].ptr, dev["W"].ptr, dev["RW"].ptr, D.ptr, DR.ptr, rows, cols, n))
            api.check(lib.af_dense_igrad(ctx.h, dev["U"].ptr, dev["UR"].ptr, dev["W"].ptr, dev["RW"].ptr, dev["Sp"].ptr, dev["RSp"].ptr,
                                         E.ptr, ER.ptr, rows, cols, n))
            return [v.numpy() for v in (Y, RY, S, RS, gW, rgW, D, DR, E, ER)]

        names = ("Y", "RY", "S", "RS", "gW", "rgW", "D", "DR", "E", "ER")
        first = run()
        for rep in range(1, 3):
            for name, a, c in zip(names, first, run()):
                if mode == "deterministic":  # (elsewhere the launches that stream-K does not take may still use atomics)
                    same_bits(a, c, f"{mode} {rows}x{cols} n={n} run {rep} {name}")
                else:
                    close(c, a, f"{mode} {rows}x{cols} n={n} run {rep} {name}")
        Wm, RWm, Xm, RXm, Um, URm = (a.reshape(s) for a, s in ((W, (rows, cols)), (RW, (rows, cols)), (X, (n, cols)),
                                                                (RX, (n, cols)), (U, (n, rows)), (UR, (n, rows))))
        y, ry = Xm @ Wm.T, RXm @ Wm.T + Xm @ RWm.T
        close(first[0], y.ravel(), "Y")
        close(first[1], ry.ravel(), "RY")
        sg = 1.0 / (1.0 + np.exp(-(y + b)))
        close(first[2], sg.ravel(), "S = sigmoid(XW^T + b)")
        close(first[3], (sg * (1 - sg) * (ry + Rb)).ravel(), "RS")
        close(first[4], G0 + (Um.T @ Xm).ravel(), "gW")
        close(first[5], G0 + (URm.T @ Xm + Um.T @ RXm).ravel(), "rgW")
        d, dr = Um @ Wm, URm @ Wm + Um @ RWm
        close(first[6], d.ravel(), "D")
        close(first[7], dr.ravel(), "DR")
        s_, rs_ = Sp.reshape(n, cols), RSp.reshape(n, cols)
        close(first[8], (s_ * (1 - s_) * d).ravel(), "E = sigmoid'(S) D")
        close(first[9], (rs_ * (1 - s_) * d - s_ * rs_ * d + s_ * (1 - s_) * dr).ravel(), "ER")
    finally:
        ctx.set_option("deterministic", 0)
        ctx.set_option("sk_fixup", 1)


@pytest.mark.parametrize("I,H,O,T,B,fused", [(16, 64, 10, 6, 70, 1), (16, 64, 10, 6, 70, 0), (30, 512, 10, 128, 256, 1)])
def test_lstm_is_bit_reproducible(api, det, I, H, O, T, B, fused):
    """bench/lstm.go forward + BPTT + R through af_lstm_*: fused time steps and the per-kernel path."""
    from autofunc_b200.lstm import LSTMPlan
    det.set_option("lstm_fused", fused)
    try:
        plan = LSTMPlan(I, H, O, T, B)
        for i in range(10):
            plan.set_param(i, ref.splitmix_uniform(900 + i, plan.param_len(i)) * (0.05 if i % 2 == 0 else 1.0))
            plan.set_rvector(i, ref.splitmix_uniform(950 + i, plan.param_len(i)))
        x = ref.splitmix_uniform(11, T * B * I, 0, 1)
        up = ref.splitmix_uniform(12, T * B * O, 0, 1)
        upr = ref.splitmix_uniform(13, T * B * O)
        first = plan.run_host(x, up, upr)
        for rep in range(1, 3):
            again = plan.run_host(x, up, upr)
            for (name, a), (_, b) in zip(_flat(first), _flat(again)):
                same_bits(a, b, f"run {rep} {name}")
        plan.close()
    finally:
        det.set_option("lstm_fused", 1)


def test_pipelined_host_steps_equal_blocking_steps_bit_for_bit(api, det):
    """af_mlp_submit_host / af_mlp_wait against af_mlp_step_host, with bit equality (the default-mode version of this test,
    tests/test_gpu_parity.py, can only compare inside the tolerance)."""
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
        same_bits(bufs[k]["out"], out, f"step {k} Output")
        same_bits(bufs[k]["rout"], rout, f"step {k} ROutput")
        for i in range(P):
            same_bits(bufs[k]["grads"][i], grads[i], f"step {k} grad[{i}]")
            same_bits(bufs[k]["rgrads"][i], rgrads[i], f"step {k} rgrad[{i}]")
    plan.close()
