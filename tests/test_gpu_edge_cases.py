"""Edge cases the reference's semantics define: empty vectors and batches, single elements, ragged / empty sequence
lanes, absent R-vectors, nil Gradient -- through the operator API and the C ABI, against the oracle."""
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
    gpu, want = np.asarray(gpu), np.asarray(want)
    assert gpu.shape == want.shape, f"{what}: {gpu.shape} vs {want.shape}"
    err = np.abs(gpu - want)
    assert not (err > ATOL + RTOL * np.abs(want)).any(), f"{what}: worst {err.max() if err.size else 0:.3e}"


def test_empty_batch_through_a_dense_network(api):
    """n = 0 samples: LinTran.Batch(in, 0) on an empty vector (lin_tran.go:52-54 accepts it: cols*0 == len(in)) gives an
    empty output and leaves the gradients untouched; the elementwise Batcher adaptor divides by n (batcher.go:40-42),
    so n = 0 is a panic there -- in Go and here."""
    w, b = api.Variable(ref.splitmix_uniform(1, 12)), api.Variable(ref.splitmix_uniform(2, 3))
    rv = {w: ref.splitmix_uniform(3, 12)}
    x = api.Variable(np.zeros(0))
    for net in (api.ComposedRBatcher([api.LinTran(w, 3, 4), api.LinAdd(b)]),
                api.fuse_dense([api.LinTran(w, 3, 4), api.LinAdd(b), api.Sigmoid()])):
        res = net.batch_r(rv, api.RVariable(x, rv), 0)
        assert res.output().shape == (0,) and res.routput().shape == (0,)
        g, rg = api.new_gradient([w, b]), api.new_rgradient([w, b])
        res.propagate_rgradient(np.zeros(0), np.zeros(0), rg, g)
        assert not g[w].any() and not g[b].any() and not rg[w].any()
        res2 = net.batch(x, 0)
        assert res2.output().shape == (0,)
        res2.propagate_gradient(np.zeros(0), g)
        assert not g[w].any()
    with pytest.raises(ZeroDivisionError):
        api.Sigmoid().batch(x, 0)


def test_empty_vectors_through_elementwise_ops_and_reductions(api):
    e = api.Variable(np.zeros(0))
    rv = {}
    for f in (api.Sigmoid(), api.Exp(), api.Log(), api.LogSigmoid(), api.Sin(), api.Tanh(), api.Softmax()):
        r = f.apply_r(rv, api.RVariable(e, rv))
        assert r.output().shape == (0,) and r.routput().shape == (0,)
        g, rg = api.new_gradient([e]), api.new_rgradient([e])
        r.propagate_rgradient(np.zeros(0), np.zeros(0), rg, g)
    assert api.sum_all(e).output().tolist() == [0.0]             # arithmetic.go:972-981: the empty sum
    assert api.squared_error(e, e).output().tolist() == [0.0]
    assert api.add(e, e).output().shape == (0,) and api.mul(e, e).output().shape == (0,)
    assert api.concat(e, e).output().shape == (0,)


def test_single_element_everything(api):
    """1 x 1 LinTran, one sample: the smallest non-empty problem takes the generic kernel (TMA needs >= 16 bytes rows)."""
    for ns in (ref, api):
        w, b, x = ns.Variable([0.75]), ns.Variable([-0.25]), ns.Variable([2.0])
        rv = {w: np.array([0.5]), x: np.array([-1.5])}
        net = ns.ComposedRFunc([ns.LinTran(w, 1, 1), ns.LinAdd(b), ns.Sigmoid()])
        res = net.apply_r(rv, ns.RVariable(x, rv))
        g, rg = ns.new_gradient([w, b, x]), ns.new_rgradient([w, b, x])
        res.propagate_rgradient(np.array([1.0]), np.array([0.25]), rg, g)
        got = [np.array(res.output()), np.array(res.routput())] + [g[v] for v in (w, b, x)] + [rg[v] for v in (w, b, x)]
        if ns is ref:
            want = got
    for i, (a, b_) in enumerate(zip(got, want)):
        close(a, b_, f"1x1 item {i}")


def test_nil_gradient_and_absent_rvector_prune_work(api):
    """PropagateRGradient(u, uR, rgrad, nil) (result.go:62-65) with a variable absent from the RVector
    (variable.go:85-91): only the R-gradient maps are filled, absent R terms vanish."""
    sizes, n = [12, 20, 6], 9
    onet = ref.MLP(sizes, weight_scale=lambda k: 1 / np.sqrt(k))
    x = onet.make_input(n)
    keep = list(onet.variables)[1:]          # W1 is absent from the RVector
    out = {}
    for ns in (ref, api):
        variables = [ns.Variable(v.vector) for v in onet.variables]
        rv = {nv: onet.rvec[ov].copy() for nv, ov in zip(variables, onet.variables) if ov in keep}
        layers = []
        for l in range(2):
            layers += [ns.LinTran(variables[2 * l], sizes[l + 1], sizes[l]), ns.LinAdd(variables[2 * l + 1]), ns.Sigmoid()]
        res = ns.ComposedRBatcher(layers).batch_r(rv, ns.RVariable(ns.Variable(x.vector), rv), n)
        rg = ns.new_rgradient(variables[2:])   # only the second layer's parameters are keys
        res.propagate_rgradient(np.ones(n * 6), np.zeros(n * 6), rg, None)
        out[ns] = [np.array(res.routput())] + [rg[v] for v in variables[2:]]
        assert set(rg) == set(variables[2:])
    for i, (w_, g_) in enumerate(zip(out[ref], out[api])):
        close(g_, w_, f"nil-grad item {i}")


def test_sequence_lists_with_only_empty_lanes(api):
    from autofunc_b200 import seq
    w = api.Variable(ref.splitmix_uniform(5, 16))
    lt = api.LinTran(w, 4, 4)
    for inp in (seq.ConstResult([[], [], []]), seq.PackedConstResult.from_host([[], [], []])):
        res = seq.MapBatcher(lt).apply_seqs(inp)
        assert res.output_seqs() == [[], [], []]
        g = api.new_gradient([w])
        res.propagate_gradient([[], [], []], g)
        assert not g[w].any()


def test_c_abi_rejects_bad_arguments_loudly(api):
    ctx = api.context()
    L, h = ctx.lib, ctx.h
    v = ctx.zeros(8)
    assert L.af_lintran_fwd(h, v.ptr, v.ptr, v.ptr, -1, 2, 2) == api._lib.AF_ESHAPE
    assert L.af_sigmoid_fwd(h, None, v.ptr, 8) == api._lib.AF_EINVAL and L.af_last_error()
    assert L.af_softmax_fwd_r(h, v.ptr, None, 0.0, v.ptr, None, -3, 2) == api._lib.AF_ESHAPE
    assert L.af_sigmoid_fwd(None, v.ptr, v.ptr, 8) == api._lib.AF_EINVAL
    assert L.af_mlp_step(None, 1, 1) == api._lib.AF_EINVAL


def test_plain_c_client_of_the_abi(api, tmp_path):
    """examples/mlp_step.c: a C99 program drives one bench iteration through the C ABI from host buffers and checks
    accumulation, linearity of the R-operator in the R-vector, and AddToVars on the device."""
    import subprocess
    from tests.test_abi_cpu import _compile_c_example
    r = subprocess.run([_compile_c_example(tmp_path)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert "mlp_step ok" in r.stdout
