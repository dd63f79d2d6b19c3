"""CPU tests: pin the oracle against the reference's own golden vectors and acceptance
procedure.  No GPU, no product code."""
import numpy as np
import pytest

from oracle import autofunc_ref as ref
from oracle.functest import RFuncChecker
from tests import fixtures as fx


def test_lintran_output_kat():  # test/lin_tran_test.go:65-73
    f = fx.LinTranFixture(ref)
    assert np.array_equal(f.mat1.apply(f.vec).output(), fx.LINTRAN_KAT_OUT)


def test_lintran_batch_output_kat():  # test/lin_tran_test.go:85-97
    f = fx.LinTranFixture(ref)
    assert np.array_equal(f.batch_test.apply(f.vec2).output(), fx.LINTRAN_KAT_BATCH_OUT)


def test_lintran_checks():  # test/lin_tran_test.go:75-83
    f = fx.LinTranFixture(ref)
    c = RFuncChecker(ref, ref.ComposedRFunc([f.mat1, f.mat2]), f.variables, f.vec, f.rv)
    assert c.full_check() == []


def test_lintran_batch_checks():  # test/lin_tran_test.go:99-107
    f = fx.LinTranFixture(ref)
    c = RFuncChecker(ref, f.batch_test, f.variables, f.vec2, f.rv)
    assert c.full_check() == []


def test_linadd_checks():  # test/lin_add_test.go:21-30
    v1, v2 = ref.Variable(fx.LINADD_VEC1), ref.Variable(fx.LINADD_VEC2)
    rv = {v1: ref.vec(fx.LINADD_RV1), v2: ref.vec(fx.LINADD_RV2)}
    f = ref.ComposedRFunc([ref.LinAdd(v1), ref.LinAdd(v2)])
    assert RFuncChecker(ref, f, [v1, v2], v2, rv).full_check() == []


@pytest.mark.parametrize("make", [
    lambda: ref.Exp(),                                                 # test/math_funcs_test.go:22-30
    lambda: ref.SquaredNorm(),                                         # :42-50
    lambda: ref.ComposedRFunc([ref.Sigmoid(), ref.Sigmoid(), ref.Sigmoid()]),  # :52-60
])
def test_math_func_checks(make):
    v = ref.Variable(fx.MATH_VEC)
    rv = {v: ref.vec(fx.MATH_RV)}
    assert RFuncChecker(ref, make(), [v], v, rv).full_check() == []


def test_arithmetic_subset_checks():
    """Sub/Scale/Mul/Add/Square/SumAll on the hot path; RVector deliberately omits one variable
    as test/arithmetic_test.go:33-34 does."""
    a, b = ref.Variable([1, -2, 3, 0.5]), ref.Variable([0.3, 2, -1, 4])
    rv = {a: ref.vec([1, 2, -3, 0.25])}

    class F:
        def apply(self, x):
            d = ref.sub(x, b)
            return ref.scale(ref.sum_all(ref.square(ref.add(ref.mul(d, x), d))), 0.5)

        def apply_r(self, r, x):
            br = ref.RVariable(b, r)
            d = ref.sub_r(x, br)
            return ref.scale_r(ref.sum_all_r(ref.square_r(ref.add_r(ref.mul_r(d, x), d))), 0.5)
    assert RFuncChecker(ref, F(), [a, b], a, rv, prec=1e-4).full_check() == []


def test_batched_mlp_equals_per_sample():
    """Batcher contract (batcher.go:5-20; seqfunc/test/map_test.go:88-120): a packed batch
    equals the per-sample results concatenated, and gradients are the per-sample sums."""
    net = ref.MLP([7, 5, 3], weight_scale=lambda k: 1 / np.sqrt(k))
    n = 4
    x = net.make_input(n)
    res, grad, rgrad = net.step_r(x, n)
    outs, g2, rg2 = [], ref.new_gradient(net.variables), ref.new_gradient(net.variables)
    per = ref.ComposedRFunc(list(net.layers))
    for i in range(n):
        xi = ref.Variable(x.vector[i * 7:(i + 1) * 7])
        r = per.apply_r(net.rvec, ref.RVariable(xi, net.rvec))
        outs.append(r.output())
        r.propagate_rgradient(np.ones(3), np.zeros(3), rg2, g2)
    np.testing.assert_allclose(res.output(), np.concatenate(outs), rtol=1e-14)
    for v in net.variables:
        np.testing.assert_allclose(grad[v], g2[v], rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(rgrad[v], rg2[v], rtol=1e-12, atol=1e-15)


def test_rfunc_batcher_adaptor_matches_native_batch():
    """batcher.go:57-84: the naive adaptor gives the same values as the batched ops."""
    net = ref.MLP([6, 4, 2], weight_scale=lambda k: 1 / np.sqrt(k))
    n = 3
    x = net.make_input(n)
    res, grad, rgrad = net.step_r(x, n)
    naive = ref.ComposedRBatcher([l if isinstance(l, ref.LinTran) else ref.RFuncBatcher(l) for l in net.layers])
    r2 = naive.batch_r(net.rvec, ref.RVariable(x, net.rvec), n)
    g2, rg2 = ref.new_gradient(net.variables), ref.new_gradient(net.variables)
    r2.propagate_rgradient(np.ones(2 * n), np.zeros(2 * n), rg2, g2)
    np.testing.assert_allclose(res.output(), r2.output(), rtol=1e-14)
    np.testing.assert_allclose(res.routput(), r2.routput(), rtol=1e-13)
    for v in net.variables:
        np.testing.assert_allclose(grad[v], g2[v], rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(rgrad[v], rg2[v], rtol=1e-12, atol=1e-15)


def test_variable_wire_format_roundtrip():  # test/variable_test.go:12-32
    v = ref.Variable(ref.splitmix_uniform(7, 17))
    b = v.serialize()
    assert len(b) == 8 + 17 * 8 and b[:8] == (17).to_bytes(8, "little")
    assert np.array_equal(ref.Variable.deserialize(b).vector, v.vector)


# ---- the reference's tests for the callers / cost heads next to the path ----
def test_arithmetic_full_checks():  # test/arithmetic_test.go:69-77
    f = fx.ArithmeticFixture(ref)
    assert RFuncChecker(ref, f.func, f.variables, f.input, f.rv).full_check() == []


def test_add_log_domain_output():  # test/arithmetic_test.go:79-94
    a, b = ref.Variable(fx.ARITH_VEC1), ref.Variable(fx.ARITH_VEC2)
    np.testing.assert_allclose(ref.add_log_domain(a, b).output(), np.log(np.exp(a.vector) + np.exp(b.vector)), atol=1e-5)


def test_sum_all_log_domain_output():  # test/arithmetic_test.go:96-107
    v = ref.Variable([3, 3.5, -1, 2])
    assert abs(ref.sum_all_log_domain(v).output()[0] - np.log(np.sum(np.exp(v.vector)))) < 1e-8


@pytest.mark.parametrize("make,positive", [
    (lambda: ref.Log(), True),                                                          # test/math_funcs_test.go:32-40
    (lambda: ref.ComposedRFunc([ref.LogSigmoid(), ref.LogSigmoid(), ref.LogSigmoid()]), False),  # :62-70
    (lambda: ref.Softmax(), False),                                                     # :72-80
    (lambda: ref.Softmax(3), False),                                                    # :82-90
    (lambda: ref.Sin(), False),                                                         # :92-100
    (lambda: ref.Norm(), False),                                                        # :122-130
    (lambda: ref.Tanh(), False),                                                        # extension
])
def test_next_math_func_checks(make, positive):
    v = ref.Variable(fx.MATH_VEC)
    inp = ref.Variable(fx.MATH_VEC_POS) if positive else v
    rv = {v: ref.vec(fx.MATH_RV)}
    if positive:  # the reference passes an input that is not in Vars / RV for TestLog
        assert RFuncChecker(ref, make(), [v], inp, rv).full_check() == []
    else:
        assert RFuncChecker(ref, make(), [v], v, rv).full_check() == []


def test_cos_and_norm_outputs():  # test/math_funcs_test.go:102-120
    for arg in ref.splitmix_uniform(5, 10, -10, 10):
        assert abs(ref.Cos().apply(ref.Variable([arg])).output()[0] - np.cos(arg)) < 1e-5
    assert abs(ref.Norm().apply(ref.Variable([3, 4])).output()[0] - 5.0) < 1e-5


def test_pool_checks():  # test/pool_test.go:29-38
    v = ref.Variable(fx.POOL_VEC)
    assert RFuncChecker(ref, fx.pool_test_func(ref), [v], v, {v: ref.vec(fx.POOL_RV)}).full_check() == []


def test_fold_output_and_checks():  # test/fold_test.go:32-65
    v0, v1 = ref.Variable(fx.FOLD_VEC), ref.Variable(fx.FOLD_STATE)
    f = fx.fold_test_func(ref, v1)
    np.testing.assert_allclose(f.apply(v0).output(), fx.FOLD_KAT_OUT, atol=1e-5)
    rng = np.random.default_rng(0)
    rv = {v0: rng.standard_normal(6), v1: rng.standard_normal(2)}
    assert RFuncChecker(ref, f, [v0, v1], v0, rv).full_check() == []


def test_matmulvec_output_and_checks():  # test/lin_alg_test.go:55-73
    f = fx.LinTranFixture(ref)
    m = fx.matmulvec_test_func(ref, f.mat1.data)
    assert np.array_equal(m.apply(f.vec).output(), fx.LINTRAN_KAT_OUT)
    assert RFuncChecker(ref, m, f.variables, f.vec, f.rv).full_check() == []


def test_squared_error_is_half_squared_distance():
    a, y = ref.Variable([1, -2, 3]), ref.Variable([0.5, 0.5, 0.5])
    assert abs(ref.squared_error(a, y).output()[0] - 0.5 * np.sum((a.vector - y.vector) ** 2)) < 1e-15


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_networks_pass_the_reference_acceptance_procedure(seed):
    """The reference's checker (functest.RFuncChecker.FullCheck) on randomly shaped batched networks that chain every
    family of node the product implements (tests/fixtures.py: RandomNetFixture)."""
    f = fx.RandomNetFixture(ref, seed)
    assert RFuncChecker(ref, f.func, f.variables, f.input, f.rv).full_check() == []


@pytest.mark.parametrize("sizes,n,scaled", [([7, 5, 3], 4, False), ([100, 130, 66, 10], 37, True), ([1000, 2000, 512, 10], 16, False)])
def test_native_like_cpu_baseline_matches_the_numpy_oracle(sizes, n, scaled):
    """oracle/native_ref.cpp (SURVEY 8d "oracle-native": one blocked OpenMP Gemm per lin_tran.go call site, the CPU
    baseline that stands in for the reference's default gonum backend) computes the same step as the numpy oracle."""
    from oracle import native
    net = ref.MLP(sizes, weight_scale=(lambda k: 1 / np.sqrt(k)) if scaled else None)
    x = net.make_input(n)
    res, grad, rgrad = net.step_r(x, n)
    out, rout, grads, rgrads = native.mlp_step_r(sizes, n, [v.vector for v in net.variables],
                                                 [net.rvec[v] for v in net.variables], x.vector)

    def close(a, b):
        return bool(np.all(np.abs(a - b) <= 1e-12 + 1e-10 * np.abs(b)))
    assert close(out, res.output()) and close(rout, res.routput())
    for i, v in enumerate(net.variables):
        assert close(grads[i], grad[v]) and close(rgrads[i], rgrad[v]), i
