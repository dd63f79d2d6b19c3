"""The reference's own test fixtures for the hot path, restated as data.

Sources (file:line in /root/reference): test/lin_tran_test.go:12-53, test/lin_add_test.go:11-19,
test/math_funcs_test.go:13-20.  `build(ns)` instantiates them against either API namespace
(oracle.autofunc_ref or autofunc_b200.api).
"""
LINTRAN_MAT1 = [1, 2, 3, 3, 4, 5, -6, 1, 7, 8, -10, 4]        # 3x4
LINTRAN_MAT2 = [4, 2, 3]                                         # 1x3
LINTRAN_VEC = [4, 3, 2, 1]
LINTRAN_VEC2 = [5, 1, -3, 2, 1, 2, 3, 4]
LINTRAN_RV_MAT1 = [0.40350, 0.75874, 0.87843, 0.86287, 0.97869, 0.27141, 0.84472, 0.43952,
                   0.49841, 0.46748, 0.71821, 0.57590]
LINTRAN_RV_MAT2 = [0.45823, 0.66692, 0.14662]
LINTRAN_RV_VEC = [0, -1, 3, -7]
LINTRAN_RV_VEC2 = [1, 3, -2, 3, -3, 2, 1, 7]
LINTRAN_KAT_OUT = [19, 20, 36]        # test/lin_tran_test.go:67
LINTRAN_KAT_BATCH_OUT = [349, 131]    # test/lin_tran_test.go:87

LINADD_VEC1 = [1, 3, 2, 1]
LINADD_VEC2 = [5, -3, 5, 4]
LINADD_RV1 = [0.5, -10, 5, 3.14]
LINADD_RV2 = [0, 5, -15, -30]

MATH_VEC = [1, -0.5, 0.3, 0.7]
MATH_RV = [0.5, -10, 5, 3.14]


class LinTranFixture:
    def __init__(self, ns):
        V = ns.Variable
        self.mat1 = ns.LinTran(V(LINTRAN_MAT1), 3, 4)
        self.mat2 = ns.LinTran(V(LINTRAN_MAT2), 1, 3)
        self.vec = V(LINTRAN_VEC)
        self.vec2 = V(LINTRAN_VEC2)
        self.variables = [self.mat1.data, self.mat2.data, self.vec, self.vec2]
        import numpy as np
        f = lambda x: np.array(x, dtype=np.float64)
        self.rv = {self.mat1.data: f(LINTRAN_RV_MAT1), self.mat2.data: f(LINTRAN_RV_MAT2),
                   self.vec: f(LINTRAN_RV_VEC), self.vec2: f(LINTRAN_RV_VEC2)}
        m1, m2 = self.mat1, self.mat2

        class BatchTest:  # test/lin_tran_test.go:55-63
            def apply(self, v):
                return m2.batch(m1.batch(v, 2), 2)

            def apply_r(self, rv, v):
                return m2.batch_r(rv, m1.batch_r(rv, v, 2), 2)
        self.batch_test = BatchTest()


# ---- the reference's own composite test functions, restated against `ns` ----
ARITH_VEC1 = [1, 2, -4, 3, -2]                                    # test/arithmetic_test.go:12-43
ARITH_VEC2 = [4, -2, 3, -0.5, 0.3]
ARITH_VEC3 = [0.99196, 0.48826, 0.51066, 0.66715, 0.44423]
ARITH_VEC4 = [0.509893, 0.874112, -0.080468, 0.372740, -0.172097]
ARITH_RV1 = [0.340162, -0.325063, 0.179612, 0.056463, -0.812274]
ARITH_RV3 = [0.59824, -0.63322, 0.13379, 0.99559, 0.53748]      # vector 2 intentionally absent (:33-34)
ARITH_RV4 = [4.2222, 5.2762, -7.5762, 2.3420, -1.8927]
MATH_VEC_POS = [1, 0.5, 0.3, 0.7]                                 # test/math_funcs_test.go:15
POOL_VEC = [1, -2, 3, -4, 5]                                      # test/pool_test.go:10-13
POOL_RV = [1, -1, 2, -2, 3]
FOLD_VEC = [1, 2, 2.5, 1.5, 3, 0.5]                               # test/fold_test.go:33-36
FOLD_STATE = [2, 1]
FOLD_KAT_OUT = [0.6, 2.0 / 3.0]                                   # test/fold_test.go:39


class ArithmeticFixture:
    """test/arithmetic_test.go:45-77 (arithmeticTestFunc + TestArithmetic)."""

    def __init__(self, ns):
        import numpy as np
        V = ns.Variable
        v1, v2, v3, v4 = V(ARITH_VEC1), V(ARITH_VEC2), V(ARITH_VEC3), V(ARITH_VEC4)
        self.variables = [v1, v2, v3, v4]
        self.input = v4
        f = lambda x: np.array(x, dtype=np.float64)
        self.rv = {v1: f(ARITH_RV1), v3: f(ARITH_RV3), v4: f(ARITH_RV4)}

        class F:
            def apply(self, r):
                sq1 = ns.square(ns.mul(ns.scale(ns.add(r, v1), -2), v2))
                sum1 = ns.add_scaler(ns.add(ns.mul(sq1, v3), ns.scale(v4, -0.5)), 2)
                powed = ns.div(ns.pow_(ns.pow_(ns.inverse(sum1), 2), 1 / 3.0), sum1)
                all_sum = ns.sum_all(ns.add_first(powed, v1))
                all_sum = ns.sub(all_sum, ns.sum_all_log_domain(v2))
                return ns.scale_first(ns.add_log_domain(v1, v2), all_sum)

            def apply_r(self, rv, r):
                r1, r2, r3, r4 = (ns.RVariable(v, rv) for v in (v1, v2, v3, v4))
                sq1 = ns.square_r(ns.mul_r(ns.scale_r(ns.add_r(r, r1), -2), r2))
                sum1 = ns.add_scaler_r(ns.add_r(ns.mul_r(sq1, r3), ns.scale_r(r4, -0.5)), 2)
                powed = ns.div_r(ns.pow_r(ns.pow_r(ns.inverse_r(sum1), 2), 1 / 3.0), sum1)
                all_sum = ns.sum_all_r(ns.add_first_r(powed, r1))
                all_sum = ns.sub_r(all_sum, ns.sum_all_log_domain_r(r2))
                return ns.scale_first_r(ns.add_log_domain_r(r1, r2), all_sum)
        self.func = F()


def pool_test_func(ns):
    """test/pool_test.go:15-40: ComposedRFunc{Sigmoid, Exp, poolTestFunc}."""
    class P:
        def apply(self, r):
            return ns.pool(r, lambda p: ns.pow_(ns.Exp().apply(ns.add_scaler(p, 2)), 0.5))

        def apply_r(self, rv, r):
            return ns.pool_r(r, lambda p: ns.pow_r(ns.Exp().apply_r(rv, ns.add_scaler_r(p, 2)), 0.5))
    return ns.ComposedRFunc([ns.Sigmoid(), ns.Exp(), P()])


def fold_test_func(ns, init_state):
    """test/fold_test.go:11-30."""
    class F:
        def apply(self, inp):
            ins = ns.split(ns.result_len(inp) // 2, inp)
            return ns.fold(init_state, ins, lambda s, i: ns.mul(ns.inverse(s), i))

        def apply_r(self, rv, inp):
            ins = ns.split_r(ns.result_len(inp) // 2, inp)
            return ns.fold_r(ns.RVariable(init_state, rv), ins, lambda s, i: ns.mul_r(ns.inverse_r(s), i))
    return F()


def matmulvec_test_func(ns, mat_var):
    """test/lin_alg_test.go:11-21."""
    class M:
        def apply(self, inp):
            return ns.mat_mul_vec(mat_var, 3, 4, inp)

        def apply_r(self, rv, inp):
            return ns.mat_mul_vec_r(ns.RVariable(mat_var, rv), 3, 4, inp)
    return M()


class RandomNetFixture:
    """A randomly shaped batched network chaining every family of node: LinTran/LinAdd/Sigmoid, MatMulVecs with a
    computed matrix, Pool, a batched Softmax head and the squared-error cost; b and t are absent from the RVector."""

    def __init__(self, ns, seed):
        import numpy as np
        rng = np.random.default_rng(seed)
        n, k, m = int(rng.integers(1, 4)), int(rng.integers(2, 5)), int(rng.integers(2, 5))
        w = ns.Variable(rng.uniform(-1, 1, m * k))
        b = ns.Variable(rng.uniform(-1, 1, m))
        q = ns.Variable(rng.uniform(0.5, 1.5, m * m))
        t = ns.Variable(rng.uniform(0, 1, n * m))
        x = ns.Variable(rng.uniform(-1, 1, n * k))
        self.variables, self.input = [w, b, q, t, x], x
        self.rv = {v: rng.uniform(-1, 1, len(v.vector)) for v in (w, q, x)}

        class Net:
            def apply(self, inp):
                h = ns.Sigmoid().apply(ns.LinAdd(b).batch(ns.LinTran(w, m, k).batch(inp, n), n))
                mixed = ns.mat_mul_vecs(ns.mul(q, q), m, m, h)
                p = ns.pool(mixed, lambda z: ns.add(z, ns.scale(z, 0.5)))
                return ns.squared_error(ns.Softmax(1.5).batch(p, n), t)

            def apply_r(self, r, inp):
                rb, rq, rt = (ns.RVariable(v, r) for v in (b, q, t))
                h = ns.Sigmoid().apply_r(r, ns.LinAdd(b).batch_r(r, ns.LinTran(w, m, k).batch_r(r, inp, n), n))
                mixed = ns.mat_mul_vecs_r(ns.mul_r(rq, rq), m, m, h)
                p = ns.pool_r(mixed, lambda z: ns.add_r(z, ns.scale_r(z, 0.5)))
                return ns.squared_error_r(ns.Softmax(1.5).batch_r(r, p, n), rt)
        self.func = Net()
