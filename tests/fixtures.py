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
