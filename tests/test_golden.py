"""Golden vectors: every fixture and known answer the parity tests use is, digit for digit, the one in the
reference's own test sources (tests/golden/reference_fixtures.json, extracted by tests/golden/make_golden.py from
/root/reference/test/*.go -- the reference is pure Go and cannot be executed here), and the oracle reproduces every
known-answer vector those tests hold.  CPU only."""
import json
import os

import numpy as np

from oracle import autofunc_ref as ref
from tests import fixtures as fx

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_fixtures.json")))


def test_fixture_constants_match_reference_sources():
    lt, la, mf = G["lin_tran_test.go"], G["lin_add_test.go"], G["math_funcs_test.go"]
    ar, po, fo = G["arithmetic_test.go"], G["pool_test.go"], G["fold_test.go"]
    pairs = [
        (fx.LINTRAN_MAT1, lt["linTranTestMat1"]), (fx.LINTRAN_MAT2, lt["linTranTestMat2"]),
        (fx.LINTRAN_VEC, lt["linTranTestVec"]), (fx.LINTRAN_VEC2, lt["linTranTestVec2"]),
        (fx.LINTRAN_RV_MAT1, lt["rv_linTranTestMat1"]), (fx.LINTRAN_RV_MAT2, lt["rv_linTranTestMat2"]),
        (fx.LINTRAN_RV_VEC, lt["rv_linTranTestVec"]), (fx.LINTRAN_RV_VEC2, lt["rv_linTranTestVec2"]),
        (fx.LINTRAN_KAT_OUT, lt["TestLinTranOutput_expected"]),
        (fx.LINTRAN_KAT_BATCH_OUT, lt["TestLinTranBatchOutput_expected"]),
        (fx.LINTRAN_KAT_OUT, G["lin_alg_test.go"]["TestMatMulVecOutput_expected"]),
        (fx.LINADD_VEC1, la["linAddTestVec1"]), (fx.LINADD_VEC2, la["linAddTestVec2"]),
        (fx.LINADD_RV1, la["rv_linAddTestVec1"]), (fx.LINADD_RV2, la["rv_linAddTestVec2"]),
        (fx.MATH_VEC, mf["mathFuncTestVec"]), (fx.MATH_VEC_POS, mf["mathFuncTestVecPos"]),
        (fx.MATH_RV, mf["rv_mathFuncTestVec"]),
        (fx.ARITH_VEC1, ar["arithmeticTestVec1"]), (fx.ARITH_VEC2, ar["arithmeticTestVec2"]),
        (fx.ARITH_VEC3, ar["arithmeticTestVec3"]), (fx.ARITH_VEC4, ar["arithmeticTestVec4"]),
        (fx.ARITH_RV1, ar["rv_arithmeticTestVec1"]), (fx.ARITH_RV3, ar["rv_arithmeticTestVec3"]),
        (fx.ARITH_RV4, ar["rv_arithmeticTestVec4"]),
        (fx.POOL_VEC, po["poolTestVar"]), (fx.POOL_RV, po["rv_poolTestVar"]),
        (fx.FOLD_VEC, fo["input"]), (fx.FOLD_STATE, fo["init_state"]), (fx.FOLD_KAT_OUT, fo["TestFoldOut_expected"]),
    ]
    for i, (ours, theirs) in enumerate(pairs):
        assert [float(x) for x in ours] == theirs, f"fixture {i} differs from the reference's test source"
    assert G["functest/func_checker.go"]["DefaultDelta_DefaultPrec"] == [1e-5, 1e-5]
    assert G["variable.go"]["wire_format"].startswith("little-endian uint64")


def test_oracle_reproduces_every_known_answer():
    lt = G["lin_tran_test.go"]
    f = fx.LinTranFixture(ref)
    assert f.mat1.apply(f.vec).output().tolist() == lt["TestLinTranOutput_expected"]             # lin_tran_test.go:65-73
    assert f.batch_test.apply(f.vec2).output().tolist() == lt["TestLinTranBatchOutput_expected"]  # :85-97
    m = fx.matmulvec_test_func(ref, f.mat1.data)
    assert m.apply(f.vec).output().tolist() == G["lin_alg_test.go"]["TestMatMulVecOutput_expected"]
    fo = G["fold_test.go"]
    out = fx.fold_test_func(ref, ref.Variable(fo["init_state"])).apply(ref.Variable(fo["input"])).output()
    assert np.max(np.abs(out - np.array(fo["TestFoldOut_expected"]))) < 1e-5                       # fold_test.go:40
    v = ref.Variable(G["arithmetic_test.go"]["TestSumAllLogDomainOutput_input"])
    assert abs(ref.sum_all_log_domain(v).output()[0] - np.log(np.sum(np.exp(v.vector)))) < 1e-8   # arithmetic_test.go:103
    assert abs(ref.Norm().apply(ref.Variable(G["math_funcs_test.go"]["TestNormOutput_input"])).output()[0] - 5.0) < 1e-5
