"""CPU tests of the C++ host layer (include/autofunc_b200.hpp): its graph logic -- gating by map keys, Constant()
pruning, nil Gradient, zero R for absent variables, accumulation, the temp ownership of upstreams, panic texts, the
Variable wire format -- exercised by the reference's own tests (tests/cpp/host_tests.cpp) against the CPU restatement
of the C ABI (oracle/abi_ref.cpp), and against the numpy oracle on seeded cases.  No GPU, no product arithmetic."""
import os
import re
import subprocess

import pytest

from oracle import autofunc_ref as ref
from tests import cpp_host


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    return cpp_host.build_host_tests(str(tmp_path_factory.mktemp("cpp_cpu")), cpp_host.build_ref_abi())


def test_reference_tests_pass_on_the_cpp_host_layer(exe):
    r = subprocess.run([exe, "selftest"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failed" in r.stdout and len(re.findall(r"^PASS ", r.stdout, re.M)) >= 46


@pytest.mark.parametrize("mode", ["ops", "fused", "plan", "nonr", "nilgrad"])
def test_cpp_mlp_matches_numpy_oracle(exe, tmp_path, mode):
    cpp_host.check_mlp(ref, exe, str(tmp_path), mode, [12, 20, 9, 4], 7)
    cpp_host.check_mlp(ref, exe, str(tmp_path), mode, [5, 3], 1, scaled=False, seed=3)


def test_cpp_lintran_and_arithmetic_match_numpy_oracle(exe, tmp_path):
    cpp_host.check_lintran(ref, exe, str(tmp_path), 6, 11, 10)
    cpp_host.check_lintran(ref, exe, str(tmp_path), 1, 1, 1, seed=5)
    cpp_host.check_arith(ref, exe, str(tmp_path), 4, 6, 3)


def test_cpp_pool_fold_matmul_and_cost_heads_match_numpy_oracle(exe, tmp_path):
    """SURVEY 8(f) rank 3 through the C++ host layer (MatMulVecs, Pool, Fold, Div, Pow, Inverse, Norm, Softmax, the
    log-domain sums, squared error): the temp-variable protocol and the gating, checked where no GPU exists."""
    cpp_host.check_next(ref, exe, str(tmp_path), 5, 7, 4)
    cpp_host.check_next(ref, exe, str(tmp_path), 1, 1, 1, seed=9)


def test_cpp_lstm_from_operators_matches_numpy_oracle(exe, tmp_path):
    """bench/lstm.go:139-196 built from the C++ host layer's operators (ConcatBatched, LinTran, LinAdd, Sigmoid, Mul,
    Add), batched over the lanes, BPTT + R-operator through the operator graph -- against the oracle's LSTMNet."""
    cpp_host.check_lstm(ref, exe, str(tmp_path), 3, 5, 2, 3, 2, case="lstmops")
    cpp_host.check_lstm(ref, exe, str(tmp_path), 6, 8, 4, 4, 3, case="lstmops")


def test_cpp_host_fails_loudly_without_the_gpu_library(tmp_path):
    """Linked against the PRODUCT library on a machine without a GPU the same binary must refuse to run: the host
    layer has no arithmetic of its own and the product has no CPU path."""
    import torch
    from autofunc_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("needs a machine without a GPU")
    exe = cpp_host.build_host_tests(str(tmp_path), _lib.LIB_PATH)
    r = subprocess.run([exe, "selftest", "TestLinTranOutput"], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and "FAIL TestLinTranOutput" in r.stdout and "cuda" in r.stdout.lower()


def test_product_never_references_the_cpu_restatement():
    pkg = os.path.join(cpp_host.ROOT, "autofunc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "abi_ref" not in text and "libautofunc_b200_ref" not in text, os.path.join(dirpath, f)
    hdr = open(os.path.join(cpp_host.INCLUDE, "autofunc_b200.hpp")).read()
    assert "abi_ref" not in hdr and "oracle" not in hdr
