"""GPU tests of the C++ host side of the drop-in (include/autofunc_b200.hpp over the C ABI): the reference's own Go
tests restated in C++ (tests/cpp/host_tests.cpp) run on the B200, and seeded cases are compared with the CPU oracle at
|gpu - oracle| <= 1e-12 + 1e-10 |oracle|."""
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu

from oracle import autofunc_ref as ref
from tests import cpp_host


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a B200"
    from autofunc_b200 import _lib
    _lib.load()
    return cpp_host.build_host_tests(str(tmp_path_factory.mktemp("cpp_gpu")), _lib.LIB_PATH)


def test_reference_tests_pass_on_the_gpu_through_cpp(exe):
    r = subprocess.run([exe, "selftest"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failed" in r.stdout and len(re.findall(r"^PASS ", r.stdout, re.M)) >= 46
    launches = int(re.search(r"(\d+) kernel launches", r.stdout).group(1))
    assert launches > 1000  # the work ran as CUDA kernels of this library


@pytest.mark.parametrize("mode", ["ops", "fused", "plan", "nonr", "nilgrad"])
def test_cpp_mlp_matches_oracle(exe, tmp_path, mode):
    cpp_host.check_mlp(ref, exe, str(tmp_path), mode, [64, 96, 32, 10], 48)
    cpp_host.check_mlp(ref, exe, str(tmp_path), mode, [100, 200, 52, 10], 130, scaled=False, seed=1)  # saturating init
    cpp_host.check_mlp(ref, exe, str(tmp_path), mode, [5, 3], 1, seed=3)


def test_cpp_lintran_c1_shape_matches_oracle(exe, tmp_path):
    """BASELINE config 1 (bench/lintran.go:13-17: 1024 x 2048, n = 10) through the C++ LinTran."""
    cpp_host.check_lintran(ref, exe, str(tmp_path), 1024, 2048, 10)
    cpp_host.check_lintran(ref, exe, str(tmp_path), 7, 13, 3, seed=2)   # odd pitches: the generic kernel


def test_cpp_arithmetic_and_slices_match_oracle(exe, tmp_path):
    cpp_host.check_arith(ref, exe, str(tmp_path), 33, 70, 40)


def test_cpp_pool_fold_matmul_and_cost_heads_match_oracle(exe, tmp_path):
    """SURVEY 8(f) rank 3 through the C++ host layer on the device: MatMulVecs with a computed matrix (the six LinTran
    kernels), Pool / Fold with the temporaries in HBM, Softmax rows, Norm, squared error, the log-domain sums."""
    cpp_host.check_next(ref, exe, str(tmp_path), 24, 40, 16)
    cpp_host.check_next(ref, exe, str(tmp_path), 10, 64, 130, seed=4)
    cpp_host.check_next(ref, exe, str(tmp_path), 3, 5, 2, seed=7)   # odd pitches: the generic contraction kernel


def test_cpp_lstm_from_operators_matches_oracle(exe, tmp_path):
    """bench/lstm.go:139-196 assembled from the C++ operators on the device (the reference's benchmark network running on
    the drop-in's LinTran / LinAdd / Sigmoid / Mul / Add / Concat), against the oracle -- the same files the fused plan
    is checked with below."""
    cpp_host.check_lstm(ref, exe, str(tmp_path), 6, 8, 4, 4, 3, case="lstmops")
    cpp_host.check_lstm(ref, exe, str(tmp_path), 30, 100, 10, 5, 3, scale=1.0, case="lstmops")


def test_cpp_lstm_plan_matches_oracle(exe, tmp_path):
    cpp_host.check_lstm(ref, exe, str(tmp_path), 6, 8, 4, 4, 3)
    cpp_host.check_lstm(ref, exe, str(tmp_path), 30, 100, 10, 5, 3, scale=1.0)   # bench/lstm.go:13-18 sizes, short
