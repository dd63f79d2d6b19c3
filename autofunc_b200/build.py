"""In-tree build of libautofunc_b200.so (nvcc, sm_100a only).  Cross-compiles without a GPU.
Sources compile to per-file objects in parallel (autofunc_b200/_build/, git-ignored) and only when stale."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libautofunc_b200.so")
SOURCES = ["capi.cu", "gemm.cu", "elementwise.cu", "mathops.cu", "lstm.cu", "small_n.cu", "small_mlp.cu", "dp.cu", "lstm_step.cu", "few_samples.cu"]
HEADERS = ["common.h", "ew.cuh", "gemm_dmma.cuh", "streamk_walk.h", os.path.join("..", "..", "include", "autofunc_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--extended-lambda",
    "-Xcompiler", "-fPIC",
]


def kernel_source_sha() -> str:
    """sha256 (first 16 hex digits) of the contraction kernel's sources: profiles/roofline_traffic.json carries the value
    of the build it was measured on, and bench.py reports `roofline.traffic` only while it still matches."""
    import hashlib
    h = hashlib.sha256()
    for name in ("gemm_dmma.cuh", "gemm.cu", "streamk_walk.h"):
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _mtime(p: str) -> float:
    return os.path.getmtime(p) if os.path.exists(p) else 0.0


def _header_time() -> float:
    return max(_mtime(os.path.join(CSRC, h)) for h in HEADERS)


def _stale(src: str, obj: str) -> bool:
    return _mtime(obj) < max(_mtime(src), _header_time(), _mtime(os.path.abspath(__file__)))


def needs_build() -> bool:
    t = _mtime(LIB)
    return t == 0.0 or any(_mtime(os.path.join(CSRC, s)) > t for s in SOURCES) or _header_time() > t


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    jobs = []
    for s in SOURCES:
        src, obj = os.path.join(CSRC, s), os.path.join(OBJ, s.replace(".cu", ".o"))
        if force or _stale(src, obj):
            cmd = [nvcc, *NVCC_FLAGS, *(["-Xptxas", "-v"] if verbose else []), "-c", src, "-o", obj]
            jobs.append(cmd)
    def run(cmd):
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True, cwd=CSRC)
    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 2) or 1) as pool:
        list(pool.map(run, jobs))
    objs = [os.path.join(OBJ, s.replace(".cu", ".o")) for s in SOURCES]
    # -ldl: dp.cu resolves libnccl with dlopen at af_dp_* time (the library itself links neither NCCL nor torch)
    run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB, *objs, "-ldl"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
