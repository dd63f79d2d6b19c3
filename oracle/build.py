"""Builds the C++ parts of the checker into oracle/_build/ (git-ignored): oracle/abi_ref.cpp -> libautofunc_b200_ref.so
(CPU restatement of the C ABI for the C++ host tests) and oracle/native_ref.cpp -> libnative_ref.so (the "oracle-native"
CPU baseline: the reference's Gemm-per-call-site structure with a blocked OpenMP Gemm).
TEST INFRASTRUCTURE: called by tests/, bench.py's CPU legs and __graft_entry__.build() (building the checker is not using it)."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
OUT_DIR = os.path.join(HERE, "_build")
REF_LIB = os.path.join(OUT_DIR, "libautofunc_b200_ref.so")


def build_abi_ref() -> str:
    src, hdr = os.path.join(HERE, "abi_ref.cpp"), os.path.join(INCLUDE, "autofunc_b200.h")
    if os.path.exists(REF_LIB) and os.path.getmtime(REF_LIB) >= max(os.path.getmtime(src), os.path.getmtime(hdr)):
        return REF_LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", "-fPIC", "-shared", "-I" + INCLUDE, src,
                    "-o", REF_LIB], check=True, capture_output=True, text=True)
    return REF_LIB


NATIVE_LIB = os.path.join(OUT_DIR, "libnative_ref.so")


def build_native_ref() -> str:
    src = os.path.join(HERE, "native_ref.cpp")
    if os.path.exists(NATIVE_LIB) and os.path.getmtime(NATIVE_LIB) >= os.path.getmtime(src):
        return NATIVE_LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    # no -march: the library is built in one container and timed on another host (and gonum's own kernels are SSE2)
    subprocess.run(["g++", "-std=c++17", "-O3", "-fopenmp", "-Wall", "-Wextra", "-Werror", "-fPIC", "-shared", src, "-o", NATIVE_LIB],
                   check=True, capture_output=True, text=True)
    return NATIVE_LIB


if __name__ == "__main__":
    print(build_abi_ref())
    print(build_native_ref())
