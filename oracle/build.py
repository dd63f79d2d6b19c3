"""Builds the C++ part of the checker: oracle/abi_ref.cpp -> oracle/_build/libautofunc_b200_ref.so (git-ignored).
TEST INFRASTRUCTURE: called by tests/ and by __graft_entry__.build() (building the checker is not using it)."""
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


if __name__ == "__main__":
    print(build_abi_ref())
