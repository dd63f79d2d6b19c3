#!/usr/bin/env python
"""Extract the golden vectors the reference's own tests hold for the path (and its callers) from the Go test
SOURCES under /root/reference, and write them to tests/golden/reference_fixtures.json.

The reference is pure Go and cannot run in the build image (no Go toolchain), so its outputs cannot be regenerated;
what CAN be pinned mechanically is that every fixture and known-answer vector the parity tests use is, digit for
digit, the one in the reference's test files.  This script is the only thing that reads /root/reference; it runs in
the build container, and the committed JSON travels to the GPU box.  tests/test_golden.py compares tests/fixtures.py
and the oracle's known answers against the JSON.

    python tests/golden/make_golden.py            # rewrite reference_fixtures.json
"""
import json
import os
import re
import sys

REF = os.environ.get("AUTOFUNC_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
NUM = r"[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?(?:\s*/\s*[-+]?(?:\d+\.\d*|\d+))?"


def floats(text):
    out = []
    for tok in re.findall(NUM, text):
        if "/" in tok:
            a, b = tok.split("/")
            out.append(float(a) / float(b))
        else:
            out.append(float(tok))
    return out


def braces_after(src, pos):
    """Contents of the first {...} literal after `pos` that follows a `[]float64` / `linalg.Vector` marker."""
    m = re.compile(r"\[\]float64\s*\{").search(src, pos)
    start = m.end()
    return src[start:src.index("}", start)], m.start()


def named_vector(src, name, occurrence=0):
    """The []float64{...} literal that follows the `occurrence`-th appearance of `name`."""
    pos = -1
    for _ in range(occurrence + 1):
        pos = src.index(name, pos + 1)
    body, _ = braces_after(src, pos)
    return floats(body)


def read(rel):
    with open(os.path.join(REF, rel)) as f:
        return f.read()


def main():
    g = {}
    lt = read("test/lin_tran_test.go")
    g["lin_tran_test.go"] = {
        "linTranTestMat1": named_vector(lt, "var linTranTestMat1"),
        "linTranTestMat2": named_vector(lt, "var linTranTestMat2"),
        "linTranTestVec": named_vector(lt, "var linTranTestVec ="),
        "linTranTestVec2": named_vector(lt, "var linTranTestVec2"),
        "rv_linTranTestMat1": named_vector(lt, "linTranTestMat1.Data: linalg.Vector"),
        "rv_linTranTestMat2": named_vector(lt, "linTranTestMat2.Data: linalg.Vector"),
        "rv_linTranTestVec": named_vector(lt, "linTranTestVec:  linalg.Vector"),
        "rv_linTranTestVec2": named_vector(lt, "linTranTestVec2: linalg.Vector"),
        "TestLinTranOutput_expected": named_vector(lt, "func TestLinTranOutput"),
        "TestLinTranBatchOutput_expected": named_vector(lt, "func TestLinTranBatchOutput"),
    }
    la = read("test/lin_add_test.go")
    g["lin_add_test.go"] = {
        "linAddTestVec1": named_vector(la, "linAddTestVec1 ="),
        "linAddTestVec2": named_vector(la, "linAddTestVec2 ="),
        "rv_linAddTestVec1": named_vector(la, "linAddTestVec1:"),
        "rv_linAddTestVec2": named_vector(la, "linAddTestVec2:"),
    }
    mf = read("test/math_funcs_test.go")
    g["math_funcs_test.go"] = {
        "mathFuncTestVec": named_vector(mf, "mathFuncTestVec "),
        "mathFuncTestVecPos": named_vector(mf, "mathFuncTestVecPos"),
        "rv_mathFuncTestVec": named_vector(mf, "mathFuncTestVec:"),
        "TestNormOutput_input": named_vector(mf, "func TestNormOutput"),
    }
    ar = read("test/arithmetic_test.go")
    g["arithmetic_test.go"] = {
        "arithmeticTestVec1": named_vector(ar, "arithmeticTestVec1 ="),
        "arithmeticTestVec2": named_vector(ar, "arithmeticTestVec2 ="),
        "arithmeticTestVec3": named_vector(ar, "arithmeticTestVec3 ="),
        "arithmeticTestVec4": named_vector(ar, "arithmeticTestVec4 ="),
        "rv_arithmeticTestVec1": named_vector(ar, "arithmeticTestVec1: linalg.Vector"),
        "rv_arithmeticTestVec3": named_vector(ar, "arithmeticTestVec3: linalg.Vector"),
        "rv_arithmeticTestVec4": named_vector(ar, "arithmeticTestVec4: linalg.Vector"),
        "TestSumAllLogDomainOutput_input": named_vector(ar, "func TestSumAllLogDomainOutput"),
    }
    po = read("test/pool_test.go")
    g["pool_test.go"] = {
        "poolTestVar": floats(re.search(r"poolTestVar = &Variable\{\[\]float64\{([^}]*)\}", po).group(1)),
        "rv_poolTestVar": named_vector(po, "poolTestVar:"),
    }
    fo = read("test/fold_test.go")
    g["fold_test.go"] = {
        "input": named_vector(fo, "func TestFoldOut"),
        "init_state": floats(re.findall(r"Vector: \[\]float64\{([^}]*)\}", fo)[1]),
        "TestFoldOut_expected": floats(re.search(r"expected := \[\]float64\{([^}]*)\}", fo).group(1)),
    }
    al = read("test/lin_alg_test.go")
    g["lin_alg_test.go"] = {
        "TestMatMulVecOutput_expected": named_vector(al, "func TestMatMulVecOutput"),
    }
    mp = read("seqfunc/test/map_test.go")
    g["seqfunc/test/map_test.go"] = {
        "TranMatrix": named_vector(mp, "TranMatrix = &autofunc.Variable"),
        "TestVars": [floats(x) for x in re.findall(r"&autofunc\.Variable\{Vector: \[\]float64\{([^}]*)\}\}", mp)],
        "rv_TranMatrix": named_vector(mp, "TranMatrix: []float64"),
    }
    fc = read("functest/func_checker.go")
    g["functest/func_checker.go"] = {
        "DefaultDelta_DefaultPrec": floats(" ".join(re.findall(r"Default(?:Delta|Prec)\s*=\s*([0-9.eE+-]+)", fc))),
    }
    vt = read("variable.go")
    g["variable.go"] = {"wire_format": "little-endian uint64 count, then little-endian float64 values"
                        if "binary.LittleEndian" in vt or "LittleEndian" in vt else "UNVERIFIED"}
    out = os.path.join(HERE, "reference_fixtures.json")
    with open(out, "w") as f:
        json.dump(g, f, indent=1, sort_keys=True)
        f.write("\n")
    print(f"wrote {out}: {sum(len(v) for v in g.values())} entries from {len(g)} reference files")


if __name__ == "__main__":
    sys.exit(main())
