"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, and refuses to run (loudly) without a B200 -- no compute calls here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from autofunc_b200 import _lib, build as af_build


@pytest.fixture(scope="session")
def lib():
    af_build.build()
    return _lib.load()


def _declared_symbols():
    text = open(_lib.HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(af_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_what_binding_binds():
    assert set(_declared_symbols()) == set(_lib.PROTOTYPES)


def test_library_exports_every_declared_symbol(lib):
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (af_[a-z0-9_]+)\b", out))
    missing = [s for s in _declared_symbols() if s not in exported]
    assert not missing, f"symbols declared in include/autofunc_b200.h but not exported: {missing}"


def test_library_is_sm100a_tma_dmma():
    """The shipped code object is sm_100a and carries the TMA + FP64 tensor-core instructions."""
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    assert "sm_100a" in sass
    assert "DMMA.8x8x4" in sass and "UTMALDG" in sass


def test_no_gpu_means_loud_failure(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = lib.af_ctx_create(0, C.byref(h))
    assert rc != 0 and lib.af_last_error()
    from autofunc_b200 import api
    with pytest.raises(_lib.AutofuncError):
        api.Context(0)


def test_product_never_imports_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "autofunc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert "oracle/" not in src, f


def test_variable_wire_format_matches_reference_layout():
    from autofunc_b200 import api
    v = api.Variable(np.arange(5, dtype=np.float64) * 0.5)
    b = v.serialize()  # variable.go:61-68: LE u64 count then LE f64s
    assert b[:8] == (5).to_bytes(8, "little") and len(b) == 48
    assert np.array_equal(api.Variable.deserialize(b).vector, v.vector)


def test_fuse_dense_pattern():
    from autofunc_b200 import api
    w, b = api.Variable(np.zeros(12)), api.Variable(np.zeros(3))
    net = [api.LinTran(w, 3, 4), api.LinAdd(b), api.Sigmoid(), api.Exp()]
    fused = api.fuse_dense(net)
    assert len(fused) == 2 and isinstance(fused[0], api.DenseSigmoid) and fused[0].activation
    fused2 = api.fuse_dense(net[:2])
    assert len(fused2) == 1 and not fused2[0].activation


def test_synth_generator_matches_oracle_generator():
    from autofunc_b200 import synth
    from oracle import autofunc_ref as ref
    for seed in (0, 123, 124, 125, 3125):
        assert np.array_equal(synth.uniform(seed, 4097), ref.splitmix_uniform(seed, 4097))
    params, rvecs = synth.mlp_params([7, 5, 3])
    net = ref.MLP([7, 5, 3])
    for p, v in zip(params, net.variables):
        assert np.array_equal(p, v.vector)
    for r, v in zip(rvecs, net.variables):
        assert np.array_equal(r, net.rvec[v])


def _split_args(text, start):
    """Arguments of the call whose opening parenthesis is at text[start]; returns (args, index after the ')')."""
    depth, args, cur, i = 0, [], [], start
    while True:
        ch = text[i]
        if ch in "([{":
            depth += 1
            if depth > 1:
                cur.append(ch)
        elif ch in ")]}":
            depth -= 1
            if depth == 0:
                break
            cur.append(ch)
        elif ch == "," and depth == 1:
            args.append("".join(cur).strip())
            cur = []
        else:
            cur.append(ch)
        i += 1
    last = "".join(cur).strip()
    if last or args:
        args.append(last)
    return args, i + 1


def _header_arity():
    text = re.sub(r"/\*.*?\*/", "", open(_lib.HEADER_PATH).read(), flags=re.S)
    arity = {}
    for m in re.finditer(r"\b(af_[a-z0-9_]+)\s*\(", text):
        args, _ = _split_args(text, m.end() - 1)
        arity[m.group(1)] = 0 if args in ([], ["void"]) else len(args)
    return arity


def _go_sources():
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "go", "b200")
    return {f: open(os.path.join(root, f)).read() for f in sorted(os.listdir(root)) if f.endswith(".go")}


def test_go_shim_binds_only_declared_symbols():
    """go/b200/*.go cannot be compiled here (no Go toolchain): at least every C.af_* it calls must be an entry point
    the header declares, and every header struct it names must exist."""
    declared = set(_declared_symbols()) | {"af_ctx", "af_mlp", "af_lstm", "af_graph"}
    used = set()
    for src in _go_sources().values():
        used |= set(re.findall(r"\bC\.(af_[a-z0-9_]+)\b", src))
    assert used and not (used - declared), f"go shim calls undeclared entry points: {sorted(used - declared)}"


def test_go_shim_calls_match_the_prototypes():
    """A lint in place of the Go compiler: every C.af_*(...) call passes as many arguments as the header's prototype
    takes, brackets balance in every file, and no forward or backward method falls back to the reference's CPU
    operators (the shim computes nothing on the host)."""
    arity = _header_arity()
    assert arity["af_lintran_fwd_r"] == 10 and arity["af_last_error"] == 0 and arity["af_dense_fwd"] == 13
    calls = 0
    for name, src in _go_sources().items():
        code = re.sub(r"//[^\n]*", "", src)
        code = re.sub(r"/\*.*?\*/", "", code, flags=re.S)
        for o, c in ("()", "[]", "{}"):
            assert code.count(o) == code.count(c), f"{name}: unbalanced {o}{c}"
        for m in re.finditer(r"\bC\.(af_[a-z0-9_]+)\(", code):
            args, _ = _split_args(code, m.end() - 1)
            assert len(args) == arity[m.group(1)], f"{name}: {m.group(1)} called with {len(args)} arguments, prototype has {arity[m.group(1)]}"
            calls += 1
        for cpu_op in ("autofunc.LinTran{", "autofunc.LinAdd{", "autofunc.Sigmoid{", "autofunc.FuncBatcher{", "autofunc.ComposedFunc{"):
            assert cpu_op not in code, f"{name}: constructs the reference's CPU operator {cpu_op}...}}"
    assert calls > 60


def _compile_c_example(tmp_path):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(str(tmp_path), "mlp_step")
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + os.path.join(root, "include"),
           os.path.join(root, "examples", "mlp_step.c"), "-L" + os.path.dirname(_lib.LIB_PATH), "-lautofunc_b200",
           "-Wl,-rpath," + os.path.dirname(_lib.LIB_PATH), "-lm", "-o", exe]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe


def test_header_is_plain_c_and_a_c_client_links(lib, tmp_path):
    """The boundary is a C ABI: the header must compile as pedantic C99 and a C program must link against the library
    with nothing but -lautofunc_b200 (no CUDA, torch or C++ runtime on the link line)."""
    exe = _compile_c_example(tmp_path)
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode != 0 and "af_ctx_create" in r.stderr  # fails loudly without a GPU: no CPU fallback


def test_variable_deserialize_rejects_truncated_buffers():
    """DeserializeVariable (variable.go:24-37) returns an error on a short read."""
    from autofunc_b200 import api
    b = api.Variable(np.arange(4, dtype=np.float64)).serialize()
    for cut in (0, 7, 8, len(b) - 1):
        with pytest.raises(ValueError, match="truncated"):
            api.Variable.deserialize(b[:cut])
    assert np.array_equal(api.Variable.deserialize(b + b"tail").vector, np.arange(4.0))  # trailing bytes belong to the caller


def test_dp_entry_points_without_a_gpu(lib):
    """The data-parallel ABI is part of the boundary: it loads without NCCL at link time (dlopen at call time), rejects
    bad arguments, and the 128-byte communicator id comes from the collective library, not from torch."""
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    assert "nccl" not in out and "torch" not in out
    assert lib.af_allreduce_sum(None, None, 4) == _lib.AF_EINVAL
    assert lib.af_dp_init(None, 2, 0, None) == _lib.AF_EINVAL
    assert lib.af_dp_calls(None) == 0
    buf = C.create_string_buffer(_lib.AF_DP_ID_BYTES)
    rc = lib.af_dp_unique_id(buf)
    assert rc in (_lib.AF_OK, _lib.AF_ENCCL)  # AF_ENCCL: no libnccl.so.2 on this host -- a loud, typed failure
    if rc == _lib.AF_OK:
        assert any(buf.raw)
    else:
        assert b"nccl" in lib.af_last_error().lower()


def test_roofline_traffic_is_tied_to_the_kernel_sources():
    """bench.py reports roofline.traffic only while profiles/roofline_traffic.json matches the contraction sources."""
    import json
    sha = af_build.kernel_source_sha()
    assert re.fullmatch(r"[0-9a-f]{16}", sha)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    tj = json.load(open(os.path.join(root, "profiles", "roofline_traffic.json")))
    assert "dram_bytes_per_launch" in tj and "kernel_source_sha" in tj


def _header_param_kinds():
    """name -> list of parameter kinds ('ptr', 'i64', 'int', 'f64') parsed from the prototypes of the header."""
    text = re.sub(r"/\*.*?\*/", "", open(_lib.HEADER_PATH).read(), flags=re.S)
    kinds = {}
    for m in re.finditer(r"\b(af_[a-z0-9_]+)\s*\(", text):
        args, _ = _split_args(text, m.end() - 1)
        ks = []
        for a in args:
            if a in ("", "void"):
                continue
            if "*" in a or "af_grad_ready_fn" in a:
                ks.append("ptr")
            elif "int64_t" in a:
                ks.append("i64")
            elif "double" in a:
                ks.append("f64")
            elif re.search(r"\bint\b", a):
                ks.append("int")
            else:
                ks.append("?")
        kinds[m.group(1)] = ks
    return kinds


def test_go_shim_argument_kinds_match_the_prototypes():
    """One step beyond arity: in every C.af_*(...) call of the uncompiled cgo shim a scalar parameter
    must be passed a value of the matching C type -- C.int64_t(...) for int64_t, C.int(...) for int, C.double(...) for double
    (or a literal, or a variable the file declares with that conversion) -- and a pointer parameter must not be passed a
    scalar conversion.  cgo performs no implicit numeric conversions, so each of these would be a compile error."""
    kinds = _header_param_kinds()
    assert kinds["af_lintran_fwd"] == ["ptr", "ptr", "ptr", "ptr", "i64", "i64", "i64"]
    assert kinds["af_mlp_create"] == ["ptr", "ptr", "int", "i64", "ptr"] and kinds["af_scale"] == ["ptr", "ptr", "f64", "i64"]
    conv = {"i64": "C.int64_t(", "int": "C.int(", "f64": "C.double("}
    checked = 0
    for name, src in _go_sources().items():
        code = re.sub(r"//[^\n]*", "", src)
        code = re.sub(r"/\*.*?\*/", "", code, flags=re.S)
        declared = {k: set(re.findall(r"\b([A-Za-z_]\w*)\s*:?=\s*" + re.escape(c), code)) for k, c in conv.items()}
        for m in re.finditer(r"\bC\.(af_[a-z0-9_]+)\(", code):
            args, _ = _split_args(code, m.end() - 1)
            for i, (arg, kind) in enumerate(zip(args, kinds[m.group(1)])):
                arg = arg.strip()
                where = f"{name}: {m.group(1)} argument {i + 1} ({arg!r}) for a {kind} parameter"
                if kind == "ptr":
                    assert not arg.startswith(tuple(conv.values())) and not re.fullmatch(r"-?[1-9]\d*(\.\d+)?", arg), where
                elif kind in conv:
                    ok = (arg.startswith(conv[kind]) or re.fullmatch(r"-?\d+" if kind != "f64" else r"-?\d+(\.\d*)?", arg) is not None
                          or arg in declared[kind])
                    assert ok, where
                checked += 1
    assert checked > 400


def test_every_option_the_library_accepts_is_documented_in_the_header():
    """af_set_option takes its knobs by name; a name the header does not explain is a knob nobody can find."""
    import re
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(ROOT, "autofunc_b200", "csrc", "capi.cu")).read()
    names = re.findall(r'!strcmp\(name, "([a-z_0-9]+)"\)', src)
    assert len(names) >= 20
    hdr = open(os.path.join(ROOT, "include", "autofunc_b200.h")).read()
    missing = [n for n in names if f'"{n}"' not in hdr]
    assert not missing, f"options without a description in include/autofunc_b200.h: {missing}"
