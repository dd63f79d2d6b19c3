"""Build and drive tests/cpp/host_tests.cpp (the C++ host side of the drop-in, include/autofunc_b200.hpp).

Two link targets: the product library (GPU tests) and oracle/abi_ref.cpp, the CPU restatement of the C ABI (CPU
tests of the host logic; test infrastructure only).  Vector files use the reference's Variable wire format
(variable.go:24-37): little-endian u64 count, then little-endian f64s."""
from __future__ import annotations

import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INCLUDE = os.path.join(ROOT, "include")
CPP = os.path.join(ROOT, "tests", "cpp")
CXXFLAGS = ["-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror"]


def build_ref_abi() -> str:
    from oracle.build import build_abi_ref
    return build_abi_ref()


def build_host_tests(out_dir: str, against: str) -> str:
    """Compile host_tests.cpp and link it against `against` (a path to the product or the reference .so)."""
    exe = os.path.join(out_dir, "host_tests")
    lib_dir = os.path.dirname(against)
    cmd = ["g++", *CXXFLAGS, "-I" + INCLUDE, "-I" + CPP, os.path.join(CPP, "host_tests.cpp"), against,
           "-Wl,-rpath," + lib_dir, "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def write_vectors(path: str, vectors) -> None:
    with open(path, "wb") as f:
        for v in vectors:
            a = np.ascontiguousarray(v, dtype="<f8").reshape(-1)
            f.write(np.uint64(a.size).astype("<u8").tobytes())
            f.write(a.tobytes())


def read_vectors(path: str) -> list[np.ndarray]:
    b = open(path, "rb").read()
    out, pos = [], 0
    while pos < len(b):
        n = int(np.frombuffer(b[pos:pos + 8], dtype="<u8")[0])
        out.append(np.frombuffer(b[pos + 8:pos + 8 + 8 * n], dtype="<f8").copy())
        pos += 8 + 8 * n
    return out


def run_case(exe: str, tmp: str, case: str, inputs, args, timeout=600) -> list[np.ndarray]:
    fin, fout = os.path.join(tmp, f"{case}.in"), os.path.join(tmp, f"{case}.out")
    write_vectors(fin, inputs)
    r = subprocess.run([exe, "run", case, fin, fout, *[str(a) for a in args]], capture_output=True, text=True,
                       timeout=timeout)
    assert r.returncode == 0, f"{case} {args}: rc={r.returncode}\n{r.stdout}\n{r.stderr}"
    assert "run ok" in r.stdout
    return read_vectors(fout)


# ---- the same cases on the oracle --------------------------------------------------------------------------
def close(got, want, what, rtol=1e-10, atol=1e-12):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, f"{what}: shape {got.shape} vs {want.shape}"
    err = np.abs(got - want)
    bad = err > atol + rtol * np.abs(want)
    assert not bad.any(), f"{what}: {bad.sum()} of {bad.size} beyond 1e-12+1e-10|ref|; worst {err.max():.3e}"


def mlp_case(ref, sizes, n, scaled=True, seed=0):
    """Inputs for `run mlp` and the oracle network they describe."""
    net = ref.MLP(sizes, seed=123 + seed, rseed=125 + seed,
                  weight_scale=(lambda k: 1 / np.sqrt(k)) if scaled else None)
    x = net.make_input(n, seed=124 + seed)
    m = sizes[-1] * n
    up, upr = ref.splitmix_uniform(77 + seed, m), ref.splitmix_uniform(78 + seed, m)
    inputs = [v.vector for v in net.variables] + [net.rvec[v] for v in net.variables] + [x.vector, up, upr]
    return net, x, up, upr, inputs


def check_mlp(ref, exe, tmp, mode, sizes, n, scaled=True, seed=0):
    net, x, up, upr, inputs = mlp_case(ref, sizes, n, scaled, seed)
    got = run_case(exe, tmp, "mlp", inputs, [mode, n, *sizes])
    P = len(net.variables)
    if mode == "nonr":
        res, grad = net.step(x, n, up)
        assert len(got) == 1 + P
        close(got[0], res.output(), "output")
        for i, v in enumerate(net.variables):
            close(got[1 + i], grad[v], f"grad[{i}]")
        return
    res, grad, rgrad = net.step_r(x, n, up, upr)
    assert len(got) == 2 + 2 * P
    close(got[0], res.output(), f"{mode} output")
    close(got[1], res.routput(), f"{mode} r-output")
    for i, v in enumerate(net.variables):
        want = np.zeros_like(grad[v]) if mode == "nilgrad" else grad[v]  # grad == nil: the map stays untouched
        close(got[2 + i], want, f"{mode} grad[{i}]")
        close(got[2 + P + i], rgrad[v], f"{mode} rgrad[{i}]")


def check_lintran(ref, exe, tmp, rows, cols, n, seed=0):
    """bench/lintran.go:60-90 with a filled W: the input is a variable too."""
    w = ref.Variable(ref.splitmix_uniform(301 + seed, rows * cols))
    x = ref.Variable(ref.splitmix_uniform(302 + seed, cols * n))
    rv = {w: ref.splitmix_uniform(303 + seed, rows * cols), x: ref.splitmix_uniform(304 + seed, cols * n)}
    up, upr = ref.splitmix_uniform(305 + seed, rows * n), ref.splitmix_uniform(306 + seed, rows * n)
    got = run_case(exe, tmp, "lintran", [w.vector, rv[w], x.vector, rv[x], up, upr], [rows, cols, n])
    res = ref.LinTran(w, rows, cols).batch_r(rv, ref.RVariable(x, rv), n)
    g, rg = ref.new_gradient([w, x]), ref.new_rgradient([w, x])
    out, rout = res.output().copy(), res.routput().copy()
    res.propagate_rgradient(up.copy(), upr.copy(), rg, g)
    for a, b, what in zip(got, [out, rout, g[w], rg[w], g[x], rg[x]], ["Y", "RY", "gW", "rgW", "gX", "rgX"]):
        close(a, b, f"lintran {rows}x{cols} n={n} {what}")


def check_arith(ref, exe, tmp, n, m, k, seed=0):
    x, b = ref.Variable(ref.splitmix_uniform(401 + seed, n * m)), ref.Variable(ref.splitmix_uniform(402 + seed, m))
    rv = {x: ref.splitmix_uniform(403 + seed, n * m), b: ref.splitmix_uniform(404 + seed, m)}
    up, upr = ref.splitmix_uniform(405 + seed, 1 + k), ref.splitmix_uniform(406 + seed, 1 + k)
    got = run_case(exe, tmp, "arith", [x.vector, b.vector, rv[x], rv[b], up, upr], [n, m, k])
    rx, rb = ref.RVariable(x, rv), ref.RVariable(b, rv)
    h = ref.add_r(rx, ref.repeat_r(rb, n))
    gte = ref.mul_r(ref.Sigmoid().apply_r(rv, h), ref.Exp().apply_r(rv, ref.scale_r(rx, 0.5)))
    o = ref.concat_r(ref.sum_all_r(ref.square_r(ref.sub_r(gte, rx))), ref.slice_r(gte, 1, 1 + k))
    g, rg = ref.new_gradient([x, b]), ref.new_rgradient([x, b])
    out, rout = o.output().copy(), o.routput().copy()
    o.propagate_rgradient(up.copy(), upr.copy(), rg, g)
    for a, w_, what in zip(got, [out, rout, g[x], g[b], rg[x], rg[b]], ["out", "rout", "gX", "gB", "rgX", "rgB"]):
        close(a, w_, f"arith {what}")


def check_next(ref, exe, tmp, rows, cols, n, seed=0):
    """SURVEY 8(f) rank 3 through the C++ host layer: host_tests.cpp RunNext, node for node on the oracle."""
    M = ref.Variable(ref.splitmix_uniform(601 + seed, rows * cols))
    V = ref.Variable(ref.splitmix_uniform(602 + seed, cols * n))
    B = ref.Variable(ref.splitmix_uniform(603 + seed, rows))
    rv = {M: ref.splitmix_uniform(604 + seed, rows * cols), V: ref.splitmix_uniform(605 + seed, cols * n),
          B: ref.splitmix_uniform(606 + seed, rows)}
    m = 2 * rows + 1
    up, upr = ref.splitmix_uniform(607 + seed, m), ref.splitmix_uniform(608 + seed, m)
    got = run_case(exe, tmp, "next", [M.vector, V.vector, B.vector, rv[M], rv[V], rv[B], up, upr], [rows, cols, n])
    rM, rV, rB = ref.RVariable(M, rv), ref.RVariable(V, rv), ref.RVariable(B, rv)
    y = ref.mat_mul_vecs_r(ref.add_scaler_r(ref.square_r(rM), 1), rows, cols, rV)
    p = ref.pool_r(y, lambda q: ref.div_r(q, ref.add_scaler_r(ref.square_r(q), 1)))
    sm = ref.Softmax().batch_r(rv, p, n)
    folded = ref.fold_r(rB, ref.split_r(n, sm), lambda state, x: ref.mul_r(
        ref.inverse_r(ref.add_scaler_r(ref.square_r(state), 1)), ref.add_r(x, state)))
    cost, nrm, lse = ref.squared_error_r(folded, rB), ref.Norm().apply_r(rv, p), ref.sum_all_log_domain_r(folded)
    pw = ref.pow_r(ref.add_scaler_r(ref.square_r(folded), 1), 0.3)
    o = ref.concat_r(ref.scale_first_r(ref.add_first_r(pw, nrm), lse), cost,
                     ref.Cos().apply_r(rv, ref.add_log_domain_r(folded, rB)))
    g, rg = ref.new_gradient([M, V, B]), ref.new_rgradient([M, V, B])
    out, rout = o.output().copy(), o.routput().copy()
    o.propagate_rgradient(up.copy(), upr.copy(), rg, g)
    names = ["out", "rout", "gM", "gV", "gB", "rgM", "rgV", "rgB"]
    for a, w_, what in zip(got, [out, rout, g[M], g[V], g[B], rg[M], rg[V], rg[B]], names):
        close(a, w_, f"next {rows}x{cols} n={n} {what}")


def check_lstm(ref, exe, tmp, I, H, O, T, B, scale=0.5, case="lstm"):
    """case "lstm": the fused plan (b200::LSTM); case "lstmops": bench/lstm.go's network assembled from the C++ host
    layer's operators.  Both against the oracle's LSTMNet, lane by lane."""
    from tests.test_gpu_lstm import _oracle
    net = ref.LSTMNet(I, H, O, weight_scale=scale)
    x = ref.splitmix_uniform(501, T * B * I, 0, 1).reshape(T, B, I)
    up = ref.splitmix_uniform(502, T * B * O, 0, 1).reshape(T, B, O)
    upr = ref.splitmix_uniform(503, T * B * O).reshape(T, B, O)
    outs, routs, grad, rgrad = _oracle(net, x, up, upr, True)
    inputs = [v.vector for v in net.variables] + [net.rvec[v] for v in net.variables] + [x, up, upr]
    got = run_case(exe, tmp, case, inputs, [I, H, O, T, B])
    close(got[0], outs.reshape(-1), "lstm outputs")
    close(got[1], routs.reshape(-1), "lstm r-outputs")
    for i, v in enumerate(net.variables):
        close(got[2 + i], grad[v], f"lstm grad[{i}]")
        close(got[12 + i], rgrad[v], f"lstm rgrad[{i}]")
