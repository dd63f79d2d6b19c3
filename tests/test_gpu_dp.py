"""Data parallelism behind the C ABI, on real GPUs (needs >= 2 devices: `gpurun --gpus 2 -- python -m pytest
tests/test_gpu_dp.py -m gpu`; skipped on a one-GPU box).  The reference's gradient sum is Gradient.Add
(gradient.go:28-30,108-120); here every rank's reduced {grad | rgrad} block must equal the single-GPU full-batch
gradient and the oracle's, at 1e-12 + 1e-10 |ref|, in all three scheduling modes (AF_DP_BLOCKING / OVERLAP / DEFERRED)."""
import os
import socket
import subprocess
import sys
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import autofunc_ref as ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RTOL, ATOL = 1e-10, 1e-12


def _need_two_gpus():
    import torch
    assert torch.cuda.is_available()
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")


def close(got, want, what):
    got, want = np.asarray(got), np.asarray(want)
    err = np.abs(got - want)
    bad = err > ATOL + RTOL * np.abs(want)
    assert not bad.any(), f"{what}: {bad.sum()} of {bad.size} beyond tolerance; worst {err.max():.3e}"


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_dp_one_process_per_gpu_matches_full_batch():
    """The launch bench.py uses: torchrun, one rank per GPU, af_dp_init with the id shipped by torch.distributed."""
    _need_two_gpus()
    import torch
    world = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dp_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, f"dp_worker failed\nstdout:\n{r.stdout[-3000:]}\nstderr:\n{r.stderr[-6000:]}"
    assert "dp_worker ok" in r.stdout


def test_dp_one_process_two_devices_matches_full_batch():
    """The launch a Go host would use: ONE process, one context (and one host thread) per device, af_dp_init_local."""
    _need_two_gpus()
    from autofunc_b200 import _lib, api, parallel
    world = 2
    ctxs = [api.Context(d) for d in range(world)]
    parallel.dp_init_local(ctxs)
    assert [c.dp_info()[:2] for c in ctxs] == [(world, 0), (world, 1)]
    sizes, n_total = [1000, 2000, 512, 10], 1283
    onet = ref.MLP(sizes, weight_scale=lambda k: 1 / np.sqrt(k))
    x = onet.make_input(n_total).vector.reshape(n_total, sizes[0])
    ores, ograd, orgrad = onet.step_r(ref.Variable(x.reshape(-1)), n_total)
    results, errors = [None] * world, []

    def rank_main(rank):
        try:
            ctx = ctxs[rank]
            start, count = parallel.shard_rows(n_total, world, rank)
            plan = api.MLPPlan(sizes, count, ctx)
            for i, pv in enumerate(onet.variables):
                plan.set_param(i, pv.vector)
                plan.set_rvector(i, onet.rvec[pv])
            xs = np.ascontiguousarray(x[start:start + count]).reshape(-1)
            got = {}
            for mode in (_lib.AF_DP_BLOCKING, _lib.AF_DP_OVERLAP, _lib.AF_DP_DEFERRED):
                plan.set_dp(mode)
                for _ in range(2):
                    out, rout, grads, rgrads = plan.step_host(xs)
                got[mode] = (start, count, out, rout, grads, rgrads)
            results[rank] = got
            plan.close()
        except Exception as e:  # surfaced by the main thread
            errors.append((rank, repr(e)))

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=600)
    assert not errors, errors
    for rank in range(world):
        for mode, (start, count, out, rout, grads, rgrads) in results[rank].items():
            close(out, ores.output().reshape(n_total, -1)[start:start + count].reshape(-1), f"Output rank {rank} mode {mode}")
            close(rout, ores.routput().reshape(n_total, -1)[start:start + count].reshape(-1), f"ROutput rank {rank} mode {mode}")
            for i, pv in enumerate(onet.variables):
                close(grads[i], ograd[pv], f"grad[{i}] rank {rank} mode {mode}")
                close(rgrads[i], orgrad[pv], f"rgrad[{i}] rank {rank} mode {mode}")
    for mode in results[0]:
        for i in range(len(onet.variables)):  # replicas hold identical sums
            assert np.array_equal(results[0][mode][4][i], results[1][mode][4][i])
    for c in ctxs:
        c.close()
