"""World-size-2 gloo test of the data-parallel path's host logic: shard the batch, run the step per
rank, sum-all-reduce the gradients, compare with the single-process full-batch result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from autofunc_b200 import parallel
from oracle import autofunc_ref as ref


def test_shard_rows_cover_and_balance():
    for n in (1, 7, 4096, 65537):
        for world in (1, 2, 3, 8):
            blocks = [parallel.shard_rows(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and sum(c for _, c in blocks) == n
            for (s0, c0), (s1, _) in zip(blocks, blocks[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1
    with pytest.raises(ValueError):
        parallel.shard_rows(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, sizes, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        net = ref.MLP(sizes, weight_scale=lambda k: 1 / np.sqrt(k))
        x = net.make_input(n).vector.reshape(n, sizes[0])
        start, count = parallel.shard_rows(n, world, rank)
        xs = ref.Variable(x[start:start + count].reshape(-1))
        res, grad, rgrad = net.step_r(xs, count)
        slab = torch.from_numpy(np.concatenate([grad[v] for v in net.variables] + [rgrad[v] for v in net.variables]))
        parallel.allreduce_sum_(slab)  # one collective over the {grad | rgrad} slab, as on the GPU
        q.put((rank, start, count, res.output().copy(), slab.numpy().copy()))
    finally:
        dist.destroy_process_group()


def test_sharded_step_allreduce_equals_full_batch():
    sizes, n, world = [12, 9, 4], 11, 2  # odd batch: ranks get 6 and 5 rows
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, sizes, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    net = ref.MLP(sizes, weight_scale=lambda k: 1 / np.sqrt(k))
    res, grad, rgrad = net.step_r(net.make_input(n), n)
    want = np.concatenate([grad[v] for v in net.variables] + [rgrad[v] for v in net.variables])
    for rank, start, count, out, slab in got:
        np.testing.assert_allclose(slab, want, rtol=1e-12, atol=1e-14)  # every rank holds the full sums
        np.testing.assert_allclose(out, res.output()[start * 4:(start + count) * 4], rtol=1e-13)  # outputs stay sharded
