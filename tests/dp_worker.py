"""One rank of the data-parallel parity check (launched by tests/test_gpu_dp.py and by hand:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tests/dp_worker.py

Every rank: shard the rows of one full batch (parallel.shard_rows), run the product's data-parallel step through the C ABI
(af_dp_init + af_mlp_step_fresh_dp / af_mlp_step_host with the plan's dp mode), and compare EVERY rank's reduced
{grad | rgrad} block with (a) the oracle's full-batch gradient and (b) the product's own single-GPU full-batch step, at
1e-12 + 1e-10 |ref|; outputs stay sharded and are compared with the oracle's rows.  torch.distributed carries only the
128-byte communicator id and the final verdict.  Exit code 0 = every check passed on this rank."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

RTOL, ATOL = 1e-10, 1e-12


def close(got, want, what):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, what
    err = np.abs(got - want)
    bad = err > ATOL + RTOL * np.abs(want)
    assert not bad.any(), f"{what}: {bad.sum()} of {bad.size} beyond tolerance; worst {err.max():.3e}"


def main():
    import torch
    import torch.distributed as dist

    from autofunc_b200 import _lib, api, parallel
    from oracle import autofunc_ref as ref

    rank, world, local = parallel.env_rank_world()
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = api.Context(local)
    api.set_context(ctx)
    parallel.dp_init_torch(ctx, local)
    nranks, r, ver = ctx.dp_info()
    assert (nranks, r) == (world, rank) and ver > 0

    # raw collective: every rank contributes rank + 1
    v = ctx.upload(np.full(1000, rank + 1.0))
    ctx.allreduce_sum(v)
    assert np.array_equal(v.numpy(), np.full(1000, world * (world + 1) / 2))

    sizes = [1000, 2000, 512, 10]
    for n_total in (world * 640 + 3, 64):  # ragged shards on the contraction path; a graph-replayed small batch
        onet = ref.MLP(sizes, weight_scale=lambda k: 1 / np.sqrt(k))
        x = onet.make_input(n_total).vector.reshape(n_total, sizes[0])
        ores, ograd, orgrad = onet.step_r(ref.Variable(x.reshape(-1)), n_total)
        start, count = parallel.shard_rows(n_total, world, rank)
        xs = np.ascontiguousarray(x[start:start + count]).reshape(-1)
        plan = api.MLPPlan(sizes, count, ctx)
        for i, pv in enumerate(onet.variables):
            plan.set_param(i, pv.vector)
            plan.set_rvector(i, onet.rvec[pv])
        P = 2 * plan.L
        # the product's own single-GPU full-batch step (rank 0's GPU is as good as any: every rank does it)
        full = api.MLPPlan(sizes, n_total, ctx)
        for i, pv in enumerate(onet.variables):
            full.set_param(i, pv.vector)
            full.set_rvector(i, onet.rvec[pv])
        fout, frout, fgrads, frgrads = full.step_host(x.reshape(-1))
        full.close()
        for mode in (_lib.AF_DP_BLOCKING, _lib.AF_DP_OVERLAP, _lib.AF_DP_DEFERRED):
            plan.set_dp(mode)
            for rep in range(3):
                # host-buffer form: upload the shard, fresh step, reduce, download everything
                out, rout, grads, rgrads = plan.step_host(xs)
                tag = f"n={n_total} mode={mode} rep={rep}"
                o = ores.output().reshape(n_total, -1)[start:start + count].reshape(-1)
                ro = ores.routput().reshape(n_total, -1)[start:start + count].reshape(-1)
                close(out, o, f"sharded Output {tag}")
                close(rout, ro, f"sharded ROutput {tag}")
                for i, pv in enumerate(onet.variables):
                    close(grads[i], ograd[pv], f"grad[{i}] vs oracle {tag}")
                    close(rgrads[i], orgrad[pv], f"rgrad[{i}] vs oracle {tag}")
                    close(grads[i], fgrads[i], f"grad[{i}] vs single-GPU full batch {tag}")
                    close(rgrads[i], frgrads[i], f"rgrad[{i}] vs single-GPU full batch {tag}")
            # resident form: inputs stay in HBM, unit upstream (bench/mlp.go:118-126), explicit join
            ctx.lib.af_upload(ctx.h, plan.input().ptr, xs.ctypes.data, xs.size)
            for rep in range(3):
                plan.step_fresh_dp(True)
            plan.dp_join()
            ctx.sync()
            for i, pv in enumerate(onet.variables):
                close(plan.grad(i).numpy(), ograd[pv], f"resident grad[{i}] n={n_total} mode={mode}")
                close(plan.rgrad(i).numpy(), orgrad[pv], f"resident rgrad[{i}] n={n_total} mode={mode}")
            # af_mlp_step_dp: the caller zeroes the accumulators and supplies the upstream (here a non-unit one)
            up = ref.splitmix_uniform(900, n_total * sizes[-1]).reshape(n_total, -1)[start:start + count].reshape(-1)
            upr = ref.splitmix_uniform(901, n_total * sizes[-1]).reshape(n_total, -1)[start:start + count].reshape(-1)
            plan.zero_grads()
            ctx.lib.af_upload(ctx.h, plan.upstream().ptr, np.ascontiguousarray(up).ctypes.data, up.size)
            ctx.lib.af_upload(ctx.h, plan.upstream_r().ptr, np.ascontiguousarray(upr).ctypes.data, upr.size)
            plan.step_dp(True)
            plan.dp_join()
            ctx.sync()
            _, og2, org2 = onet.step_r(ref.Variable(x.reshape(-1)), n_total,
                                       ref.splitmix_uniform(900, n_total * sizes[-1]), ref.splitmix_uniform(901, n_total * sizes[-1]))
            for i, pv in enumerate(onet.variables):
                close(plan.grad(i).numpy(), og2[pv], f"step_dp grad[{i}] n={n_total} mode={mode}")
                close(plan.rgrad(i).numpy(), org2[pv], f"step_dp rgrad[{i}] n={n_total} mode={mode}")
            plan.step_fresh_dp(True)  # back to the unit-upstream gradients for the checks below
            plan.dp_join()
            ctx.sync()
            # every rank holds bit-identical sums (replicas must not drift apart under AddToVars)
            slab = plan.grad_slab().numpy()
            t = torch.from_numpy(slab.copy()).cuda()
            lo, hi = t.clone(), t.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            assert torch.equal(lo, hi), f"ranks disagree on the reduced block (n={n_total} mode={mode})"
        # sharded host download: the ranks' host buffers together hold the reduced gradient exactly once
        plan.set_dp(_lib.AF_DP_OVERLAP)
        plan.set_host_shard(True)
        grads = [np.zeros(plan.param_len(i)) for i in range(P)]
        rgrads = [np.zeros(plan.param_len(i)) for i in range(P)]
        plan.step_host(xs, grads=grads, rgrads=rgrads)
        flat = torch.from_numpy(np.concatenate(grads + rgrads)).cuda()
        dist.all_reduce(flat)  # disjoint shards: the sum is the union
        want = np.concatenate([ograd[pv] for pv in onet.variables] + [orgrad[pv] for pv in onet.variables])
        close(flat.cpu().numpy(), want, f"union of the ranks' host shards n={n_total}")
        touched = sum(int(np.count_nonzero(g)) for g in grads + rgrads)
        assert touched <= 2 * sum(plan.param_len(i) for i in range(P)) // world + 64 * P, "a rank downloaded more than its shard"
        plan.set_host_shard(False)
        plan.close()

    # LSTM: lanes shard the same way (bench/lstm.go one sequence per lane)
    from autofunc_b200.lstm import LSTMPlan
    I, H, O, T, B = 6, 16, 4, 5, 2 * world
    net = ref.LSTMNet(I, H, O, weight_scale=0.3)
    x = ref.splitmix_uniform(71, T * B * I, 0, 1).reshape(T, B, I)
    up = ref.splitmix_uniform(72, T * B * O, 0, 1).reshape(T, B, O)
    grad = {pv: np.zeros_like(pv.vector) for pv in net.variables}
    rgrad = {pv: np.zeros_like(pv.vector) for pv in net.variables}
    for b in range(B):
        _, _, g, rg = net.run_r([x[t, b] for t in range(T)], [up[t, b] for t in range(T)], None)
        for pv in net.variables:
            grad[pv] += g[pv]
            rgrad[pv] += rg[pv]
    lanes = slice(2 * rank, 2 * rank + 2)
    lp = LSTMPlan(I, H, O, T, 2, ctx)
    for i, pv in enumerate(net.variables):
        lp.set_param(i, pv.vector)
        lp.set_rvector(i, net.rvec[pv])
    lp.set_inputs(ctx.upload(np.ascontiguousarray(x[:, lanes]).reshape(-1)))
    for rep in range(3):
        ups = np.ascontiguousarray(up[:, lanes]).reshape(-1)
        ctx.lib.af_upload(ctx.h, lp.upstream().ptr, ups.ctypes.data, ups.size)
        api.check(ctx.lib.af_zero(ctx.h, lp.upstream_r().ptr, ups.size))
        lp.step_dp(True)
    ctx.sync()
    for i, pv in enumerate(net.variables):
        close(lp.grad(i).numpy(), grad[pv], f"lstm grad[{i}]")
        close(lp.rgrad(i).numpy(), rgrad[pv], f"lstm rgrad[{i}]")
    lp.close()

    calls = ctx.dp_calls()
    assert calls > 0
    dist.barrier()
    if rank == 0:
        print(f"dp_worker ok: world={world} nccl={ver} collectives issued per rank={calls}", flush=True)
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
