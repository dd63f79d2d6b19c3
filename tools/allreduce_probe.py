#!/usr/bin/env python
"""Time the data-parallel collective alone: NCCL sum all-reduce of the bench step's {grad | rgrad} slab
(6.06 M doubles = 48.5 MB) and of C3 / C5 sized slabs, device-timed, max over ranks.  Run under torchrun;
NCCL_* environment variables select algorithm / protocol for A/B runs."""
import json
import os
import sys

import torch
import torch.distributed as dist


def main():
    rank, world, local = (int(os.environ.get(k, "0")) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sizes = [int(s) for s in (sys.argv[1:] or ["6062044", "33570816"])]
    for n in sizes:
        t = torch.ones(n, dtype=torch.float64, device="cuda")
        for _ in range(5):
            dist.all_reduce(t)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 20
        e0.record()
        for _ in range(iters):
            dist.all_reduce(t)
        e1.record()
        e1.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda", dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rank == 0:
            ms = float(ms.item())
            print(json.dumps({"doubles": n, "MB": round(n * 8 / 1e6, 1), "world": world, "ms": round(ms, 4),
                              "algbw_GBps": round(n * 8 / ms / 1e6, 1),
                              "busbw_GBps": round(n * 8 / ms / 1e6 * 2 * (world - 1) / world, 1),
                              "env": {k: v for k, v in os.environ.items() if k.startswith("NCCL_")}}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
