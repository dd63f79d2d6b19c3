#!/bin/bash
# every BASELINE config data-parallel at N ranks (weak scaling: the same per-GPU problem on every rank, the sum over ranks scheduled by the library)
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "few_cols or few_rows or head_fusion" 2>&1 | tail -2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) tools/bench_configs.py c1 c2 c3 c4 c5 > gpurun_out/r02_configs_n${N}.jsonl 2> gpurun_out/r02_configs_n${N}.err; echo "configs n$N rc=$?"
cut -c1-230 gpurun_out/r02_configs_n${N}.jsonl; tail -3 gpurun_out/r02_configs_n${N}.err | grep -v OMP_NUM
