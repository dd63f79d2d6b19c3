#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dp.py -x -q -m gpu 2>&1 | tail -2
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@"; }
run bench.py --gpus $N --steps 20 --warmup 5 --no-strong 2>/dev/null | grep '^{"metric"' | python -c "
import json,sys
b=json.loads(sys.stdin.readline()); dp=b['data_parallel']
print('value %.0f ms %.3f | alone %.3f | by mode %s'%(b['value'], b['ms_per_step'], dp['allreduce_alone_ms'], {k: round(v) for k,v in dp['samples_per_s_by_mode'].items()}))"
run tools/bench_configs.py c1 c2 c4 2>/dev/null | cut -c1-200
