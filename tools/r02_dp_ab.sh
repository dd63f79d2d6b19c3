#!/bin/bash
N=${1:-2}
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@" 2>/dev/null | grep '^{"metric"' | python -c "
import json,sys
b=json.loads(sys.stdin.readline()); dp=b['data_parallel']
print('   value %.0f ms %.3f clocks %s | by mode %s'%(b['value'], b['ms_per_step'], b['clocks'], {k: round(v) for k,v in dp['samples_per_s_by_mode'].items()}))"; }
for v in "--steps 10" "--steps 20 --warmup 5" "--no-clocks --steps 20 --warmup 5"; do echo "== $v"; run bench.py --gpus $N --no-strong $v; done
timeout 100 python bench.py --steps 20 --warmup 5 --no-cpu-baseline | python -c "
import json,sys
b=json.loads(sys.stdin.readline()); print('n1', b['value'], b['ms_per_step'], b['clocks'])"
