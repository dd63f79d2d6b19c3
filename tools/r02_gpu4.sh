#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "plan or bench or step" 2>&1 | tail -3
timeout 200 python bench.py --steps 20 --warmup 5 --trace --no-cpu-baseline > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r02d_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['frac_of_cublas_on_the_same_shapes'], d['e2e']['value'], d['gpu_launches'])"
cat gpurun_out/r02d_bench.err | tail -12
for v in "AF_OPT_LSTM_FUSED=1" "AF_OPT_LSTM_FUSED=0" "AF_OPT_LSTM_FUSED=1 AF_OPT_PDL=1"; do echo "== $v"; env $v timeout 200 python tools/bench_configs.py c4 2>&1 | tail -1; done
