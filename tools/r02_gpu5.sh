#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_lstm.py tests/test_gpu_edge_cases.py tests/test_gpu_api_ops.py -x -q -m gpu 2>&1 | tail -3
timeout 200 python bench.py --steps 20 --warmup 5 --trace --no-cpu-baseline > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r02e_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['frac_of_cublas_on_the_same_shapes'], d['e2e']['value'], d['gpu_launches'])"
cat gpurun_out/r02e_bench.err | tail -12
