#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_graph.py -x -q -m gpu 2>&1 | tail -3
for v in "" "AF_OPTIONS=small_wgrad_max=8"; do echo "== C1 $v"; env $v AF_TRACE=1 timeout 100 python tools/bench_configs.py c1 2>&1 | cut -c1-230; done
timeout 200 python bench.py --steps 20 --warmup 5 --trace --no-cpu-baseline 2> gpurun_out/r02g_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])"; grep gemm_dual gpurun_out/r02g_bench.err | awk '{printf "%s ", $4}'; echo
timeout 300 python tools/bench_configs.py c3 2>&1 | cut -c1-200
