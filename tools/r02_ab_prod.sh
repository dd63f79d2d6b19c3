#!/bin/bash
# A/B: rotating producer role (default build) vs the fixed producer in warp 0 (round 1), same box, alternating
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_lstm.py tests/test_gpu_streamk.py tests/test_gpu_edge_cases.py -x -q -m gpu 2>&1 | tail -3
for v in new old new old; do
  lib=""; [ $v = old ] && lib="AUTOFUNC_B200_LIB=$PWD/autofunc_b200/_ab/libautofunc_b200_fixedprod.so"
  env $lib timeout 120 python bench.py --steps 20 --warmup 5 --trace --no-cpu-baseline > gpurun_out/ab_prod_${v}.json 2> gpurun_out/ab_prod_${v}.err
  echo "== producer=$v rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/ab_prod_${v}.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])"
  grep "gemm_dual" gpurun_out/ab_prod_${v}.err | awk '{printf "%s ", $4}'; echo
done
for v in new old; do
  lib=""; [ $v = old ] && lib="AUTOFUNC_B200_LIB=$PWD/autofunc_b200/_ab/libautofunc_b200_fixedprod.so"
  echo "== configs producer=$v"; env $lib timeout 300 python tools/bench_configs.py c3 c4 2>&1 | cut -c1-200
done
