#!/bin/bash
# A/B: bounded mbarrier wait (default build) vs round 1's unbounded wait, and the fresh-step entry point vs round 1's four calls.
mkdir -p gpurun_out
for v in new old; do
  for st in fresh legacy; do
    lib=""; [ $v = old ] && lib="AUTOFUNC_B200_LIB=$PWD/autofunc_b200/_ab/libautofunc_b200_oldwait.so"
    flag=""; [ $st = legacy ] && flag="--legacy-step"
    env $lib timeout 120 python bench.py --steps 10 --warmup 3 --trace --no-cpu-baseline $flag > gpurun_out/ab_wait_${v}_${st}.json 2> gpurun_out/ab_wait_${v}_${st}.err
    echo "== wait=$v step=$st rc=$?"; python -c "
import json,sys
d=json.loads(open('gpurun_out/ab_wait_${v}_${st}.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])"
    grep "gemm_dual" gpurun_out/ab_wait_${v}_${st}.err | awk '{printf \"%s \", \$4}'; echo
  done
done
