#!/bin/bash
# LSTM parity + C4 timing with the fused steps and the streaming x projection; bench A/B bounded vs unbounded mbarrier wait on one box.
mkdir -p gpurun_out
bash tools/r02_lstm.sh r02lstm3 "AF_OPT_LSTM_FUSED=1;AF_OPT_LSTM_FUSED=0"
for v in new old new old; do
  lib=""; [ $v = old ] && lib="AUTOFUNC_B200_LIB=$PWD/autofunc_b200/_ab/libautofunc_b200_oldwait.so"
  env $lib timeout 120 python bench.py --steps 20 --warmup 5 --trace --no-cpu-baseline > gpurun_out/ab2_wait_${v}.json 2> gpurun_out/ab2_wait_${v}.err
  echo "== wait=$v rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/ab2_wait_${v}.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])"
  grep "gemm_dual" gpurun_out/ab2_wait_${v}.err | awk '{printf "%s ", $4}'; echo
done
