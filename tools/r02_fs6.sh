#!/bin/bash
t=${1:-r02fsw}
mkdir -p gpurun_out
for v in "AF_C1_ONLY=fwd" "AF_C1_ONLY=bwd" ""; do echo "== C1 $v"; env $v timeout 100 python tools/bench_configs.py c1 2>&1 | cut -c1-160; done
timeout 300 ncu --set full --import-source on --clock-control none -k regex:few_samples -s 4 -c 2 -f -o gpurun_out/${t}_c1_full env AF_C1_GRAPH=0 python tools/bench_configs.py c1 > gpurun_out/${t}_ncu_full.log 2>&1; echo "ncu rc=$?"
