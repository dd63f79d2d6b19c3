#!/bin/bash
# timing experiments on the few-samples kernels (wrong results by design): which part of a band costs the time
t=${1:-r02fsx}
mkdir -p gpurun_out
for dbg in 0 1 2 3 4 5 7; do
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:few_samples -s 4 -c 2 --csv --log-file gpurun_out/${t}_dbg$dbg.csv env AF_C1_GRAPH=0 AF_OPT_FEW_SAMPLES_DEBUG=$dbg python tools/bench_configs.py c1 > gpurun_out/${t}_dbg$dbg.log 2>&1
echo "dbg=$dbg $(grep -o 'few_samples_[a-z]*_kernel[^"]*\|"[0-9.]*"$' gpurun_out/${t}_dbg$dbg.csv | paste - - | cut -c1-80 | tr '\n' ' ')"
done
