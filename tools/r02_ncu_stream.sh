#!/bin/bash
# DRAM bytes of the short-row softmax kernels (beside the staged ones) under ncu: traffic must equal the algorithmic bytes
t=${1:-r02ncus}
mkdir -p gpurun_out
export AF_STREAM_N=$((1<<25)) AF_STREAM_STEPS=2 AF_STREAM_WARM=2 AF_STREAM_ONLY="rows of 10"
timeout 120 python tools/bench_streaming.py > /dev/null 2>&1 || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread --clock-control none -k regex:softmax --csv --log-file gpurun_out/${t}_softmax.csv python tools/bench_streaming.py > gpurun_out/${t}_softmax.log 2>&1; echo "ncu rc=$?"
