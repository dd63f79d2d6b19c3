#!/bin/bash
# Round profile pass (run under gpurun, one GPU): plain runs first, then ncu over the same commands.
set -x
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1 || exit 1
python bench.py --steps 3 --warmup 3 --impl reference > gpurun_out/bench_ref.log 2>&1
python tools/bench_configs.py c1 c2 c3 c4 > gpurun_out/configs.log 2>&1
python tools/bench_streaming.py > gpurun_out/streaming.log 2>&1
# launch list of the bench command (shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
# streaming kernels: DRAM bytes and time per launch
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread
AF_STREAM_N=33554432 AF_STREAM_WARM=2 AF_STREAM_STEPS=2 ncu --metrics $M --clock-control none -k regex:'ew_kernel|softmax|colsum|reduce2_partial|sum_partial' \
    -c 120 --csv --log-file gpurun_out/launches_streaming.csv python tools/bench_streaming.py > gpurun_out/ncu_streaming.log 2>&1
# the persistent small-batch kernel
ncu --metrics $M,l2__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct --clock-control none \
    -k regex:small_mlp_step -s 30 -c 4 --csv --log-file gpurun_out/launches_small_mlp.csv \
    python tools/bench_configs.py c2 > gpurun_out/ncu_small.log 2>&1
ls -la gpurun_out | head -30
