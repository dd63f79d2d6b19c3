#!/bin/bash
# Round-2 evidence pass (one B200): the full GPU suite and smoke() on the final code, the bench line and the CPU arm, every
# BASELINE config, the streaming kernels, then ncu over the same commands (a number printed under ncu is never a bench value).
t=${1:-r02f}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${t}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${t}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${t}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${t}_smoke.log
timeout 300 python bench.py --steps 20 --warmup 5 --trace > gpurun_out/${t}_bench.json 2> gpurun_out/${t}_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/${t}_bench_ref.json 2> gpurun_out/${t}_bench_ref.err; echo "ref rc=$?"
timeout 600 python tools/bench_configs.py c1 c2 c3 c4 c4a c5 > gpurun_out/${t}_configs.jsonl 2> gpurun_out/${t}_configs.err; echo "configs rc=$?"; cat gpurun_out/${t}_configs.jsonl | cut -c1-260
timeout 300 python tools/bench_streaming.py > gpurun_out/${t}_streaming.jsonl 2> gpurun_out/${t}_streaming.err; echo "streaming rc=$?"
timeout 100 python bench.py --steps 3 --warmup 3 --profile-step > gpurun_out/${t}_profile_step.json 2>&1 || exit 1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${t}_launches_step.csv \
    python bench.py --steps 3 --warmup 3 --profile-step > gpurun_out/${t}_ncu_step.log 2>&1; echo "ncu list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_dmma -s 18 -c 6 -f -o gpurun_out/${t}_gemm_full \
    python bench.py --steps 2 --warmup 3 --profile-step > gpurun_out/${t}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | grep ${t}
