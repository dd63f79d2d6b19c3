#!/bin/bash
# Session-5 GPU pass (one B200, ~3 min): the full GPU suite and smoke() on the final code, then the evidence runs --
# plain bench first, then ncu over the same commands (a number printed under ncu is never a bench value).
mkdir -p gpurun_out
timeout 150 python -m pytest tests -x -q -m gpu > gpurun_out/s5_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/s5_pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s5_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/s5_smoke.log
timeout 90 python bench.py --steps 10 --warmup 3 > gpurun_out/s5_bench.json 2> gpurun_out/s5_bench.err; echo "bench rc=$?"
timeout 60 python bench.py --steps 3 --warmup 3 --profile-step > gpurun_out/s5_profile_step.json 2>&1 || exit 1
timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s5_launches_step.csv \
    python bench.py --steps 3 --warmup 3 --profile-step > gpurun_out/s5_ncu_step.log 2>&1; echo "ncu list rc=$?"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:gemm_dmma -s 14 -c 7 -f -o gpurun_out/s5_gemm_full \
    python bench.py --steps 2 --warmup 3 --profile-step > gpurun_out/s5_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | tail -8
