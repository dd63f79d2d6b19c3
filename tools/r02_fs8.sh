#!/bin/bash
t=${1:-r02fsu}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/${t}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${t}_pytest.log
for v in "" "AF_OPT_FEW_SAMPLES=0" "AF_OPT_FEW_SAMPLES_MIN=5"; do echo "== C1 C2 $v"; env $v timeout 200 python tools/bench_configs.py c1 c2 2>&1 | cut -c1-200; done
