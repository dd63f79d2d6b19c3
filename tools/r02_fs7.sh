#!/bin/bash
t=${1:-r02fsv}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_few_samples.py tests/test_gpu_determinism.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/${t}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${t}_pytest.log
for v in "" "AF_C1_ONLY=fwd" "AF_C1_ONLY=bwd" "AF_OPT_FEW_SAMPLES_TMA=0" "AF_C1_SPLIT=1 AF_OPT_FEW_SAMPLES=0"; do echo "== C1 $v"; env $v timeout 100 python tools/bench_configs.py c1 2>&1 | cut -c1-160; done
