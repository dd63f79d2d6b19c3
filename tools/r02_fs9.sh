#!/bin/bash
t=${1:-r02fst}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_few_samples.py tests/test_gpu_parity.py tests/test_gpu_determinism.py tests/test_gpu_edge_cases.py tests/test_gpu_graph.py tests/test_gpu_api_ops.py -x -q -m gpu > gpurun_out/${t}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${t}_pytest.log
for v in "" "AF_OPT_FEW_SAMPLES_MIN=1 AF_OPT_SMALL_MLP=0" "AF_OPT_FEW_SAMPLES_MIN=2 AF_OPT_SMALL_MLP=0"; do echo "== C2 $v"; env $v timeout 200 python tools/bench_configs.py c2 2>&1 | cut -c1-130 | head -4; done
