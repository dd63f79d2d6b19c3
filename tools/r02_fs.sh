#!/bin/bash
# few-samples kernels: parity, then C1 and the small C2 batches with and without them (one B200)
t=${1:-r02fs}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_few_samples.py tests/test_gpu_parity.py tests/test_gpu_determinism.py tests/test_gpu_graph.py tests/test_gpu_edge_cases.py -x -q -m gpu > gpurun_out/${t}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/${t}_pytest.log
for v in "" "AF_C1_SPLIT=1" "AF_C1_SPLIT=1 AF_OPT_FEW_SAMPLES=0"; do echo "== C1 $v"; env $v AF_TRACE=1 timeout 100 python tools/bench_configs.py c1 2>&1 | cut -c1-230; done
