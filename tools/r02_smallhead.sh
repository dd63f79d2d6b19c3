#!/bin/bash
# persistent small-batch step with the head run locally: parity, then C2 n = 1 / 4 with and without it
t=${1:-r02sh}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_small_mlp.py tests/test_gpu_parity.py -x -q -m gpu -k "small_mlp or mlp_plan or missing_rvectors or pruning" > gpurun_out/${t}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${t}_pytest.log
for h in 1 0 1 0; do
  echo "== small_mlp_head=$h"; AF_OPT_SMALL_MLP_HEAD=$h timeout 200 python tools/bench_configs.py c2 2>&1 | grep -E "n=1\"|n=4\"|n=2\"" | cut -c1-150
done
