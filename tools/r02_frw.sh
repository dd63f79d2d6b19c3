#!/bin/bash
# few-rows weight gradient (few_rows_wgrad_kernel): parity + bit-reproducibility tests, then the bench step A/B
t=${1:-r02frw}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_few_samples.py tests/test_gpu_determinism.py -x -q -m gpu -k "few_rows or reproducible" > gpurun_out/${t}_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/${t}_pytest.log
for cfg in "1 0" "1 1" "0 0" "0 2"; do
  set -- $cfg
  timeout 300 python bench.py --steps 20 --warmup 5 --trace --no-cpu-baseline --deterministic $1 --opt few_rows_wgrad=$2 > gpurun_out/${t}_bench_d$1_f$2.json 2> gpurun_out/${t}_bench_d$1_f$2.err; echo "bench det=$1 frw=$2 rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/${t}_bench_d$1_f$2.json").readline())
print("deterministic=$1 few_rows_wgrad=$2", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["gpu_launches"])
PY
  grep -vE "gemm_dmma_kernel<[01], [01], 2, 2" gpurun_out/${t}_bench_d$1_f$2.err | tail -30 | cut -c1-160
done
