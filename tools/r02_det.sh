#!/bin/bash
# deterministic mode: the reproducibility tests, then the bench step with and without it (one B200)
t=${1:-r02det}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_determinism.py -x -q -m gpu --durations=5 > gpurun_out/${t}_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/${t}_pytest.log
for d in 1 0; do
  timeout 300 python bench.py --steps 20 --warmup 5 --trace --no-cpu-baseline --deterministic $d > gpurun_out/${t}_bench_d$d.json 2> gpurun_out/${t}_bench_d$d.err; echo "bench d=$d rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/${t}_bench_d$d.json").readline())
print("deterministic=$d", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["gpu_launches"])
PY
  grep -E "gemm|colsum" gpurun_out/${t}_bench_d$d.err | tail -24 | cut -c1-150
done
