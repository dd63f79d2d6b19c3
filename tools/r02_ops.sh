#!/bin/bash
# compile-time R-operand presence in the big contractions (gemm_dmma_kernel OPS): parity, bench A/B
t=${1:-r02ops}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_determinism.py tests/test_gpu_edge_cases.py -x -q -m gpu > gpurun_out/${t}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${t}_pytest.log
for f in 0 1 0 1; do
  timeout 300 python bench.py --steps 20 --warmup 5 --trace --no-cpu-baseline --opt ops_runtime=$f > gpurun_out/${t}_bench_f$f.json 2> gpurun_out/${t}_bench_f$f.err; echo "bench ops_runtime=$f rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/${t}_bench_f$f.json").readline())
print("ops_runtime=$f", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["gpu_launches"])
PY
done
for f in 0 1; do grep -E "^trace.*gemm_dual" gpurun_out/${t}_bench_f$f.err | tail -6 | cut -c1-120; done
