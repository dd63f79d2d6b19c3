#!/bin/bash
# batched-row sigmoid-backward epilogue of the input-gradient launches: parity tests, the bench step, C3
t=${1:-r02epi}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_gpu_determinism.py -x -q -m gpu -k "bias_gradient_fused or mlp_plan or bench_configuration or reproducible or lintran_kernels or fused_dense or c5" > gpurun_out/${t}_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/${t}_pytest.log
for f in 1 0 1; do
  timeout 300 python bench.py --steps 20 --warmup 5 --trace --no-cpu-baseline --opt fused_colsum=$f > gpurun_out/${t}_bench_f$f.json 2> gpurun_out/${t}_bench_f$f.err; echo "bench fused_colsum=$f rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/${t}_bench_f$f.json").readline())
print("fused_colsum=$f", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["gpu_launches"])
PY
done
for f in 1 0; do grep -E "^trace" gpurun_out/${t}_bench_f$f.err | tail -12 | cut -c1-120; done
timeout 300 python tools/bench_configs.py c3 2>&1 | cut -c1-260
