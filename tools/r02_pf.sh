#!/bin/bash
# epilogue prefetch A/B on the bench step + ncu of few_rows_wgrad_kernel
t=${1:-r02pf}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "bias_gradient_fused or mlp_plan or lintran_kernels or fused_dense" > gpurun_out/${t}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${t}_pytest.log
for f in 1 0 1 0; do
  timeout 300 python bench.py --steps 20 --warmup 5 --trace --no-cpu-baseline --opt epi_prefetch=$f > gpurun_out/${t}_bench_f$f.json 2> gpurun_out/${t}_bench_f$f.err; echo "bench epi_prefetch=$f rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/${t}_bench_f$f.json").readline())
print("epi_prefetch=$f", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["gpu_launches"])
PY
  grep -E "^trace" gpurun_out/${t}_bench_f$f.err | tail -9 | sed -n 7,8p | cut -c1-120
done
timeout 120 python tools/ncu_frw.py && timeout 300 ncu --set full --clock-control none --import-source on -k regex:few_rows_wgrad -s 2 -c 1 -o gpurun_out/${t}_frw python tools/ncu_frw.py > gpurun_out/${t}_ncu.log 2>&1; echo "ncu rc=$?"
