#!/bin/bash
# batched bias + sigmoid epilogue of interior forward tiles: parity, bench trace, C3
t=${1:-r02fwdepi}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_gpu_determinism.py -x -q -m gpu -k "not lstm" > gpurun_out/${t}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${t}_pytest.log
for f in 1 2; do
  timeout 300 python bench.py --steps 20 --warmup 5 --trace --no-cpu-baseline > gpurun_out/${t}_bench_$f.json 2> gpurun_out/${t}_bench_$f.err; echo "bench rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/${t}_bench_$f.json").readline())
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["gpu_launches"])
PY
done
grep -E "^trace" gpurun_out/${t}_bench_2.err | tail -9 | cut -c1-120
timeout 300 python tools/bench_configs.py c3 2>&1 | cut -c1-260
