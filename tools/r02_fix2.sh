#!/bin/bash
# store-type launches through the ordered fix-up in the default mode; sk_fix_bias A/B (103 = model's own choice, 90 / 80 push
# the multi-wave forward / input-gradient launches to stream-K too)
t=${1:-r02fixb}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_determinism.py tests/test_gpu_parity.py tests/test_gpu_full_size.py -x -q -m gpu -k "not lstm" > gpurun_out/${t}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${t}_pytest.log
for b in 103 90 80 103 90; do
  timeout 300 python bench.py --steps 20 --warmup 5 --trace --no-cpu-baseline --opt sk_fix_bias=$b > gpurun_out/${t}_bench_b$b.json 2> gpurun_out/${t}_bench_b$b.err; echo "bench sk_fix_bias=$b rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/${t}_bench_b$b.json").readline())
print("sk_fix_bias=$b", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["gpu_launches"])
PY
done
for b in 103 90 80; do echo "== $b"; grep -E "^trace" gpurun_out/${t}_bench_b$b.err | tail -9 | cut -c1-110; done
