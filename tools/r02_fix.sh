#!/bin/bash
# stream-K with the ordered fix-up: parity + bit-equality tests, then the bench step in deterministic mode (sk_fixup 0 / 1)
# and in the default mode (sk_fixup 1 / 2)
t=${1:-r02fix}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_determinism.py -x -q -m gpu --durations=5 > gpurun_out/${t}_pytest.log 2>&1; echo "pytest rc=$?"; tail -9 gpurun_out/${t}_pytest.log
for cfg in "1 0" "1 1" "0 1" "0 2" "1 1" "0 2"; do
  set -- $cfg
  timeout 300 python bench.py --steps 20 --warmup 5 --trace --no-cpu-baseline --deterministic $1 --opt sk_fixup=$2 > gpurun_out/${t}_bench_d$1_x$2.json 2> gpurun_out/${t}_bench_d$1_x$2.err; echo "bench det=$1 sk_fixup=$2 rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/${t}_bench_d$1_x$2.json").readline())
print("deterministic=$1 sk_fixup=$2", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["gpu_launches"])
PY
done
for f in d1_x0 d1_x1 d0_x1 d0_x2; do echo "== $f"; grep -E "^trace" gpurun_out/${t}_bench_$f.err | tail -16 | grep -v "reduce  \|elementwise  " | cut -c1-110; done
