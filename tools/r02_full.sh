#!/bin/bash
# whole GPU suite + smoke + bench (our arm with cpu baseline) + per-launch trace; t = tag
t=${1:-r02full}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --durations=8 > gpurun_out/${t}_pytest.log 2>&1; echo "pytest rc=$?"; tail -14 gpurun_out/${t}_pytest.log
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --warmup 5 --trace --no-cpu-baseline > gpurun_out/${t}_bench.json 2> gpurun_out/${t}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/${t}_bench.json").readline())
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["gpu_launches"])
PY
grep -E "^trace" gpurun_out/${t}_bench.err | tail -9 | cut -c1-120
