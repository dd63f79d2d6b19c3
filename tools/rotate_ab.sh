#!/bin/bash
# A/B of the stream-K rotated k-walk (af_set_option "sk_rotate"): parity at full size with it on, then the bench both ways.
mkdir -p gpurun_out
AF_OPTIONS=sk_rotate=1 timeout 150 python -m pytest tests/test_gpu_parity.py -x -q -k "full_size or mlp_plan_matches" > gpurun_out/rot_parity.log 2>&1; echo "parity rc=$?"; tail -2 gpurun_out/rot_parity.log
timeout 60 python bench.py --no-cpu-baseline --trace --sk-rotate 0 > gpurun_out/rot0.json 2> gpurun_out/rot0.trace; echo "rot0 rc=$?"
timeout 60 python bench.py --no-cpu-baseline --trace --sk-rotate 1 > gpurun_out/rot1.json 2> gpurun_out/rot1.trace; echo "rot1 rc=$?"
AF_OPTIONS=sk_rotate=1 timeout 60 python tools/bench_configs.py c3 c4 > gpurun_out/rot1_configs.jsonl 2> gpurun_out/rot1_configs.err; echo "configs rc=$?"
python - <<'PY'
import json
for f in ("rot0", "rot1"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"])
    except Exception as e:
        print(f, "failed", e)
PY
cat gpurun_out/rot1.trace | grep gemm_dual
