#!/bin/bash
# Round-2 8-GPU confirmation: DP parity on 4 ranks, the 1-GPU line of the same box, the default bench line at N ranks
# (deferred join; reports all three schedules, the collective alone, e2e with sharded downloads, strong scaling).
N=${1:-8}; tag=${2:-r02dp8}
mkdir -p gpurun_out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@"; }
timeout 600 python -m pytest tests/test_gpu_dp.py -x -q -m gpu --durations=5 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest dp rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; echo "bench n1 rc=$?"
run bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${tag}_bench_n${N}.json 2> gpurun_out/${tag}_bench_n${N}.err; echo "bench n$N rc=$?"
run tools/bench_configs.py c3 c4 > gpurun_out/${tag}_configs_n${N}.jsonl 2> gpurun_out/${tag}_configs_n${N}.err; echo "configs n$N rc=$?"; cat gpurun_out/${tag}_configs_n${N}.jsonl | cut -c1-250
python - <<PY
import json
def load(p):
    try:
        for l in reversed(open(p).read().splitlines()):
            if l.startswith('{"metric"'): return json.loads(l)
    except Exception as e: return None
b1=load("gpurun_out/${tag}_bench_n1.json")
print("n1", b1 and (b1["value"], b1["ms_per_step"], b1["roofline"]["frac"], b1["e2e"]["value"]))
b=load("gpurun_out/${tag}_bench_n${N}.json")
if b:
    dp=b["data_parallel"]
    print("value %.0f ms %.3f eff %.3f e2e %.0f e2e_eff %.3f | alone %.3f local %.3f exposed %.3f calls/step %.1f modes %s clocks %s"%(
        b["value"], b["ms_per_step"], b["value"]/($N*b1["value"]) if b1 else 0, b["e2e"]["value"], b["e2e"]["value"]/($N*b1["e2e"]["value"]) if b1 else 0,
        dp["allreduce_alone_ms"], dp["step_without_reduce_ms"], dp["exposed_ms_per_step"], dp["collectives_per_step"],
        {k: round(v) for k,v in dp["samples_per_s_by_mode"].items()}, b["clocks"]))
    print("  strong:", json.dumps(b.get("strong_scaling")))
PY
