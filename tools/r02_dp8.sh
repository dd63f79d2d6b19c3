#!/bin/bash
# Round-2 8-GPU confirmation: DP parity on 4 ranks, the 1-GPU line of the same box, the default bench line at N ranks
# (deferred join; reports all three schedules, the collective alone, e2e with sharded downloads, strong scaling), and the
# in-step overlap line for comparison.
N=${1:-8}; tag=${2:-r02dp8}
mkdir -p gpurun_out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@"; }
timeout 600 python -m pytest tests/test_gpu_dp.py -x -q -m gpu --durations=5 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest dp rc=$?"; tail -6 gpurun_out/${tag}_pytest.log
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; echo "bench n1 rc=$?"
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,TUNING run bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${tag}_bench_n${N}_deferred.json 2> gpurun_out/${tag}_bench_n${N}_deferred.err; echo "bench n$N deferred rc=$?"
grep -i "nvls\|comm 0x.*nranks\|Algo\|channels" gpurun_out/${tag}_bench_n${N}_deferred.err | sort | uniq -c | sort -rn | head -12
run bench.py --gpus $N --steps 20 --warmup 5 --dp-mode overlap --no-strong > gpurun_out/${tag}_bench_n${N}_overlap.json 2> gpurun_out/${tag}_bench_n${N}_overlap.err; echo "bench n$N overlap rc=$?"
python - <<PY
import json
def load(p):
    try: return json.loads(open(p).read().strip().splitlines()[-1])
    except Exception as e: return None
b1=load("gpurun_out/${tag}_bench_n1.json")
print("n1", b1 and (b1["value"], b1["ms_per_step"], b1["roofline"]["frac"], b1["e2e"]["value"]))
for mode in ("deferred","overlap"):
    b=load("gpurun_out/${tag}_bench_n${N}_%s.json"%mode)
    if not b: print(mode, "no line"); continue
    dp=b["data_parallel"]
    print(mode, "value %.0f ms %.3f eff %.3f e2e %.0f e2e_eff %.3f | alone %.3f local %.3f exposed %.3f calls/step %.1f modes %s"%(
        b["value"], b["ms_per_step"], b["value"]/($N*b1["value"]) if b1 else 0, b["e2e"]["value"], b["e2e"]["value"]/($N*b1["e2e"]["value"]) if b1 else 0,
        dp["allreduce_alone_ms"], dp["step_without_reduce_ms"], dp["exposed_ms_per_step"], dp["collectives_per_step"],
        {k: round(v) for k,v in dp["samples_per_s_by_mode"].items()}))
    print("  strong:", json.dumps(b.get("strong_scaling")))
PY
