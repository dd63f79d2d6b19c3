#!/bin/bash
# Round-2 multi-GPU pass (gpurun --gpus N): the data-parallel parity tests, then the bench at N ranks in the three reduce
# schedules (the bench line itself reports all three + the collective alone + the step without it), and the 1-GPU line
# of the same box for the efficiency.
N=${1:-2}; tag=${2:-r02dp}; extra=${3:-}
mkdir -p gpurun_out
run() { timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@"; }
timeout 900 python -m pytest tests/test_gpu_dp.py -x -q -m gpu --durations=5 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest dp rc=$?"; tail -8 gpurun_out/${tag}_pytest.log
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; echo "bench n1 rc=$?"
for mode in ${MODES:-overlap deferred blocking}; do
  run bench.py --gpus $N --steps 10 --warmup 3 --dp-mode $mode --no-strong $extra > gpurun_out/${tag}_bench_n${N}_${mode}.json 2> gpurun_out/${tag}_bench_n${N}_${mode}.err; echo "bench n$N $mode rc=$?"
  tail -3 gpurun_out/${tag}_bench_n${N}_${mode}.err
done
python - <<PY
import json
def load(p):
    try: return json.loads(open(p).read().strip().splitlines()[-1])
    except Exception as e: return None
b1=load("gpurun_out/${tag}_bench_n1.json")
print("n1", b1 and (b1["value"], b1["ms_per_step"], b1["roofline"]["frac"], b1["e2e"]["value"]))
for mode in ("overlap","deferred","blocking"):
    b=load("gpurun_out/${tag}_bench_n${N}_%s.json"%mode)
    if not b: print(mode, "no line"); continue
    dp=b["data_parallel"]
    print(mode, "value %.0f ms %.3f eff %.3f e2e %.0f e2e_eff %.3f | alone %.3f local %.3f exposed %.3f calls/step %.1f modes %s"%(
        b["value"], b["ms_per_step"], b["value"]/($N*b1["value"]) if b1 else 0, b["e2e"]["value"], b["e2e"]["value"]/($N*b1["e2e"]["value"]) if b1 else 0,
        dp["allreduce_alone_ms"], dp["step_without_reduce_ms"], dp["exposed_ms_per_step"], dp["collectives_per_step"],
        {k: round(v) for k,v in dp["samples_per_s_by_mode"].items()}))
PY
