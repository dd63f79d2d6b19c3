#!/bin/bash
# few-samples kernels, second pass: parity, C1 as a graph (prefetch on / off, split calls), per-kernel times under ncu
t=${1:-r02fs2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_few_samples.py tests/test_gpu_determinism.py -x -q -m gpu > gpurun_out/${t}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${t}_pytest.log
for v in "" "AF_C1_SPLIT=1" "AF_C1_SPLIT=1 AF_OPT_FEW_SAMPLES=0"; do echo "== C1 $v"; env $v timeout 100 python tools/bench_configs.py c1 2>&1 | cut -c1-200; done
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:few_samples -c 12 --csv --log-file gpurun_out/${t}_ncu_c1.csv env AF_C1_GRAPH=0 python tools/bench_configs.py c1 > gpurun_out/${t}_ncu_c1.log 2>&1; echo "ncu rc=$?"
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/${t}_ncu_c1.csv")))
hdr=[i for i,r in enumerate(rows) if r and r[0]=="ID"]
if hdr:
    h=rows[hdr[0]]; ki=h.index("Kernel Name"); mi=h.index("Metric Name"); vi=h.index("Metric Value"); idi=h.index("ID")
    out={}
    for r in rows[hdr[0]+1:]:
        if len(r)>vi: out.setdefault((r[idi], r[ki][:40]),{})[r[mi]]=r[vi]
    for k,v in list(out.items())[-6:]: print(k, v)
PY
