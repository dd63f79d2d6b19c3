#!/bin/bash
t=${1:-r02fsz}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_few_samples.py tests/test_gpu_determinism.py -x -q -m gpu > gpurun_out/${t}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${t}_pytest.log
for v in "AF_OPT_FEW_SAMPLES_BLOCKS=3" "AF_OPT_FEW_SAMPLES_BLOCKS=2" "AF_OPT_FEW_SAMPLES_BLOCKS=4" "AF_OPT_FEW_SAMPLES_BLOCKS=3 AF_OPT_FEW_SAMPLES_TMA=0"; do echo "== C1 $v"; env $v timeout 100 python tools/bench_configs.py c1 2>&1 | cut -c1-160; done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:few_samples -s 4 -c 2 --csv --log-file gpurun_out/${t}_ncu.csv env AF_C1_GRAPH=0 python tools/bench_configs.py c1 > gpurun_out/${t}_ncu.log 2>&1
echo "ncu: $(grep -o 'few_samples_[a-z_]*_kernel[^"]*\|"[0-9.]*"$' gpurun_out/${t}_ncu.csv | paste - - | cut -c1-90 | tr '\n' ' ')"
