#!/bin/bash
# head-fusion parity (bench config + plan tests) and bench; ncu --set full of one fused LSTM forward step and one backward step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_small_mlp.py tests/test_gpu_graph.py tests/test_gpu_lstm.py tests/test_gpu_cpp_host.py -x -q -m gpu > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02c_pytest.log
timeout 200 python bench.py --steps 20 --warmup 5 --trace --no-cpu-baseline > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r02c_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['frac_of_cublas_on_the_same_shapes'], d['e2e']['value'], d['gpu_launches'])"
cat gpurun_out/r02c_bench.err | tail -16
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_dmma -s 10 -c 1 -f -o gpurun_out/r02_lstm_fwd_step python tools/ncu_lstm.py > gpurun_out/r02_ncu_lstm_fwd.log 2>&1; echo "ncu fwd rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_dmma -s 140 -c 1 -f -o gpurun_out/r02_lstm_bwd_step python tools/ncu_lstm.py > gpurun_out/r02_ncu_lstm_bwd.log 2>&1; echo "ncu bwd rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -3
