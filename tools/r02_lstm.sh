#!/bin/bash
# Round-2 LSTM pass (one B200): parity of the fused step kernels at every size, then C4 timing fused vs per-kernel steps.
tag=${1:-r02lstm}; variants=${2:-"AF_OPT_LSTM_FUSED=1;AF_OPT_LSTM_FUSED=1 AF_OPT_PDL=1;AF_OPT_LSTM_FUSED=0"}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lstm.py tests/test_gpu_full_size.py -x -q -m gpu -k "lstm" --durations=6 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${tag}_pytest.log
IFS=';' read -ra VS <<< "$variants"
for v in "${VS[@]}"; do
  echo "== $v"; env $v AF_TRACE=1 timeout 200 python tools/bench_configs.py c4 c4a 2>&1 | grep -v "kind=[345]" | tail -12
done
