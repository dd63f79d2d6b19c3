#!/bin/bash
# short-row softmax kernels: parity tests, then the streaming bench (copy + softmax cases)
t=${1:-r02soft}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_next.py -x -q -m gpu -k "softmax" > gpurun_out/${t}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${t}_pytest.log
AF_STREAM_ONLY=copy timeout 300 python tools/bench_streaming.py > gpurun_out/${t}_streaming.jsonl 2> gpurun_out/${t}_streaming.err
AF_STREAM_ONLY=softmax timeout 300 python tools/bench_streaming.py >> gpurun_out/${t}_streaming.jsonl 2>> gpurun_out/${t}_streaming.err
cut -c1-200 gpurun_out/${t}_streaming.jsonl; tail -3 gpurun_out/${t}_streaming.err
