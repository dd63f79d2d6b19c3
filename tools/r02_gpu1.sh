#!/bin/bash
# Round-2 single-GPU pass: full GPU suite (incl. the C4 / C5 full-size parity tests), smoke(), the bench line and the CPU arm.
tag=${1:-r02a}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --durations=12 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -22 gpurun_out/${tag}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${tag}_smoke.log
timeout 200 python bench.py --steps 10 --warmup 3 --trace > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/${tag}_bench.json; tail -20 gpurun_out/${tag}_bench.err
timeout 200 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; echo "ref rc=$?"; tail -c 800 gpurun_out/${tag}_bench_ref.json
