#!/bin/bash
# last pass on the final library: full GPU suite, smoke(), bench line (with roofline.traffic from the matching ncu capture), CPU arm
t=r02z; mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${t}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${t}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${t}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${t}_smoke.log
timeout 300 python bench.py --steps 20 --warmup 5 --trace > gpurun_out/${t}_bench.json 2> gpurun_out/${t}_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/${t}_bench_ref.json 2> gpurun_out/${t}_bench_ref.err; echo "ref rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/${t}_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['traffic'], d['e2e']['value'], d['cpu_baseline']['value'], d['gpu_launches'])"
