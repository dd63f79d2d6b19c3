#!/bin/bash
# usage: tools/scale_run.sh "<ranks list>" "<overlap list>"  -> gpurun_out/scale_n<N>_ov<K>.log
mkdir -p gpurun_out
for n in $1; do
  for ov in $2; do
    out=gpurun_out/scale_n${n}_ov${ov}.log
    if [ "$n" = "1" ]; then
      timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > $out 2>&1
    else
      timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
        --master-port $((29600 + n * 10 + ov)) bench.py --gpus $n --steps 10 --warmup 3 --overlap $ov > $out 2>&1
    fi
    tail -1 $out | cut -c1-160
  done
done
