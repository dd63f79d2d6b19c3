run() { echo "== $*"; env "$@" timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) tools/allreduce_probe.py 2>&1 | grep -E '^\{|rror' | cut -c1-220; }
run NCCL_DEBUG=WARN
run NCCL_ALGO=Ring
run NCCL_ALGO=Tree
run NCCL_ALGO=NVLS
run NCCL_ALGO=NVLSTree
run NCCL_ALGO=Ring NCCL_PROTO=Simple NCCL_MIN_NCHANNELS=32
run NCCL_NVLS_ENABLE=0
