#!/bin/bash
N=${1:-8}
run() { timeout 60 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 scripts/nccl_time.py 2>&1 | grep "all_reduce" ; }
run X=1
run NCCL_ALGO=NVLS
run NCCL_PROTO=LL
run NCCL_ALGO=Tree NCCL_PROTO=LL
run NCCL_PROTO=LL128
