#!/bin/bash
# quick validation of the slab path on 2 GPUs: parity tests (all exchange variants) + one bench line
O=gpurun_out; TAG=${1:-r1k}; mkdir -p $O
( timeout 400 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -4 ) > $O/${TAG}_pytest_dist.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 \
   bench.py --gpus 2 --steps 5 --warmup 3 > $O/${TAG}_bench2.json 2> $O/${TAG}_bench2.err
cat $O/${TAG}_pytest_dist.log; head -c 300 $O/${TAG}_bench2.json
exit 0
