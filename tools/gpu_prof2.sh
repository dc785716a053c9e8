#!/bin/bash
# per-level profile of the slab path for several agglomeration thresholds.  Usage: bash tools/gpu_prof2.sh NGPU tag
N=${1:-2}; TAG=${2:-r1s}
O=gpurun_out; mkdir -p $O
( timeout 300 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -5 ) > $O/${TAG}_pytest_dist.log
for mp in 8 32 64; do
B200NP_DIST_MIN_PLANES=$mp B200NP_PROFILE=1 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
   --master-port 29513 bench.py --gpus $N --steps 2 --warmup 1 --no-e2e > $O/${TAG}_prof_mp$mp.json 2> $O/${TAG}_prof_mp$mp.err
B200NP_DIST_MIN_PLANES=$mp timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
   --master-port 29514 bench.py --gpus $N --steps 5 --warmup 3 --no-e2e > $O/${TAG}_bench_mp$mp.json 2> $O/${TAG}_bench_mp$mp.err
done
tail -3 $O/${TAG}_pytest_dist.log
exit 0
