#!/bin/bash
# per-phase profile (B200NP_PROFILE=1) on 1 GPU and on N GPUs + interp kernel check.  Usage: bash tools/gpu_prof_round.sh NGPU tag
N=${1:-2}; TAG=${2:-r1p}
O=gpurun_out; mkdir -p $O
if [ -z "$SKIP1" ]; then
( timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_project.py -x -q 2>&1 | tail -5 ) > $O/${TAG}_pytest_k.log
timeout 200 python tools/kernel_bench.py 256 rt > $O/${TAG}_kb_256_rt.log 2>&1
fi
( timeout 300 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -5 ) > $O/${TAG}_pytest_dist.log
B200NP_PROFILE=1 timeout 200 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > $O/${TAG}_prof1.json 2> $O/${TAG}_prof1.err
if [ "$N" -gt 1 ]; then
B200NP_PROFILE=1 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
   --master-port 29513 bench.py --gpus $N --steps 2 --warmup 1 --no-e2e > $O/${TAG}_profN.json 2> $O/${TAG}_profN.err
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
   --master-port 29514 bench.py --gpus $N --steps 5 --warmup 3 --no-e2e > $O/${TAG}_benchN.json 2> $O/${TAG}_benchN.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
   --master-port 29515 bench.py --gpus $N --steps 3 --warmup 3 --no-e2e --scaling strong --size 512 > $O/${TAG}_strong512.json 2> $O/${TAG}_strong512.err
fi
tail -3 $O/${TAG}_pytest_dist.log; grep "ms/step\|value" $O/${TAG}_benchN.json | head -c 300
exit 0
