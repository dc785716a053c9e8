#!/bin/bash
# 4- (or 8-) GPU visit: slab parity incl. interior ranks, weak + strong bench.  Usage: bash tools/gpu_dist4.sh NGPU tag
N=${1:-4}; TAG=${2:-r1g4}; O=gpurun_out; mkdir -p $O
( timeout 700 python -m pytest tests/test_gpu_dist.py -x -q -k "nccl or default or fused" 2>&1 | tail -8 ) > $O/${TAG}_pytest_dist.log
pr() { python - "$1" "$2" <<'PY'
import sys, json
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print(sys.argv[2], 'n_gpus', d['n_gpus'], 'ms/step', round(d['ms_per_step'], 2), 'Mcell/s', round(d['value'], 1), 'vcycles', d['config']['vcycles'], d['scaling'])
except Exception as e:
    print(sys.argv[2], 'FAILED', e)
PY
}
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 \
   bench.py --gpus $N --steps 5 --warmup 3 --no-e2e > $O/${TAG}_bench_weak.json 2> $O/${TAG}_bench_weak.err
pr $O/${TAG}_bench_weak.json weak256
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 \
   bench.py --gpus $N --steps 3 --warmup 3 --no-e2e --scaling strong --size 512 > $O/${TAG}_strong512.json 2> $O/${TAG}_strong512.err
pr $O/${TAG}_strong512.json strong512
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29516 \
   bench.py --gpus $N --steps 2 --warmup 3 --no-e2e --scaling strong --size 1024 > $O/${TAG}_strong1024.json 2> $O/${TAG}_strong1024.err
pr $O/${TAG}_strong1024.json strong1024
cat $O/${TAG}_pytest_dist.log
exit 0
