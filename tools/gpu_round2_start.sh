#!/bin/bash
# First multi-GPU visit of round 2 (gpurun --gpus 4 or 8): the measurements round 1 ran out of budget for.
#   1. per-level phase profile of the slab path (where do 5.1 ms per V-cycle at 4 GPUs go?)
#   2. zero-start pre-smooth on slabs (default off there: unmeasured) -- parity + A/B
#   3. weak 256^3/GPU and strong 512^3 / 1024^3 bench lines
# Usage: bash tools/gpu_round2_start.sh NGPU tag
N=${1:-4}; TAG=${2:-r2a}; O=gpurun_out; mkdir -p $O
tr() { port=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port "$@"; }
pr() { python - "$1" "$2" <<'PY'
import sys, json
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print(sys.argv[2], 'n_gpus', d['n_gpus'], 'ms/step', round(d['ms_per_step'], 2), 'Mcell/s', round(d['value'], 1), 'vcycles', d['config']['vcycles'], d['scaling'])
except Exception as e:
    print(sys.argv[2], 'FAILED', e)
PY
}
B200NP_PROFILE=1 tr 29531 bench.py --gpus $N --steps 2 --warmup 1 --no-e2e > $O/${TAG}_prof.json 2> $O/${TAG}_prof.err
grep -A70 "profile (rank 0)" $O/${TAG}_prof.err | tail -71 > $O/${TAG}_phase_profile_rank0.txt
( B200NP_ZERO_START=1 B200NP_DIST_MIN_PLANES=8 timeout 300 python -m pytest tests/test_gpu_dist.py -x -q -k "fused and not unfused" 2>&1 | tail -4 ) > $O/${TAG}_pytest_dist_zero.log
for z in 0 1; do
  B200NP_ZERO_START=$z tr 29532 bench.py --gpus $N --steps 5 --warmup 3 --no-e2e > $O/${TAG}_weak_zero$z.json 2> $O/${TAG}_weak_zero$z.err
  pr $O/${TAG}_weak_zero$z.json "weak256 zero_start=$z"
done
tr 29533 bench.py --gpus $N --steps 3 --warmup 3 --no-e2e --scaling strong --size 512 > $O/${TAG}_strong512.json 2> $O/${TAG}_strong512.err
pr $O/${TAG}_strong512.json strong512
tr 29534 bench.py --gpus $N --steps 2 --warmup 3 --no-e2e --scaling strong --size 1024 > $O/${TAG}_strong1024.json 2> $O/${TAG}_strong1024.err
pr $O/${TAG}_strong1024.json strong1024
tr 29535 tools/kernel_bench_dist.py 256 > $O/${TAG}_kbd.log 2>&1; grep -A6 "^rank 0" $O/${TAG}_kbd.log
cat $O/${TAG}_pytest_dist_zero.log
exit 0
