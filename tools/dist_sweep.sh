#!/bin/bash
# 2-GPU sweep of the slab-path knobs: agglomeration threshold x graph capture.  Usage: tools/dist_sweep.sh NGPU out.log
N=${1:-2}; OUT=${2:-gpurun_out/dist_sweep.log}
: > $OUT
for mp in 8 32 64; do for g in 1 0; do
  B200NP_DIST_MIN_PLANES=$mp B200NP_DIST_GRAPH=$g timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
    --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus $N --steps 5 --warmup 3 --no-e2e 2>/dev/null | grep '^{' | \
    python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('mp=$mp graph=$g ms/step', round(d['ms_per_step'],2), 'Mcell/s', round(d['value'],1), 'vcycles', d['config']['vcycles'])" >> $OUT 2>&1
done; done
cat $OUT
