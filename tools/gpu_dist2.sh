#!/bin/bash
N=${1:-2}; TAG=${2:-r1j}; O=gpurun_out; mkdir -p $O
( timeout 500 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -8 ) > $O/${TAG}_pytest_dist.log
run() { name=$1; shift
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
     tools/kernel_bench_dist.py 256 > $O/${TAG}_kbd_$name.log 2>&1; grep -A6 "^rank 0" $O/${TAG}_kbd_$name.log; }
run fused_mp8 B200NP_DIST_MIN_PLANES=8
for mp in 64 8; do
B200NP_DIST_MIN_PLANES=$mp timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
   --master-port 29514 bench.py --gpus $N --steps 5 --warmup 3 --no-e2e > $O/${TAG}_bench_mp$mp.json 2> $O/${TAG}_bench_mp$mp.err
python - $O/${TAG}_bench_mp$mp.json $mp <<'PY'
import sys, json
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print('mp', sys.argv[2], 'ms/step', round(d['ms_per_step'], 2), 'Mcell/s', round(d['value'], 1), 'vcycles', d['config']['vcycles'])
except Exception as e:
    print('FAILED', e)
PY
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
   --master-port 29515 bench.py --gpus $N --steps 3 --warmup 3 --no-e2e --scaling strong --size 512 > $O/${TAG}_strong512.json 2> $O/${TAG}_strong512.err
python - $O/${TAG}_strong512.json <<'PY'
import sys, json
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print('strong 512^3: ms/step', round(d['ms_per_step'], 2), 'Mcell/s', round(d['value'], 1), 'vcycles', d['config']['vcycles'], d['scaling'])
except Exception as e:
    print('FAILED', e)
PY
cat $O/${TAG}_pytest_dist.log
exit 0
