#!/bin/bash
# N-GPU check of the split halo protocol: dist parity tests, weak bench (parity block inside), per-level profile
N=${1:-2}; TAG=${2:-r2n}
O=gpurun_out; mkdir -p $O
tr() { port=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port "$@"; }
( B200NP_DIST_MIN_PLANES=8 timeout 500 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -6 ) > $O/${TAG}_pytest_dist.log
cat $O/${TAG}_pytest_dist.log
tr 29551 bench.py --gpus $N --steps 5 --warmup 3 --no-strong > $O/${TAG}_weak.json 2> $O/${TAG}_weak.err
grep -i "b200np" $O/${TAG}_weak.err | head
python - $O/${TAG}_weak.json <<'PY'
import sys, json
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print('n_gpus', d.get('n_gpus'), 'ms/step', d.get('ms_per_step'), 'vcycles', d.get('config', {}).get('vcycles'), 'transport', d.get('config', {}).get('halo_transport'), 'e2e', d.get('e2e', {}).get('ms_per_step'))
    print('parity', d.get('parity'))
except Exception as e:
    print('FAILED', e)
PY
B200NP_PROFILE=1 tr 29552 bench.py --gpus $N --steps 1 --warmup 1 --no-e2e --no-parity --no-strong > /dev/null 2> $O/${TAG}_prof.err
grep -A75 "profile (rank 0)" $O/${TAG}_prof.err | grep -v "rank [1-9]" | grep "L[012] smooth\|L0 resid\|between" | sort | uniq | head
exit 0
