#!/bin/bash
# Multi-GPU visit of round 2 (gpurun --gpus N): transport diagnosis, parity on N ranks, weak line, phase profile, strong record.
# Usage: bash tools/gpu_r2_multi.sh NGPU tag [full] [tests]
N=${1:-4}; TAG=${2:-r2m}; FULL=${3:-}; TESTS=${4:-}
O=gpurun_out; mkdir -p $O
tr() { port=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port "$@"; }
pr() { python - "$1" "$2" <<'PY'
import sys, json
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    if 'error' in d: print(sys.argv[2], 'ERROR', d)
    else:
        print(sys.argv[2], 'n_gpus', d['n_gpus'], 'ms/step', round(d['ms_per_step'], 2), 'Mcell/s', round(d['value'], 1), 'vcycles', d['config']['vcycles'], d['scaling'],
              'transport', d['config'].get('halo_transport'), 'e2e', d['e2e'].get('ms_per_step'))
        print('  parity', d.get('parity'))
        if d.get('strong'): print('  strong', json.dumps(d['strong']))
except Exception as e:
    print(sys.argv[2], 'FAILED', e)
PY
}
nvidia-smi topo -m > $O/${TAG}_topo.txt 2>&1
timeout 120 tools/bin/ipc_probe $N > $O/${TAG}_ipc_probe.txt 2>&1
sort $O/${TAG}_ipc_probe.txt | uniq -c | sort -rn | head -30
STRONG=--no-strong; [ "$FULL" = "full" ] && STRONG=
tr 29541 bench.py --gpus $N --steps 5 --warmup 3 $STRONG > $O/${TAG}_weak.json 2> $O/${TAG}_weak.err
grep -i "b200np" $O/${TAG}_weak.err | head -20
pr $O/${TAG}_weak.json weak256
B200NP_PROFILE=1 tr 29542 bench.py --gpus $N --steps 1 --warmup 1 --no-e2e --no-parity --no-strong > /dev/null 2> $O/${TAG}_prof.err
grep -A75 "profile (rank 0)" $O/${TAG}_prof.err | tail -76 > $O/${TAG}_phase_profile_rank0.txt
if [ -n "$TESTS" ]; then
( B200NP_DIST_MIN_PLANES=8 timeout 600 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -5 ) > $O/${TAG}_pytest_dist.log
cat $O/${TAG}_pytest_dist.log
fi
exit 0
