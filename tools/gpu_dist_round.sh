#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): slab parity test, then bench with NCCL halos vs NVLink peer-memory halos.
# Usage: bash tools/gpu_dist_round.sh NGPU tag
N=${1:-2}; TAG=${2:-r1}
O=gpurun_out
mkdir -p $O
nvidia-smi topo -m > $O/${TAG}_topo.txt 2>&1
( timeout 400 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -15 ) > $O/${TAG}_pytest_dist.log
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
     --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-e2e > $O/${TAG}_dist_${name}.json 2> $O/${TAG}_dist_${name}.err
  python - "$O/${TAG}_dist_${name}.json" "$name" <<'PY' >> $O/${TAG}_dist_summary.txt
import sys, json
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print(sys.argv[2], 'ms/step', round(d['ms_per_step'], 2), 'Mcell/s', round(d['value'], 1), 'vcycles', d['config']['vcycles'])
except Exception as e:
    print(sys.argv[2], 'FAILED', e)
PY
}
: > $O/${TAG}_dist_summary.txt
run nccl_mp8      B200NP_P2P=0 B200NP_DIST_MIN_PLANES=8
run p2p_unfused   B200NP_P2P=1 B200NP_FUSE_HALO=0
run p2p_fused_mp8 B200NP_P2P=1 B200NP_DIST_MIN_PLANES=8
run p2p_fused_mp16 B200NP_P2P=1 B200NP_DIST_MIN_PLANES=16
run p2p_fused_mp32 B200NP_P2P=1 B200NP_DIST_MIN_PLANES=32
if [ -n "$FULL" ]; then
  ( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > $O/${TAG}_pytest_gpu.log
  EXTRA="--scaling strong --n 512" 
  env B200NP_P2P=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
     --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --no-e2e --scaling strong --n 512 > $O/${TAG}_dist_strong512.json 2> $O/${TAG}_dist_strong512.err
fi
cat $O/${TAG}_dist_summary.txt
tail -5 $O/${TAG}_dist_p2p_mp8.err
