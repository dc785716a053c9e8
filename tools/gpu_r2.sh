#!/bin/bash
# One 1-GPU box visit of round 2: parity tests, bench line (+ reference arm), per-kernel timings, phase profile.
# Usage (under gpurun): bash tools/gpu_r2.sh [tag] [quick]
TAG=${1:-r2}; QUICK=${2:-}
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_gpu.txt 2>&1
nproc > $O/${TAG}_nproc.txt
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > $O/${TAG}_pytest_gpu.log
timeout 600 python bench.py > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err
tail -3 $O/${TAG}_pytest_gpu.log
python - $O/${TAG}_bench_n1.json <<'PY'
import sys, json
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print('bench: ms/step', round(d['ms_per_step'], 3), 'Mcell/s', round(d['value'], 1), 'vcycles', d['config']['vcycles'], 'e2e', d['e2e'].get('ms_per_step'),
          'roofline', round(d['roofline']['frac'], 3), 'us/sweep', round(d['roofline']['us_per_launch'], 1), 'whole', round(d['roofline']['whole_solve']['frac'], 3))
    print('parity', d.get('parity')); print('cpu', d.get('cpu_baseline'))
except Exception as e:
    print('bench FAILED', e); print(open(sys.argv[1].replace('.json', '.err')).read()[-3000:])
PY
[ -n "$QUICK" ] && exit 0
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err
cat $O/${TAG}_bench_ref.json | cut -c1-400
timeout 300 python bench.py --n 512 --steps 3 --warmup 3 --no-cpu --no-parity > $O/${TAG}_bench_512.json 2> $O/${TAG}_bench_512.err
timeout 200 python tools/kernel_bench.py 256 rt  > $O/${TAG}_kb_256_rt.log 2>&1
KB_ALL=1 timeout 200 python tools/kernel_bench.py 512 rt  > $O/${TAG}_kb_512_rt.log 2>&1
B200NP_PROFILE=1 timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-parity > /dev/null 2> $O/${TAG}_phase_profile_256.txt
cat $O/${TAG}_kb_256_rt.log
exit 0
