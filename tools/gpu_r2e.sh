#!/bin/bash
# 1-GPU visit: parity tests, bench with the direct top level / tail cut-offs, phase profile
TAG=${1:-r2e}
O=gpurun_out; mkdir -p $O
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > $O/${TAG}_pytest_gpu.log
tail -3 $O/${TAG}_pytest_gpu.log
b() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-parity > $O/${TAG}_bench_$name.json 2> $O/${TAG}_bench_$name.err
  python - $O/${TAG}_bench_$name.json $name <<'PY'
import sys, json
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print(sys.argv[2], 'ms/step', round(d['ms_per_step'], 3), 'vcycles', d['config']['vcycles'], 'resid', d['config']['resid_over_bnorm'], 'whole', round(d['roofline']['whole_solve']['frac'], 3))
except Exception as e:
    print(sys.argv[2], 'FAILED', e)
PY
}
b default A=1
b tail0 B200NP_TAIL=0
b tail125 B200NP_TAIL_NODES=125
b tail4913 B200NP_TAIL_NODES=4913
b nodirect B200NP_TOP_DIRECT=0
b nodirect_tail0 B200NP_TOP_DIRECT=0 B200NP_TAIL=0
B200NP_PROFILE=1 timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-parity > /dev/null 2> $O/${TAG}_phase_profile_256.txt
tail -40 $O/${TAG}_phase_profile_256.txt
exit 0
