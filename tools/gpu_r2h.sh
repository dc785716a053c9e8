#!/bin/bash
TAG=${1:-r2h}
O=gpurun_out; mkdir -p $O
( timeout 900 python -m pytest tests/test_gpu_multibox.py tests/test_prob_bc.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -15 ) > $O/${TAG}_pytest_gpu.log
tail -15 $O/${TAG}_pytest_gpu.log
b() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-parity > $O/${TAG}_bench_$name.json 2> $O/${TAG}_bench_$name.err
  python - $O/${TAG}_bench_$name.json $name <<'PY'
import sys, json
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print(sys.argv[2], 'ms/step', round(d['ms_per_step'], 3), 'vcycles', d['config']['vcycles'], 'whole', round(d['roofline']['whole_solve']['frac'], 3))
except Exception as e:
    print(sys.argv[2], 'FAILED', e)
PY
}
b direct1 B200NP_TOP_DIRECT=1
b correction1 B200NP_TOP_DIRECT=0
b direct2 B200NP_TOP_DIRECT=1
b correction2 B200NP_TOP_DIRECT=0
exit 0
