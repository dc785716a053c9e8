#!/bin/bash
# 1-GPU visit: parity tests + the default bench line
TAG=${1:-q}
O=gpurun_out; mkdir -p $O
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $O/${TAG}_pytest_gpu.log
tail -15 $O/${TAG}_pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-parity > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
python - $O/${TAG}_bench.json <<'PY'
import sys, json
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print('ms/step', round(d['ms_per_step'], 3), 'vcycles', d['config']['vcycles'], 'e2e', d['e2e'].get('ms_per_step'), 'whole', round(d['roofline']['whole_solve']['frac'], 3))
except Exception as e:
    print('FAILED', e); print(open(sys.argv[1].replace('.json', '.err')).read()[-2000:])
PY
exit 0
