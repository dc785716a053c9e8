#!/bin/bash
# 1-GPU evidence refresh: smoke, whole GPU suite, the driver's bench command (both arms)
TAG=${1:-r3e}
O=gpurun_out; mkdir -p $O
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) > $O/${TAG}_smoke.log; cat $O/${TAG}_smoke.log
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 ) > $O/${TAG}_pytest_gpu.log
tail -6 $O/${TAG}_pytest_gpu.log
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err
tail -2 $O/${TAG}_bench_n1.err
python - $O/${TAG}_bench_n1.json <<'PY'
import sys, json
d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print('ms/step', round(d['ms_per_step'], 3), 'e2e', d['e2e'], 'roofline frac', round(d['roofline']['frac'], 3), 'whole', round(d['roofline']['whole_solve']['frac'], 3))
print('parity', d['parity'])
print('mac', d.get('mac_projection'))
print('eb', d.get('eb_projection'))
print('cpu', d.get('cpu_baseline'))
PY
