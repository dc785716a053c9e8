#!/bin/bash
TAG=${1:-r3n}
O=gpurun_out; mkdir -p $O
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err
tail -3 $O/${TAG}_bench_n1.err
python - $O/${TAG}_bench_n1.json <<'PY'
import sys, json
d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print('ms/step', round(d['ms_per_step'], 3), 'e2e ms', d['e2e']['ms_per_step'], 'roofline frac', round(d['roofline']['frac'], 3), 'whole', round(d['roofline']['whole_solve']['frac'], 3))
print('512:', d.get('north_star_512'))
print('mac', {k: d['mac_projection'][k] for k in ('ms_per_projection', 'vcycles')})
e = d['eb_projection']; print('eb', e['ms_per_projection'], e['sweep'], e.get('parity'), e.get('cpu_port', {}).get('seconds'))
PY
