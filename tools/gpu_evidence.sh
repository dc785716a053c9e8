#!/bin/bash
# evidence refresh: smoke, whole GPU suite, the driver's bench command
TAG=${1:-r3p}
O=gpurun_out; mkdir -p $O
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) > $O/${TAG}_smoke.log; cat $O/${TAG}_smoke.log
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 ) > $O/${TAG}_pytest_gpu.log
tail -6 $O/${TAG}_pytest_gpu.log
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err
tail -2 $O/${TAG}_bench_n1.err
python - $O/${TAG}_bench_n1.json <<'PY'
import sys, json
d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print('ms/step', round(d['ms_per_step'], 3), 'e2e ms', d['e2e']['ms_per_step'], 'roofline frac', round(d['roofline']['frac'], 3), 'whole', round(d['roofline']['whole_solve']['frac'], 3))
n = d.get('north_star_512'); print('512:', n['ms_per_solve'], n['smoother_sweep'], n['residual'], n['clocks'])
print('mac', {k: d['mac_projection'][k] for k in ('ms_per_projection', 'vcycles')})
e = d['eb_projection']; print('eb', e['ms_per_projection'], e['sweep']['frac_of_measured_peak'], e.get('parity'))
PY
