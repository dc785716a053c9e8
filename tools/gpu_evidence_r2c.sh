#!/bin/bash
# round-2 closing evidence on one B200: smoke, whole GPU suite, the driver's bench command, phase profile, ncu launch list of the bench
TAG=${1:-r2c}; O=gpurun_out; mkdir -p $O
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) > $O/${TAG}_smoke.log; cat $O/${TAG}_smoke.log
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 ) > $O/${TAG}_pytest_gpu.log; tail -3 $O/${TAG}_pytest_gpu.log
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err; tail -2 $O/${TAG}_bench_n1.err
B200NP_PROFILE=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-mac --no-eb --no-512 --no-parity > /dev/null 2> $O/${TAG}_phase_profile_256.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-mac --no-eb --no-512 --no-parity > $O/${TAG}_ncu_bench.log 2>&1
python tools/launch_summary.py $O/${TAG}_launches.csv > $O/${TAG}_launch_list.txt 2>&1; head -8 $O/${TAG}_launch_list.txt; gzip -f $O/${TAG}_launches.csv
python - $O/${TAG}_bench_n1.json <<'PY'
import sys, json
d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print('ms/step', round(d['ms_per_step'], 3), 'e2e ms', d['e2e']['ms_per_step'], 'roofline frac', round(d['roofline']['frac'], 3), 'whole', round(d['roofline']['whole_solve']['frac'], 3))
n = d.get('north_star_512'); print('512:', n['ms_per_solve'], n['smoother_sweep'], n['residual'], n['clocks'])
print('mac', {k: d['mac_projection'][k] for k in ('ms_per_projection', 'vcycles')})
e = d['eb_projection']; print('eb', e['ms_per_projection'], e['sweep']['frac_of_measured_peak'], e.get('parity'))
PY
