#!/bin/bash
# 1-GPU visit: whole GPU suite, default bench line, kernel bench, full ncu capture of the level-0 smoother
TAG=${1:-r2p}
O=gpurun_out; mkdir -p $O
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > $O/${TAG}_pytest_gpu.log
tail -4 $O/${TAG}_pytest_gpu.log
timeout 600 python bench.py > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err
python - $O/${TAG}_bench_n1.json <<'PY'
import sys, json
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print('bench: ms/step', round(d['ms_per_step'], 3), 'Mcell/s', round(d['value'], 1), 'vcycles', d['config']['vcycles'], 'e2e', d['e2e'].get('ms_per_step'),
          'roofline', round(d['roofline']['frac'], 3), 'us/sweep', round(d['roofline']['us_per_launch'], 1), 'whole', round(d['roofline']['whole_solve']['frac'], 3))
    print('parity', d.get('parity', {}).get('ok'), 'cpu', d.get('cpu_baseline', {}).get('value'))
except Exception as e:
    print('bench FAILED', e); print(open(sys.argv[1].replace('.json', '.err')).read()[-3000:])
PY
timeout 200 python tools/kernel_bench.py 256 rt  > $O/${TAG}_kb_256_rt.log 2>&1; head -8 $O/${TAG}_kb_256_rt.log
B200NP_PROFILE=1 timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-parity > /dev/null 2> $O/${TAG}_phase_profile_256.txt
grep "setup\|divu\|mknewu\|top \|between" $O/${TAG}_phase_profile_256.txt | tail -7
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^k_smooth_iso$' -s 20 -c 1 -o $O/${TAG}_prof_smooth_256 -f \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-parity > $O/${TAG}_ncu_smooth.log 2>&1
tail -3 $O/${TAG}_ncu_smooth.log
exit 0
