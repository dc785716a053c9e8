#!/bin/bash
# 1-GPU visit: whole GPU suite, prefetch-depth A/B, kernel bench, full ncu capture of the LEVEL-0 smoother launch
TAG=${1:-r2q}
O=gpurun_out; mkdir -p $O
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > $O/${TAG}_pytest_gpu.log
tail -4 $O/${TAG}_pytest_gpu.log
b() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-parity > $O/${TAG}_bench_$name.json 2> $O/${TAG}_bench_$name.err
  python - $O/${TAG}_bench_$name.json $name <<'PY'
import sys, json
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print(sys.argv[2], 'ms/step', round(d['ms_per_step'], 3), 'vcycles', d['config']['vcycles'], 'us/sweep', round(d['roofline']['us_per_launch'], 1), 'frac', round(d['roofline']['frac'], 3), 'whole', round(d['roofline']['whole_solve']['frac'], 3))
except Exception as e:
    print(sys.argv[2], 'FAILED', e)
PY
}
b pf1_a B200NP_PREFETCH=1
b pf2_a B200NP_PREFETCH=2
b pf1_b B200NP_PREFETCH=1
b pf2_b B200NP_PREFETCH=2
B200NP_PREFETCH=2 timeout 300 python bench.py --n 512 --steps 3 --warmup 3 --no-cpu --no-e2e --no-parity > $O/${TAG}_bench_512_pf2.json 2>/dev/null
B200NP_PREFETCH=1 timeout 300 python bench.py --n 512 --steps 3 --warmup 3 --no-cpu --no-e2e --no-parity > $O/${TAG}_bench_512_pf1.json 2>/dev/null
python - $O/${TAG}_bench_512_pf1.json $O/${TAG}_bench_512_pf2.json <<'PY'
import sys, json
for f in sys.argv[1:]:
    try:
        d = json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f.split('/')[-1], 'ms/step', round(d['ms_per_step'], 2), 'us/sweep', round(d['roofline']['us_per_launch'], 1), 'frac', round(d['roofline']['frac'], 3), 'whole', round(d['roofline']['whole_solve']['frac'], 3))
    except Exception as e:
        print(f, 'FAILED', e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^k_smooth_iso$' -s 3 -c 1 -o $O/${TAG}_prof_smooth_256 -f \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-parity > $O/${TAG}_ncu_smooth.log 2>&1
tail -2 $O/${TAG}_ncu_smooth.log
exit 0
