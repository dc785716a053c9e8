#!/bin/bash
# End-of-round evidence on one B200: smoke, full GPU test suite, bench line, kernel timings, ncu launch list + full capture.
TAG=${1:-r1f}; O=gpurun_out; mkdir -p $O
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 ) > $O/${TAG}_pytest_gpu.log
timeout 600 python bench.py > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err
timeout 300 python bench.py --n 512 --steps 3 --warmup 3 --no-cpu > $O/${TAG}_bench_512.json 2> $O/${TAG}_bench_512.err
timeout 200 python tools/kernel_bench.py 256 rt  > $O/${TAG}_kb_256_rt.log 2>&1
timeout 200 python tools/kernel_bench.py 512 rt  > $O/${TAG}_kb_512_rt.log 2>&1
B200NP_PROFILE=1 timeout 200 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > /dev/null 2> $O/${TAG}_phase_profile_256.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > $O/${TAG}_ncu_bench.log 2>&1
for op in smooth interp; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_smooth|k_interp" -c 2 \
      -f -o $O/${TAG}_prof_${op}_256 python tools/prof_op.py 256 rt $op 0 2 > $O/${TAG}_ncu_${op}.log 2>&1
done
tail -3 $O/${TAG}_smoke.log; tail -3 $O/${TAG}_pytest_gpu.log; head -c 400 $O/${TAG}_bench_n1.json
exit 0
