#!/bin/bash
# One GPU-box visit: parity tests, bench line, per-kernel timings, ncu launch list and full captures.
# Usage (under gpurun): bash tools/gpu_round.sh [tag]
TAG=${1:-r1}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_gpu.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > $O/${TAG}_pytest_gpu.log
timeout 600 python bench.py > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err
timeout 300 python bench.py --n 512 --steps 3 --warmup 3 --no-cpu > $O/${TAG}_bench_512.json 2> $O/${TAG}_bench_512.err
timeout 200 python tools/kernel_bench.py 256 rt  > $O/${TAG}_kb_256_rt.log 2>&1
timeout 200 python tools/kernel_bench.py 512 rt  > $O/${TAG}_kb_512_rt.log 2>&1
timeout 200 python tools/kernel_bench.py 512 tgv > $O/${TAG}_kb_512_tgv.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > $O/${TAG}_ncu_bench.log 2>&1
for op in smooth residual interp restrict; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_smooth|k_residual|k_interp|k_restrict" -c 2 \
      -f -o $O/${TAG}_prof_${op}_512 python tools/prof_op.py 512 rt $op 0 2 > $O/${TAG}_ncu_${op}.log 2>&1
done
ls -la $O
