#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_eb.py -q 2>&1 | tail -5
for cfg in "1 4096" "0 4096" "1 20000"; do
  set -- $cfg
  B200EB_FLAGS=$1 B200EB_SMALL_NODES=$2 timeout 600 python tools/eb_bench.py 512 128 128 3 > gpurun_out/r2x_eb_bench_512_f$1_s$2.json 2> gpurun_out/r2x_eb_bench.err
  python - <<PY
import json
d = json.load(open("gpurun_out/r2x_eb_bench_512_f$1_s$2.json"))
print("flags=$1 small=$2", "ms", round(d["ms_per_projection"],2), "solve", round(d["ms_solve"],2), "vcycles", d["vcycles"], "launches", d["launches"], "resid", d["resid_over_bnorm"])
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2x_eb_launches.csv python tools/eb_bench.py 512 128 128 1 > gpurun_out/r2x_ncu_eb.log 2>&1
python tools/launch_summary.py gpurun_out/r2x_eb_launches.csv big > gpurun_out/r2x_eb_launch_list.txt 2>&1; head -32 gpurun_out/r2x_eb_launch_list.txt
gzip -f gpurun_out/r2x_eb_launches.csv
