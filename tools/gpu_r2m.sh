#!/bin/bash
# 1-GPU visit: MAC / shim tests, MAC bench, ncu launch list of the bench command + one full capture of the level-0 smoother
TAG=${1:-r2m}
O=gpurun_out; mkdir -p $O
( timeout 600 python -m pytest tests/test_gpu_mac.py tests/test_cpp_shim.py -m gpu -q 2>&1 | tail -8 ) > $O/${TAG}_pytest_gpu.log
tail -4 $O/${TAG}_pytest_gpu.log
timeout 300 python tools/mac_bench.py 256 5 > $O/${TAG}_mac_bench_256.json 2> $O/${TAG}_mac_bench_256.err; cat $O/${TAG}_mac_bench_256.json; tail -3 $O/${TAG}_mac_bench_256.err
timeout 300 python tools/mac_bench.py 128 5 > $O/${TAG}_mac_bench_128.json 2>/dev/null; cat $O/${TAG}_mac_bench_128.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-parity > $O/${TAG}_ncu_bench.log 2>&1
python tools/launch_summary.py $O/${TAG}_launches.csv > $O/${TAG}_launch_list.txt 2>&1; head -25 $O/${TAG}_launch_list.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_smooth_iso -s 40 -c 1 -o $O/${TAG}_prof_smooth_256 -f \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-parity > $O/${TAG}_ncu_smooth.log 2>&1
ncu -i $O/${TAG}_prof_smooth_256.ncu-rep --page raw --csv > $O/${TAG}_ncu_smooth_raw.csv 2>/dev/null
python tools/ncu_summary.py $O/${TAG}_ncu_smooth_raw.csv 2>/dev/null | head -40
gzip -f $O/${TAG}_launches.csv
exit 0
