#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mac.py tests/test_cpp_shim.py -q 2>&1 | tail -4
timeout 300 python tools/mac_bench.py 256 5 > gpurun_out/r3g_mac_bench_256.json 2> gpurun_out/r3g_mac.err; cut -c1-330 gpurun_out/r3g_mac_bench_256.json
timeout 300 python tools/mac_bench.py 128 5 > gpurun_out/r3g_mac_bench_128.json 2>> gpurun_out/r3g_mac.err; cut -c1-330 gpurun_out/r3g_mac_bench_128.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r3g_mac_launches.csv python tools/mac_bench.py 256 1 > gpurun_out/r3g_ncu_mac.log 2>&1
python tools/launch_summary.py gpurun_out/r3g_mac_launches.csv big > gpurun_out/r3g_mac_launch_list.txt 2>&1; grep "k_mac" gpurun_out/r3g_mac_launch_list.txt | head -24
gzip -f gpurun_out/r3g_mac_launches.csv
