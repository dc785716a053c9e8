#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_eb.py -q 2>&1 | tail -5
timeout 600 python tools/eb_bench.py 512 128 128 3 > gpurun_out/r2z_eb_bench_512.json 2> gpurun_out/r2z_eb_bench.err; cut -c1-420 gpurun_out/r2z_eb_bench_512.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2z_eb_launches.csv python tools/eb_bench.py 512 128 128 1 > gpurun_out/r2z_ncu_eb.log 2>&1
python tools/launch_summary.py gpurun_out/r2z_eb_launches.csv big > gpurun_out/r2z_eb_launch_list.txt 2>&1; sed -n 1,8p gpurun_out/r2z_eb_launch_list.txt;  sed -n 20,30p gpurun_out/r2z_eb_launch_list.txt
gzip -f gpurun_out/r2z_eb_launches.csv
