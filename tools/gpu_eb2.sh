#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_eb.py -q 2>&1 | tail -60 > gpurun_out/r2t_pytest_eb.log
tail -40 gpurun_out/r2t_pytest_eb.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2t_eb_launches.csv python tools/eb_bench.py 512 128 128 1 > gpurun_out/r2t_ncu_eb.log 2>&1
python tools/launch_summary.py gpurun_out/r2t_eb_launches.csv big > gpurun_out/r2t_eb_launch_list.txt 2>&1; head -40 gpurun_out/r2t_eb_launch_list.txt
gzip -f gpurun_out/r2t_eb_launches.csv
