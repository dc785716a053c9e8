#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r3f_mac_launches.csv python tools/mac_bench.py 256 1 > gpurun_out/r3f_ncu_mac.log 2>&1
python tools/launch_summary.py gpurun_out/r3f_mac_launches.csv big > gpurun_out/r3f_mac_launch_list.txt 2>&1; head -45 gpurun_out/r3f_mac_launch_list.txt
gzip -f gpurun_out/r3f_mac_launches.csv
