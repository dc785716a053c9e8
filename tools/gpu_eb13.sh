#!/bin/bash
mkdir -p gpurun_out
B200EB_TILE_ZC=8 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r3j_eb_launches.csv python tools/eb_bench.py 512 128 128 1 > gpurun_out/r3j_ncu_eb.log 2>&1
python tools/launch_summary.py gpurun_out/r3j_eb_launches.csv big > gpurun_out/r3j_eb_launch_list.txt 2>&1; grep "k_eb_gs" gpurun_out/r3j_eb_launch_list.txt | head
B200EB_TILE_ZC=8 timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_eb_gs_tile --launch-skip 3 --launch-count 1 -o gpurun_out/r3j_eb_gs_tile -f python tools/eb_bench.py 512 128 128 1 > gpurun_out/r3j_ncu2.log 2>&1
rm -f gpurun_out/r3j_eb_launches.csv
