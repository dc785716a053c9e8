#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_eb_gs --launch-skip 3 --launch-count 1 -o gpurun_out/r3t_eb_gs_l0 -f python tools/eb_bench.py 512 128 128 1 > gpurun_out/r3t_ncu1.log 2>&1
tail -2 gpurun_out/r3t_ncu1.log
