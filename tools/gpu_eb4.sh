#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/eb_bench.py 512 128 128 3 > gpurun_out/r2w_eb_bench_512.json 2> gpurun_out/r2w_eb_bench.err; cat gpurun_out/r2w_eb_bench_512.json | cut -c1-600
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_eb_gs$ --launch-skip 3 --launch-count 1 -o gpurun_out/r2w_eb_gs_l0 -f python tools/eb_bench.py 512 128 128 1 > gpurun_out/r2w_ncu1.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_eb_gs$ --launch-skip 131 --launch-count 1 -o gpurun_out/r2w_eb_gs_l1 -f python tools/eb_bench.py 512 128 128 1 > gpurun_out/r2w_ncu2.log 2>&1
tail -3 gpurun_out/r2w_ncu1.log
