#!/bin/bash
# ncu --set full of the two hot MAC kernels on the finest level (256^3), and of the EB residual
O=gpurun_out; mkdir -p $O
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_mac_gsrb --launch-skip 1 --launch-count 1 -o $O/r3x_mac_gsrb_l0 -f python tools/mac_bench.py 256 1 > $O/r3x_ncu1.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_mac_residual --launch-skip 0 --launch-count 1 -o $O/r3x_mac_residual_l0 -f python tools/mac_bench.py 256 1 > $O/r3x_ncu2.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_eb_residual --launch-skip 0 --launch-count 1 -o $O/r3x_eb_residual_l0 -f python tools/eb_bench.py 512 128 128 1 > $O/r3x_ncu3.log 2>&1
for f in $O/r3x_ncu1.log $O/r3x_ncu2.log $O/r3x_ncu3.log; do tail -n 1 $f; done
