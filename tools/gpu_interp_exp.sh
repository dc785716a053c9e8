#!/bin/bash
O=gpurun_out; TAG=${1:-r1m}; mkdir -p $O
( B200NP_INTERP_TZ=4 timeout 200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_composite.py -x -q 2>&1 | tail -3 ) > $O/${TAG}_pytest_tz4.log
for tz in 8 4; do
  B200NP_INTERP_TZ=$tz timeout 100 python tools/kernel_bench.py 512 rt 2>&1 | grep -E "interp|iters" > $O/${TAG}_kb512_tz$tz.log
  B200NP_INTERP_TZ=$tz timeout 100 python tools/kernel_bench.py 512 tgv 2>&1 | grep -E "interp|iters" >> $O/${TAG}_kb512_tz$tz.log
  B200NP_INTERP_TZ=$tz timeout 100 python tools/kernel_bench.py 256 rt 2>&1 | grep -E "interp|iters" >> $O/${TAG}_kb512_tz$tz.log
  echo "== tz $tz"; cat $O/${TAG}_kb512_tz$tz.log
done
cat $O/${TAG}_pytest_tz4.log
exit 0
