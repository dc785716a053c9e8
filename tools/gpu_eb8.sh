#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_eb.py tests/test_cpp_shim.py -q 2>&1 | tail -5
for bb in 4000000 0 100000000; do
B200EB_BATCH_BELOW=$bb timeout 600 python tools/eb_bench.py 512 128 128 3 > gpurun_out/r3a_eb_bench_512_bb$bb.json 2> gpurun_out/r3a_eb_bench.err; echo "batch_below=$bb"; cut -c300-420 gpurun_out/r3a_eb_bench_512_bb$bb.json
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r3a_eb_launches.csv python tools/eb_bench.py 512 128 128 1 > gpurun_out/r3a_ncu_eb.log 2>&1
python tools/launch_summary.py gpurun_out/r3a_eb_launches.csv big > gpurun_out/r3a_eb_launch_list.txt 2>&1; sed -n 1,8p gpurun_out/r3a_eb_launch_list.txt;  sed -n 20,30p gpurun_out/r3a_eb_launch_list.txt
gzip -f gpurun_out/r3a_eb_launches.csv
