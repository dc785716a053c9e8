#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mac.py tests/test_gpu_eb.py -q 2>&1 | tail -5
for pdl in 1 0; do
for n in 256 128; do
B200MAC_PDL=$pdl timeout 300 python tools/mac_bench.py $n 5 > gpurun_out/r3m_mac_bench_${n}_pdl$pdl.json 2> gpurun_out/r3m_mac.err; echo "pdl=$pdl n=$n $(cut -c95-250 gpurun_out/r3m_mac_bench_${n}_pdl$pdl.json)"
done; done
