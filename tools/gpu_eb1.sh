#!/bin/bash
# first GPU visit of the EB path: parity tests, then the configs[4] timing at two sizes
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_eb.py -q -x 2>&1 | tail -40 > gpurun_out/r2s_pytest_eb.log
cat gpurun_out/r2s_pytest_eb.log | tail -25
timeout 300 python tools/eb_bench.py 128 32 32 3 > gpurun_out/r2s_eb_bench_128.json 2> gpurun_out/r2s_eb_bench_128.err; cat gpurun_out/r2s_eb_bench_128.json; tail -3 gpurun_out/r2s_eb_bench_128.err
timeout 600 python tools/eb_bench.py 512 128 128 3 > gpurun_out/r2s_eb_bench_512.json 2> gpurun_out/r2s_eb_bench_512.err; cat gpurun_out/r2s_eb_bench_512.json; tail -3 gpurun_out/r2s_eb_bench_512.err
