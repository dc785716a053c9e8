#!/bin/bash
O=gpurun_out; TAG=${1:-r1n}; mkdir -p $O
( timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -4 ) > $O/${TAG}_pytest_gpu.log
timeout 300 python bench.py > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err
timeout 100 python tools/composite_bench.py 256 2 > $O/${TAG}_comp.log 2>&1
cat $O/${TAG}_pytest_gpu.log; head -c 330 $O/${TAG}_bench_n1.json; echo; tail -2 $O/${TAG}_comp.log
exit 0
