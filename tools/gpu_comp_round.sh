#!/bin/bash
# composite solver on the GPU: parity tests + configs[3] at full size.  Usage: bash tools/gpu_comp_round.sh tag
TAG=${1:-r1c}; O=gpurun_out; mkdir -p $O
( timeout 600 python -m pytest tests/test_gpu_composite.py -x -q 2>&1 | tail -30 ) > $O/${TAG}_pytest_comp.log
timeout 300 python tools/composite_bench.py 256 3 > $O/${TAG}_comp_bench.log 2>&1
tail -5 $O/${TAG}_pytest_comp.log; tail -4 $O/${TAG}_comp_bench.log
exit 0
