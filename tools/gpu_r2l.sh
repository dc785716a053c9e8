#!/bin/bash
TAG=${1:-r2l}
O=gpurun_out; mkdir -p $O
( timeout 900 python -m pytest tests/test_gpu_mac.py tests/test_gpu_multibox.py tests/test_cpp_shim.py -m gpu -q 2>&1 | tail -40 ) > $O/${TAG}_pytest_gpu.log
tail -40 $O/${TAG}_pytest_gpu.log
exit 0
