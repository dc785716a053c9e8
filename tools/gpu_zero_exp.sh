#!/bin/bash
O=gpurun_out; TAG=${1:-r1z}; mkdir -p $O
( B200NP_ZERO_START=1 timeout 200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_project.py tests/test_gpu_composite.py tests/test_golden.py -x -q 2>&1 | tail -3 ) > $O/${TAG}_pytest_zero.log
for z in 0 1; do
  B200NP_ZERO_START=$z timeout 100 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('zero=$z 256^3 ms', round(d['ms_per_step'],3), 'min', round(d['config']['ms_per_step_min'],3), 'vc', d['config']['vcycles'], 'launches', d['gpu_launches'])"
  B200NP_ZERO_START=$z timeout 100 python bench.py --n 512 --steps 3 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('zero=$z 512^3 ms', round(d['ms_per_step'],3), 'min', round(d['config']['ms_per_step_min'],3), 'vc', d['config']['vcycles'])"
done
cat $O/${TAG}_pytest_zero.log
exit 0
