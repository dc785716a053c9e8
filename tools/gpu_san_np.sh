#!/bin/bash
# compute-sanitizer over the nodal path (b200np_*): memcheck on the level kernels (resident and ring-slot smoother, zero start, residual,
# restriction, interpolation, V-cycle routing) and on whole projections (every BC case, both top-level forms); racecheck (shared-memory
# hazards) on the ring-slot smoother, the residual and the interpolation tile.  Usage: gpurun -- bash tools/gpu_san_np.sh [tag]
TAG=${1:-r2c}
O=gpurun_out; mkdir -p $O
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_kernels.py -q -x -p no:cacheprovider \
   -k "(nonresident and True-4) or test_residual or test_restriction or test_interpolation or (routing and rt_walls_z) or test_vcycle or zero_start_resident" > $O/${TAG}_sanitizer_np_kernels.log 2>&1
echo "memcheck kernels rc=$?"; tail -4 $O/${TAG}_sanitizer_np_kernels.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_project.py -q -x -p no:cacheprovider -k "bc_cases or (parity and 32)" > $O/${TAG}_sanitizer_np_project.log 2>&1
echo "memcheck project rc=$?"; tail -4 $O/${TAG}_sanitizer_np_project.log
timeout 420 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_kernels.py -q -x -p no:cacheprovider \
   -k "(nonresident and rt_walls_z-True-4) or (test_residual and rt_walls_z) or (test_interpolation and rt_walls_z) or (test_smoother_sweeps and rt_walls_z-True-4)" > $O/${TAG}_racecheck_np.log 2>&1
echo "racecheck rc=$?"; tail -4 $O/${TAG}_racecheck_np.log
