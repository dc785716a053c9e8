#!/bin/bash
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_eb.py -q -x -k "stencil_hierarchy or project_parity or finest or odd or apply_nodal or set_eb_flow or rhs" > gpurun_out/r3h_sanitizer_eb.log 2>&1
echo "rc=$?"; tail -15 gpurun_out/r3h_sanitizer_eb.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_mac.py -q -x -k "project" > gpurun_out/r3h_sanitizer_mac.log 2>&1
echo "rc=$?"; tail -6 gpurun_out/r3h_sanitizer_mac.log
