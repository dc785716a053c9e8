#!/bin/bash
mkdir -p gpurun_out
B200EB_BATCH_BELOW=0 B200EB_SMALL_NODES=0 timeout 600 python -m pytest tests/test_gpu_eb.py -q 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_eb.py -q 2>&1 | tail -3
run() { name=$1; var=$2; shift; shift; env "$@" timeout 600 python tools/eb_bench.py 512 128 128 3 $var > gpurun_out/r3v_eb_bench_512_$name.json 2> gpurun_out/r3v_eb_bench.err
  python - gpurun_out/r3v_eb_bench_512_$name.json $name <<'PY'
import json, sys
d = json.load(open(sys.argv[1])); l = d["levels"][0]
print(sys.argv[2], "ms/projection %.2f  solve %.2f  vcycles %d  level-0 sweep %.1f us  residual %.1f us" % (d["ms_per_projection"], d["ms_solve"], d["vcycles"], l["us_per_sweep"], l["us_per_residual"]))
PY
}
run const const X=1
run var var X=1
run var_noflags var B200EB_FLAGS=0
