#!/bin/bash
# EB: one-CTA smoother with the level vector in shared memory (k_eb_gs_smem) -- parity tests with the threshold at 0 / default / max,
# then configs[4] timing per threshold.  Usage: gpurun -- bash tools/gpu_eb_smem_ab.sh
O=gpurun_out; mkdir -p $O
for T in 28000 3000; do B200EB_SMEM_NODES=$T timeout 600 python -m pytest tests/test_gpu_eb.py -q -x 2>&1 | tail -2; done
run() { name=$1; shift; env "$@" timeout 600 python tools/eb_bench.py 512 128 128 3 > $O/ebs_bench_512_$name.json 2> $O/ebs_bench.err
  python - $O/ebs_bench_512_$name.json $name <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print(sys.argv[2], "ms/projection %.2f  solve %.2f  vcycles %d  launches %d  sweeps us:" % (d["ms_per_projection"], d["ms_solve"], d["vcycles"], d["launches"]), [round(l["us_per_sweep"], 1) for l in d["levels"]])
PY
}
run smem0 B200EB_SMEM_NODES=0
run smem1000 B200EB_SMEM_NODES=1000
run smem3000 B200EB_SMEM_NODES=3000
run smem20000 B200EB_SMEM_NODES=20000
