#!/bin/bash
mkdir -p gpurun_out
for v in 3 1; do
B200EB_BIG_VARIANT=$v B200EB_BATCH_BELOW=0 B200EB_SMALL_NODES=0 timeout 600 python -m pytest tests/test_gpu_eb.py -q -k "project_parity or level_operators" 2>&1 | tail -2
B200EB_BIG_VARIANT=$v timeout 600 python tools/eb_bench.py 512 128 128 3 > gpurun_out/r3r_eb_bench_512_v$v.json 2> gpurun_out/r3r_eb_bench.err; echo "big_variant=$v"
python - <<PY
import json
d = json.load(open("gpurun_out/r3r_eb_bench_512_v$v.json"))
print("ms", round(d["ms_per_projection"],2), "solve", round(d["ms_solve"],2), "vcycles", d["vcycles"], "resid", d["resid_over_bnorm"])
for l in d["levels"][:2]: print("  lev", l["lev"], l["nodes"], "us/sweep %.1f" % l["us_per_sweep"], "us/residual %.1f" % l["us_per_residual"], "GB/s@49 %.0f" % l["sweep_GBs_at_49B_per_node"])
PY
done
