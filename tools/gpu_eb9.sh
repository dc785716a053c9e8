#!/bin/bash
mkdir -p gpurun_out
for bb in 4000000 0 100000000; do
B200EB_BATCH_BELOW=$bb timeout 600 python tools/eb_bench.py 512 128 128 3 > gpurun_out/r3b_eb_bench_512_bb$bb.json 2> gpurun_out/r3b_eb_bench.err; echo "batch_below=$bb"
python - <<PY
import json
d = json.load(open("gpurun_out/r3b_eb_bench_512_bb$bb.json"))
print("ms", round(d["ms_per_projection"],2), "solve", round(d["ms_solve"],2))
for l in d["levels"]: print("  lev", l["lev"], l["nodes"], "us/sweep %.1f" % l["us_per_sweep"], "us/residual %.1f" % l["us_per_residual"], "GB/s@49 %.0f" % l["sweep_GBs_at_49B_per_node"])
PY
done
