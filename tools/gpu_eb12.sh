#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_eb.py -q 2>&1 | tail -8
for zc in 8 4 16 0; do
B200EB_TILE_ZC=$zc timeout 600 python tools/eb_bench.py 512 128 128 3 > gpurun_out/r3i_eb_bench_512_zc$zc.json 2> gpurun_out/r3i_eb_bench.err; echo "tile_zc=$zc"
python - <<PY
import json
d = json.load(open("gpurun_out/r3i_eb_bench_512_zc$zc.json"))
print("ms", round(d["ms_per_projection"],2), "solve", round(d["ms_solve"],2), "vcycles", d["vcycles"], "resid", d["resid_over_bnorm"])
for l in d["levels"][:2]: print("  lev", l["lev"], l["nodes"], "us/sweep %.1f" % l["us_per_sweep"], "us/residual %.1f" % l["us_per_residual"], "GB/s@49 %.0f" % l["sweep_GBs_at_49B_per_node"])
PY
done
