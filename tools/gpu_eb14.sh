#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_eb.py -q 2>&1 | tail -5
for pdl in 1 0; do
B200EB_PDL=$pdl timeout 600 python tools/eb_bench.py 512 128 128 3 > gpurun_out/r3k_eb_bench_512_pdl$pdl.json 2> gpurun_out/r3k_eb_bench.err; echo "pdl=$pdl"
python - <<PY
import json
d = json.load(open("gpurun_out/r3k_eb_bench_512_pdl$pdl.json"))
print("ms", round(d["ms_per_projection"],2), "solve", round(d["ms_solve"],2), "vcycles", d["vcycles"], "resid", d["resid_over_bnorm"])
for l in d["levels"][:5]: print("  lev", l["lev"], l["nodes"], "us/sweep %.1f" % l["us_per_sweep"], "us/residual %.1f" % l["us_per_residual"])
PY
done
