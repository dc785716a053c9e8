#!/bin/bash
mkdir -p gpurun_out
B200EB_BATCH_BELOW=0 B200EB_SMALL_NODES=0 timeout 600 python -m pytest tests/test_gpu_eb.py -q 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_eb.py -q 2>&1 | tail -2
timeout 600 python tools/eb_bench.py 512 128 128 3 > gpurun_out/r3u_eb_bench_512.json 2> gpurun_out/r3u_eb_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/r3u_eb_bench_512.json"))
print("ms", round(d["ms_per_projection"],2), "solve", round(d["ms_solve"],2), "vcycles", d["vcycles"], "resid", d["resid_over_bnorm"])
for l in d["levels"][:4]: print("  lev", l["lev"], l["nodes"], "us/sweep %.1f" % l["us_per_sweep"], "us/residual %.1f" % l["us_per_residual"], "GB/s@49 %.0f" % l["sweep_GBs_at_49B_per_node"])
PY
