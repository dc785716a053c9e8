#!/bin/bash
mkdir -p gpurun_out
for sn in 4096 1000 100 0; do
B200EB_SMALL_NODES=$sn timeout 600 python tools/eb_bench.py 512 128 128 3 > gpurun_out/r3l_eb_bench_512_sn$sn.json 2> gpurun_out/r3l_eb_bench.err; echo "small_nodes=$sn"
python - <<PY
import json
d = json.load(open("gpurun_out/r3l_eb_bench_512_sn$sn.json"))
print("ms", round(d["ms_per_projection"],2), "solve", round(d["ms_solve"],2), "launches", d["launches"])
for l in d["levels"][3:]: print("  lev", l["lev"], l["nodes"], "us/sweep %.1f" % l["us_per_sweep"])
PY
done
