#!/bin/bash
# EB path on one B200: parity tests with every kernel variant forced, then the configs[4] timing under the A/B switches of
# csrc/b200eb.cu (B200EB_FLAGS canonical-row flags, B200EB_BIG_VARIANT finest-level kernel, B200EB_BATCH_BELOW level threshold of the
# load-batching kernels, B200EB_SMALL_NODES one-CTA smoother threshold, B200EB_PDL programmatic dependent launch), an ncu launch list and a
# full capture of one colour launch of the finest level.  Usage: gpurun -- bash tools/gpu_eb_ab.sh [tag]
TAG=${1:-eb}
O=gpurun_out; mkdir -p $O
B200EB_BATCH_BELOW=0 B200EB_SMALL_NODES=0 timeout 600 python -m pytest tests/test_gpu_eb.py -q 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_eb.py -q 2>&1 | tail -2
run() { name=$1; shift; env "$@" timeout 600 python tools/eb_bench.py 512 128 128 3 > $O/${TAG}_bench_512_$name.json 2> $O/${TAG}_bench.err
  python - $O/${TAG}_bench_512_$name.json $name <<'PY'
import json, sys
d = json.load(open(sys.argv[1])); l = d["levels"][0]
print(sys.argv[2], "ms/projection %.2f  solve %.2f  vcycles %d  level-0 sweep %.1f us" % (d["ms_per_projection"], d["ms_solve"], d["vcycles"], l["us_per_sweep"]))
PY
}
run default
run flags0 B200EB_FLAGS=0
run big1 B200EB_BIG_VARIANT=1
run big0 B200EB_BIG_VARIANT=0
run batch4m B200EB_BATCH_BELOW=4000000
run small4096 B200EB_SMALL_NODES=4096
run pdl0 B200EB_PDL=0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/${TAG}_launches.csv python tools/eb_bench.py 512 128 128 1 > $O/${TAG}_ncu.log 2>&1
python tools/launch_summary.py $O/${TAG}_launches.csv big > $O/${TAG}_launch_list.txt 2>&1; head -12 $O/${TAG}_launch_list.txt; gzip -f $O/${TAG}_launches.csv
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_eb_gs --launch-skip 3 --launch-count 1 -o $O/${TAG}_gs_l0 -f python tools/eb_bench.py 512 128 128 1 > $O/${TAG}_ncu2.log 2>&1
