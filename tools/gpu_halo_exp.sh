#!/bin/bash
N=${1:-2}; TAG=${2:-r1h}; O=gpurun_out; mkdir -p $O
run() { name=$1; shift
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
     tools/kernel_bench_dist.py 256 > $O/${TAG}_kbd_$name.log 2>&1; grep -A6 "^rank 0" $O/${TAG}_kbd_$name.log; }
run dbg1_noflags        B200NP_DIST_MIN_PLANES=8 B200NP_DBG_HALO=1
run dbg2_nopush         B200NP_DIST_MIN_PLANES=8 B200NP_DBG_HALO=2
run dbg3_nopush_notick  B200NP_DIST_MIN_PLANES=8 B200NP_DBG_HALO=3
run dbg4_push_notick    B200NP_DIST_MIN_PLANES=8 B200NP_DBG_HALO=4
run dbg5_plain_nohalo   B200NP_DIST_MIN_PLANES=8 B200NP_DBG_HALO=5 B200NP_FUSE_HALO=0
timeout 100 python tools/kernel_bench.py 256 rt > $O/${TAG}_kb1.log 2>&1; grep smooth $O/${TAG}_kb1.log
exit 0
