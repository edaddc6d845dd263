#!/bin/bash
# Final round-2 profiling pass (run under gpurun on ONE B200): every ncu command line has first exited 0 WITHOUT ncu.
#   1. launch list of the default bench command (cold-cache, serialised device times: compare SHARES)
#   2. ncu --set full of k_laneconv_v2 at the benchmark size (batch 128: DRAM traffic per launch for bench.py's roofline)
#      and on BASELINE config 4 (MapNet only, 100,800-node graph)
#   3. ncu --set full of the head-of-forward kernels: k_actor_net, k_csr_finish, k_plan_build, k_pairs_thin
set -x
O=gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eager-reference"
C4="python bench.py --config 4 --steps 2 --warmup 3"
$B > $O/${P:-r2h}_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/${P:-r2h}_launches.csv $B > $O/${P:-r2h}_ncu_launches.log 2>&1
$B > $O/${P:-r2h}_plain_bench2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_laneconv_v2 -s 9 -c 2 -o $O/${P:-r2h}_v2_b128 $B > $O/${P:-r2h}_ncu_v2.log 2>&1
$C4 > $O/${P:-r2h}_plain_c4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_laneconv_v2 -s 9 -c 2 -o $O/${P:-r2h}_v2_c4 $C4 > $O/${P:-r2h}_ncu_c4.log 2>&1
$B > $O/${P:-r2h}_plain_bench3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:k_actor_net|k_actor_gn|k_csr_finish|k_plan_build|k_pairs_thin" -s 8 -c 8 -o $O/${P:-r2h}_head $B > $O/${P:-r2h}_ncu_head.log 2>&1
ls -la $O/${P:-r2h}_*
