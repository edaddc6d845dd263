#!/bin/bash
# Round-2 profiling pass (run under gpurun on ONE B200): plain runs first, then ncu on the same command lines.
#   1. launch list of the default bench command (cold-cache, serialised device times: compare SHARES)
#   2. ncu --set full of the LaneConv kernels on BASELINE config 4 (MapNet only, 100,800-node graph):
#      k_laneconv_fused (default path), k_gather_gn_relu + k_wide_tc (split path)
#   3. ncu --set full of k_actor_net / k_pred_net from the default bench
set -x
O=gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eager-reference"
C4="python bench.py --config 4 --steps 2 --warmup 3"
$B > $O/r2p_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/r2p_launches.csv $B > $O/r2p_ncu_launches.log 2>&1
$C4 > $O/r2p_plain_c4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_laneconv_fused -s 18 -c 6 -o $O/r2p_c4_fused $C4 > $O/r2p_ncu_c4a.log 2>&1
$C4 > $O/r2p_plain_c4b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_gather_gn_relu -s 6 -c 2 -o $O/r2p_c4_gather $C4 > $O/r2p_ncu_c4b.log 2>&1
$C4 > $O/r2p_plain_c4c.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_wide_tc -s 6 -c 2 -o $O/r2p_c4_wide $C4 > $O/r2p_ncu_c4c.log 2>&1
$B > $O/r2p_plain_bench2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:k_actor_net|k_pred_net" -s 4 -c 2 -o $O/r2p_actor_pred $B > $O/r2p_ncu_ap.log 2>&1
ls -la $O/r2p_*
