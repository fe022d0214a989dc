#!/bin/bash
# Captures `ncu --set full` reports of the hot kernels of one bench frame (run under gpurun, one GPU).
# usage: scripts/ncu_capture.sh <tag> [kernels...]   -> gpurun_out/<tag>_<kernel>.ncu-rep
TAG=${1:-r1}; shift
KERNELS=${@:-"k_trace k_primary k_shade k_mlp_forward"}
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
NCU="ncu --set full --clock-control none --import-source on -f"
for k in $KERNELS; do
  case $k in
    # frame 3 (after 3 warm-up frames): 40 k_trace / k_shade launches per frame
    k_trace) $NCU -k regex:k_trace --launch-skip 120 --launch-count 2 -o $OUT/${TAG}_k_trace $BENCH > $OUT/${TAG}_ncu_trace.log 2>&1 ;;
    k_primary) $NCU -k regex:k_primary --launch-skip 3 --launch-count 1 -o $OUT/${TAG}_k_primary $BENCH > $OUT/${TAG}_ncu_primary.log 2>&1 ;;
    k_shade) $NCU -k regex:k_shade --launch-skip 120 --launch-count 2 -o $OUT/${TAG}_k_shade $BENCH > $OUT/${TAG}_ncu_shade.log 2>&1 ;;
    k_mlp_forward) $NCU -k regex:k_mlp_forward_tc --launch-skip 6 --launch-count 2 -o $OUT/${TAG}_k_mlp_forward $BENCH > $OUT/${TAG}_ncu_mlp.log 2>&1 ;;
  esac
done
ls -la $OUT/*.ncu-rep
