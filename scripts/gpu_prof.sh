#!/bin/bash
# ncu captures of the real-scene frame: main-piece launches of frame 3 (k_primary, k_shade, k_trace, k_shade, k_trace)
set -x
TAG=${1:-r2b}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_primary|k_shade|k_trace' -s 162 -c 5 \
    -o gpurun_out/${TAG}_main -f python bench.py --steps 1 --warmup 3 --no-others --no-gate --no-cpu-baseline > gpurun_out/${TAG}_main.log 2>&1
echo "ncu main exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_mlp|k_wgrad|k_adam' -s 12 -c 7 \
    -o gpurun_out/${TAG}_mlp -f python bench.py --steps 1 --warmup 3 --no-others --no-gate --no-cpu-baseline > gpurun_out/${TAG}_mlp.log 2>&1
echo "ncu mlp exit $?"
ls -la gpurun_out/*.ncu-rep
