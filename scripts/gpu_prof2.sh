#!/bin/bash
# ncu capture of the real-scene frame with merged tail pieces: warm-up frames 0-2 enqueue 3 x 5 main-piece launches, their
# merged tail 38 x (k_shade + k_trace); the timed frame's main piece (k_primary, k_shade, k_trace, k_shade, k_trace) follows
TAG=${1:-r2aa}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_primary|k_shade|k_trace' -s 91 -c 5 \
    -o gpurun_out/${TAG}_main -f python bench.py --steps 1 --warmup 3 --no-others --no-gate --no-cpu-baseline > gpurun_out/${TAG}_main.log 2>&1
echo "ncu main exit $?"
ls -la gpurun_out/*.ncu-rep
