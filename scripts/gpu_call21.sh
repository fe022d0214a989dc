#!/bin/bash
# pooled solver: parity first (hit ids bit-exact vs the host build, real hair), then A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pt.py tests/test_gpu_real_scene.py -x -q 2>&1 | tail -8 | tee gpurun_out/r2v_tests.txt
bash scripts/sweep_variants.sh r2v nopool base pool16 pool32
