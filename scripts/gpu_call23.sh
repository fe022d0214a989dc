#!/bin/bash
# pooled prim step (HM_TRACE_POOL=2): parity on one variant, then A/B
mkdir -p gpurun_out
HM_LIB=$PWD/hairmsnn_b200/lib/variants/libhairmsnn_pp24.so timeout 900 python -m pytest tests/test_gpu_pt.py tests/test_gpu_real_scene.py -x -q 2>&1 | tail -6 | tee gpurun_out/r2ab_tests.txt
bash scripts/sweep_variants.sh r2ab base pp16 pp24 pp32
