#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/r2x_tests.txt
bash scripts/sweep_variants.sh r2x base
