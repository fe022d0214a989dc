#!/bin/bash
set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -25 gpurun_out/pytest_gpu.log
timeout 300 python scripts/mlp_bench.py > gpurun_out/mlp_bench_r2e.txt 2>&1
HM_MLP_TRAIN=split timeout 300 python scripts/mlp_bench.py >> gpurun_out/mlp_bench_r2e.txt 2>&1
cat gpurun_out/mlp_bench_r2e.txt
