#!/bin/bash
set -x
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -40 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline --no-gate > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err
echo "bench exit $?"; tail -4 gpurun_out/bench_r2f.err; head -c 600 gpurun_out/bench_r2f.json
