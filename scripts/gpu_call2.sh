#!/bin/bash
set -x
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err
echo "bench exit $?"; tail -5 gpurun_out/bench_r2a.err; cat gpurun_out/bench_r2a.json | head -c 3000
timeout 600 python bench.py --impl reference --steps 8 --warmup 2 > gpurun_out/bench_r2a_ref.json 2> gpurun_out/bench_r2a_ref.err
echo "ref exit $?"; cat gpurun_out/bench_r2a_ref.json | head -c 1500
