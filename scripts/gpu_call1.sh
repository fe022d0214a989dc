#!/bin/bash
# first GPU call of round 2: tcnn golden vectors, headline-scene image gate, GPU test-suite
set -x
mkdir -p gpurun_out/tcnn12 gpurun_out/tcnn9 gpurun_out/golden
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
for ch in 12 9; do
  timeout 300 oracle/_ref/tcnn_golden assets/scenes/curly/tcnn_hairmsnn.json $ch gpurun_out/tcnn$ch > gpurun_out/tcnn$ch.log 2>&1
  echo "tcnn_golden $ch exit $?"
  python oracle/make_tcnn_golden.py gpurun_out/tcnn$ch $ch gpurun_out/golden/tcnn_$ch.npz >> gpurun_out/tcnn$ch.log 2>&1
  echo "reduce $ch exit $?"
  rm -f gpurun_out/tcnn$ch/params*_f16.bin gpurun_out/tcnn$ch/params1_f32.bin gpurun_out/tcnn$ch/params_reset1_f32.bin gpurun_out/tcnn$ch/inputs.bin
done
timeout 900 python scripts/curly_gate.py --nrc > gpurun_out/curly_gate.log 2>&1
echo "curly gate exit $?"
tail -40 gpurun_out/curly_gate.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"
tail -15 gpurun_out/pytest_gpu.log
