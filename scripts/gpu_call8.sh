#!/bin/bash
set -x
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -12 gpurun_out/pytest_gpu.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_r2i.json 2> gpurun_out/bench_r2i.err
echo "bench exit $?"; tail -6 gpurun_out/bench_r2i.err; head -c 400 gpurun_out/bench_r2i.json; echo
timeout 300 python scripts/tail_probe.py > gpurun_out/tail_probe_r2i.txt 2>&1; cat gpurun_out/tail_probe_r2i.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_mlp' -s 6 -c 3 -o gpurun_out/r2i_mlp -f python scripts/mlp_bench.py > gpurun_out/r2i_mlp_ncu.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/*.ncu-rep
