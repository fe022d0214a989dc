#!/bin/bash
# merged tail pieces: parity of the new schedule, then the A/B on the bench workload
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_msnn.py -x -q 2>&1 | tail -15 > gpurun_out/r2u_tests.txt
cat gpurun_out/r2u_tests.txt
timeout 600 python scripts/tail_group_sweep.py 24 "1:8,4:12,2:8,4:16,8:16,8:24,1:8,4:12" 2> gpurun_out/r2u_sweep.err | grep "^group" | tee gpurun_out/r2u_tail_group.txt
tail -5 gpurun_out/r2u_sweep.err
