#!/bin/bash
set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --workload straight4096 --beta-sweep 1,2,4,8,11 --steps 12 --warmup 3 > gpurun_out/bench_r2s_straight4096_sweep.json 2> gpurun_out/bench_r2s_straight4096_sweep.err
echo "exit $?"; grep "beta sweep" gpurun_out/bench_r2s_straight4096_sweep.err
