#!/bin/bash
set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -25 gpurun_out/pytest_gpu.log
timeout 300 python scripts/mlp_bench.py > gpurun_out/mlp_bench_r2c.txt 2>&1; cat gpurun_out/mlp_bench_r2c.txt
for n in 2; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 16 --warmup 4 > gpurun_out/bench_r2c_${n}gpu.json 2> gpurun_out/bench_r2c_${n}gpu.err
echo "bench $n exit $?"; tail -3 gpurun_out/bench_r2c_${n}gpu.err; head -c 1200 gpurun_out/bench_r2c_${n}gpu.json
done
