#!/bin/bash
set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 16 --warmup 4 > gpurun_out/bench_r2h_8gpu.json 2> gpurun_out/bench_r2h_8gpu.err
echo "bench 8 exit $?"; tail -3 gpurun_out/bench_r2h_8gpu.err; head -c 700 gpurun_out/bench_r2h_8gpu.json; echo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --workload straight4096 --steps 16 --warmup 4 > gpurun_out/bench_r2h_straight4096_8gpu.json 2> gpurun_out/bench_r2h_straight4096_8gpu.err
echo "bench straight4096 exit $?"; tail -3 gpurun_out/bench_r2h_straight4096_8gpu.err; head -c 900 gpurun_out/bench_r2h_straight4096_8gpu.json; echo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 4 --workload straight4096 --steps 16 --warmup 4 > gpurun_out/bench_r2h_straight4096_4gpu.json 2> gpurun_out/bench_r2h_straight4096_4gpu.err
echo "bench straight4096 x4 exit $?"; head -c 300 gpurun_out/bench_r2h_straight4096_4gpu.json; echo
timeout 300 hairmsnn_b200/bin/render_hair_msnn assets/scenes/curly/config.json 1 --spp 504 --gpus 8 --out gpurun_out/exe_8gpu.png > gpurun_out/exe_8gpu.log 2>&1
echo "exe exit $?"; tail -4 gpurun_out/exe_8gpu.log; rm -f gpurun_out/exe_8gpu*.exr
