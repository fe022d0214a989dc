#!/bin/bash
set -x
timeout 900 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_pt.py tests/test_gpu_real_scene.py -m gpu -q > gpurun_out/pytest_gpu2.log 2>&1
echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload straight4096 --size 2048 --steps 8 --warmup 3 > gpurun_out/bench_r2g_straight2048_2gpu.json 2> gpurun_out/bench_r2g_straight2048_2gpu.err
echo "bench straight exit $?"; tail -5 gpurun_out/bench_r2g_straight2048_2gpu.err; head -c 900 gpurun_out/bench_r2g_straight2048_2gpu.json
timeout 600 hairmsnn_b200/bin/render_hair_msnn assets/scenes/curly/config.json 1 --spp 32 --gpus 2 --out gpurun_out/exe_2gpu.png > gpurun_out/exe_2gpu.log 2>&1
echo "exe exit $?"; tail -5 gpurun_out/exe_2gpu.log
timeout 600 hairmsnn_b200/bin/render_path_tracing assets/scenes/curly/config.json --spp 16 --gpus 2 --shard bands --out gpurun_out/exe_pt_bands.png > gpurun_out/exe_pt_bands.log 2>&1
echo "exe exit $?"; tail -3 gpurun_out/exe_pt_bands.log
# ncu of the network kernels alone
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_mlp' -s 6 -c 4 -o gpurun_out/r2g_mlp -f python scripts/mlp_bench.py > gpurun_out/r2g_mlp_ncu.log 2>&1
echo "ncu exit $?"
