#!/bin/bash
set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 4 > gpurun_out/bench_r2r_8gpu.json 2> gpurun_out/bench_r2r_8gpu.err
echo "bench 8 exit $?"; tail -2 gpurun_out/bench_r2r_8gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 20 --warmup 4 > gpurun_out/bench_r2r_4gpu.json 2> gpurun_out/bench_r2r_4gpu.err
echo "bench 4 exit $?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --workload straight4096 --steps 20 --warmup 4 > gpurun_out/bench_r2r_straight4096_8gpu.json 2> gpurun_out/bench_r2r_straight4096_8gpu.err
echo "bench straight exit $?"
python - <<'PY'
import json
for f in ("bench_r2r_8gpu","bench_r2r_4gpu","bench_r2r_straight4096_8gpu"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read())
        print("RESULT", f, round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'loss', d['training_loss'])
    except Exception as e: print("RESULT", f, "failed", e)
PY
