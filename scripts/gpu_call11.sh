#!/bin/bash
set -x
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -6 gpurun_out/pytest_gpu.log
export HM_BVH_CACHE=/dev/shm/hm_bvh_sweep; mkdir -p $HM_BVH_CACHE
for i in 1 2; do
timeout 300 python bench.py --no-others --no-gate --no-cpu-baseline --steps 20 --warmup 4 > gpurun_out/bench_r2l_$i.json 2> gpurun_out/bench_r2l_$i.err
python - gpurun_out/bench_r2l_$i.json <<'PY'
import json, sys
d=json.loads(open(sys.argv[1]).read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['stage_ms_per_step'])
PY
done
