#!/bin/bash
set -x
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -6 gpurun_out/pytest_gpu.log
export HM_BVH_CACHE=/dev/shm/hm_bvh_sweep; mkdir -p $HM_BVH_CACHE
for v in two_level one_level two_level one_level; do
if [ $v = one_level ]; then export HM_ENV_ONE_LEVEL=1; else unset HM_ENV_ONE_LEVEL; fi
timeout 300 python bench.py --no-others --no-gate --no-cpu-baseline --steps 24 --warmup 6 > gpurun_out/bench_r2n_$v.json 2> gpurun_out/bench_r2n_$v.err
python - $v gpurun_out/bench_r2n_$v.json <<'PY'
import json, sys
d=json.loads(open(sys.argv[2]).read()); s=d['stage_ms_per_step']; print(sys.argv[1], round(d['value'],1), round(d['ms_per_step'],3), 'shade', round(s['shade'],3), 'trace', round(s['trace'],3), 'primary', round(s['primary'],3))
PY
done
