#!/bin/bash
set -x
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
export HM_BVH_CACHE=/dev/shm/hm_bvh_sweep; mkdir -p $HM_BVH_CACHE
for i in 1 2; do
timeout 300 python bench.py --no-others --no-gate --no-cpu-baseline --steps 24 --warmup 6 > gpurun_out/bench_r2p_$i.json 2> gpurun_out/bench_r2p_$i.err
python - gpurun_out/bench_r2p_$i.json <<'PY'
import json, sys
d=json.loads(open(sys.argv[1]).read()); s=d['stage_ms_per_step']; print('RESULT', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3), 'primary', round(s['primary'],3), 'shade', round(s['shade'],3), 'trace', round(s['trace'],3))
PY
done
