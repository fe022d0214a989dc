#!/bin/bash
set -x
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -6 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_r2j.json 2> gpurun_out/bench_r2j.err
echo "bench exit $?"; tail -3 gpurun_out/bench_r2j.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2j.json').read())
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['mlp'], d['cpu_baseline']['value'], d['stage_ms_per_step'])
PY
timeout 300 python scripts/mlp_bench.py > gpurun_out/mlp_bench_r2j.txt 2>&1; cat gpurun_out/mlp_bench_r2j.txt
HM_MLP_CTAS=3 timeout 300 python scripts/mlp_bench.py 2>&1 | head -1
