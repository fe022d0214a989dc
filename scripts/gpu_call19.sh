#!/bin/bash
export HM_BVH_CACHE=/dev/shm/hm_bvh_sweep; mkdir -p $HM_BVH_CACHE
run() { label=$1; shift
  env "$@" timeout 300 python bench.py --no-others --no-gate --no-cpu-baseline --steps 24 --warmup 6 > gpurun_out/r2t_$label.json 2> gpurun_out/r2t_$label.err
  python - "$label" gpurun_out/r2t_$label.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read()); s = d["stage_ms_per_step"]
    print(f"{sys.argv[1]:14s} value {d['value']:7.1f}  ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['value']:7.1f}  tail {s['tail_piece']:.1f} trace {s['trace']:.2f}")
except Exception as e:
    print(sys.argv[1], "no result", e)
PY
}
run base A=1
run tb4096 HM_TAIL_BOUND=4096
run tb2048 HM_TAIL_BOUND=2048
run tb1024 HM_TAIL_BOUND=1024
run tb32768 HM_TAIL_BOUND=32768
run tail_lowprio HM_TAIL_LOW_PRIO=1
