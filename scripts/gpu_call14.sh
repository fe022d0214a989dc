#!/bin/bash
export HM_BVH_CACHE=/dev/shm/hm_bvh_sweep; mkdir -p $HM_BVH_CACHE
run() { label=$1; shift
  env "$@" timeout 300 python bench.py --no-others --no-gate --no-cpu-baseline --steps 24 --warmup 6 > gpurun_out/r2o_$label.json 2> gpurun_out/r2o_$label.err
  python - "$label" gpurun_out/r2o_$label.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read()); s = d["stage_ms_per_step"]
    print(f"{sys.argv[1]:14s} value {d['value']:7.1f}  ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['value']:7.1f}  primary {s['primary']:.2f} shade {s['shade']:.2f} trace {s['trace']:.2f}")
except Exception as e:
    print(sys.argv[1], "no result", e)
PY
}
run base A=1
run streams2 HM_MAIN_STREAMS=2
run streams3 HM_MAIN_STREAMS=3
run streams2_fif16 HM_MAIN_STREAMS=2 HM_FRAMES_IN_FLIGHT=16
run base_b A=1
