#!/bin/bash
export HM_BVH_CACHE=/dev/shm/hm_bvh_sweep; mkdir -p $HM_BVH_CACHE
run() { # label, env..., args
  label=$1; shift
  env "$@" timeout 300 python bench.py --no-others --no-gate --no-cpu-baseline $ARGS > gpurun_out/r2k_$label.json 2> gpurun_out/r2k_$label.err
  python - "$label" gpurun_out/r2k_$label.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read()); s = d["stage_ms_per_step"]
    print(f"{sys.argv[1]:22s} value {d['value']:7.1f}  ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['value']:7.1f}  steps {d['steps']}  tail {s['tail_piece']:.1f} trace {s['trace']:.2f}")
except Exception as e:
    print(sys.argv[1], "no result", e)
PY
}
ARGS="--steps 16 --warmup 4"
run base16 A=1
run mega16 HM_TAIL_MEGA=1
run fif4_16 HM_FRAMES_IN_FLIGHT=4
run fif12_16 HM_FRAMES_IN_FLIGHT=12
ARGS="--steps 64 --warmup 8"
run base64 A=1
run mega64 HM_TAIL_MEGA=1
ARGS="--steps 20 --warmup 3"
run base20 A=1
run mega20 HM_TAIL_MEGA=1
