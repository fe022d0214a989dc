#!/bin/bash
# A/B of library variants on the headline workload: scripts/sweep_variants.sh <tag> <variant> [<variant> ...]
# ("base" = the default library).  One line per variant in gpurun_out/<tag>_sweep.txt.
TAG=$1; shift
export HM_BVH_CACHE=/dev/shm/hm_bvh_sweep
mkdir -p $HM_BVH_CACHE gpurun_out
OUT=gpurun_out/${TAG}_sweep.txt
: > $OUT
for v in "$@"; do
  if [ "$v" = "base" ]; then unset HM_LIB; else export HM_LIB=$PWD/hairmsnn_b200/lib/variants/libhairmsnn_$v.so; fi
  for rep in 1 2; do
    timeout 300 python bench.py --steps 24 --warmup 6 --no-others --no-gate --no-cpu-baseline > gpurun_out/${TAG}_$v.json 2> gpurun_out/${TAG}_$v.err || echo "$v FAILED" >> $OUT
    python - "$v" gpurun_out/${TAG}_$v.json >> $OUT <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read())
    s = d["stage_ms_per_step"]
    print(f"{sys.argv[1]:10s} {d['value']:7.1f} Mpaths/s  ms/step {d['ms_per_step']:.3f}  primary {s['primary']:.2f} shade {s['shade']:.2f} trace {s['trace']:.2f} tail {s['tail_piece']:.1f} train {s['train']:.2f} infer {s['infer']:.2f}  frac {d['roofline']['frac']:.3f} nodes/ray {d['roofline']['nodes_per_ray']:.1f}")
except Exception as e:
    print(sys.argv[1], "no result", e)
PY
  done
done
cat $OUT
