#!/bin/bash
# A/B sweep on the bench workload (run under gpurun): BVH build parameters (run-time env) x library variants.
run() { # label, env..., lib
  label=$1; shift
  env "$@" python bench.py --steps 24 --warmup 6 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$label', round(d['value'],1),'Mpaths/s', round(d['ms_per_step'],3),'ms nodes/ray', round(r['nodes_per_ray'],1), 'prims/ray', round(r['prims_per_ray'],2), {k:round(v,2) for k,v in d['stage_ms_per_step'].items()})"
}
V=$PWD/hairmsnn_b200/lib/variants
