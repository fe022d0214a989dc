#!/bin/bash
# A/B of library variants on the bench workload: scripts/ab.sh name1 name2 ... (names under hairmsnn_b200/lib/variants; "base" = the default build)
for v in "$@"; do
  if [ "$v" = base ]; then lib=""; else lib=$PWD/hairmsnn_b200/lib/variants/libhairmsnn_$v.so; fi
  HM_LIB=$lib python bench.py --steps 24 --warmup 6 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v', round(d['value'],1),'Mpaths/s', round(d['ms_per_step'],3),'ms', {k:round(v,2) for k,v in d['stage_ms_per_step'].items()})"
done
