#!/bin/bash
bash scripts/sweep_variants.sh r2m base shade6 shade5 shade4 shade10
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_shade' -s 80 -c 2 -o gpurun_out/r2m_shade -f python bench.py --steps 1 --warmup 3 --no-others --no-gate --no-cpu-baseline > gpurun_out/r2m_shade.log 2>&1
echo "ncu exit $?"
