#!/bin/bash
# Builds A/B variants of libhairmsnn.so with different -D switches: scripts/build_variants.sh name1 "flags1" name2 "flags2" ...
# -> hairmsnn_b200/lib/variants/libhairmsnn_<name>.so ; select at run time with HM_LIB=<path>.
set -e
mkdir -p hairmsnn_b200/lib/variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  rm -f build/hm_wavefront.o build/hm_renderer.o hairmsnn_b200/lib/libhairmsnn.so
  make -s EXTRA="$flags" hairmsnn_b200/lib/libhairmsnn.so
  cp hairmsnn_b200/lib/libhairmsnn.so hairmsnn_b200/lib/variants/libhairmsnn_$name.so
  grep -A1 "k_trace" build/hm_wavefront.ptxas.log | grep -o "Used [0-9]* registers" | head -3 | tr '\n' ' '; echo " <- $name"
done
rm -f build/hm_wavefront.o build/hm_renderer.o hairmsnn_b200/lib/libhairmsnn.so
make -s all
