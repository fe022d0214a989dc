#!/bin/bash
# A/B build of the wavefront kernels: scripts/build_variant.sh <name> "<extra nvcc flags>" ["<extra flags for the shading TU>"]
# -> hairmsnn_b200/lib/variants/libhairmsnn_<name>.so (other objects reused from build/); select it with HM_LIB=<path>.
set -e
NAME=$1; EXTRA=$2; EXTRA_SHADE=${3:-}
cd "$(dirname "$0")/.."
make -s lib
mkdir -p build_var/$NAME hairmsnn_b200/lib/variants
NVCC=/usr/local/cuda/bin/nvcc
ARCH="-gencode arch=compute_100a,code=sm_100a"
$NVCC -O3 -std=c++17 $ARCH -lineinfo -Xcompiler -fPIC,-O3,-ffp-contract=off -fmad=false --expt-relaxed-constexpr -Xptxas -v $EXTRA \
   -c hairmsnn_b200/csrc/hm_wavefront.cu -o build_var/$NAME/hm_wavefront.o 2> build_var/$NAME/ptxas.log || { cat build_var/$NAME/ptxas.log; exit 1; }
$NVCC -O3 -std=c++17 $ARCH -lineinfo -Xcompiler -fPIC,-O3,-ffp-contract=off -fmad=false --expt-relaxed-constexpr -Xptxas -v $EXTRA $EXTRA_SHADE \
   -c hairmsnn_b200/csrc/hm_shade_kernels.cu -o build_var/$NAME/hm_shade_kernels.o 2> build_var/$NAME/ptxas_shade.log || { cat build_var/$NAME/ptxas_shade.log; exit 1; }
NCCL=/opt/prime-rl/.venv/lib/python3.12/site-packages/nvidia/nccl
OBJS="build_var/$NAME/hm_wavefront.o build_var/$NAME/hm_shade_kernels.o build/hm_renderer.o build/hm_mlp.o build/hm_capi.o build/hm_io.o build/hm_piz.o build/hm_scene_util.o build/hm_bvh_build.o build/hm_comm.o"
$NVCC $ARCH -shared -o hairmsnn_b200/lib/variants/libhairmsnn_$NAME.so $OBJS -lz -L$NCCL/lib -l:libnccl.so.2 -Xlinker -rpath,$NCCL/lib -Xlinker -rpath,/usr/local/cuda/lib64
grep -A1 "k_shade\b\|k_traceE\|k_primary\|7k_shade\|7k_trace" build_var/$NAME/ptxas.log | grep -o "Used [0-9]* registers.*" | head -5
