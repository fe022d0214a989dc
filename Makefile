# Builds the C-ABI shared library (sm_100a only) and the three headless executables.
NVCC ?= /usr/local/cuda/bin/nvcc
CSRC := hairmsnn_b200/csrc
# A/B builds: `make lib BUILD=build_v1 LIB=hairmsnn_b200/lib/libhairmsnn_v1.so EXTRA=-DHM_TRACE_CTAS=8`, then HM_LIB=<that .so>
BUILD ?= build
LIB ?= hairmsnn_b200/lib/libhairmsnn.so
ARCH := -gencode arch=compute_100a,code=sm_100a
# -fmad=false: the traversal/intersection code must round like its host build (hit-id parity)
NVFLAGS := -O3 -std=c++17 $(ARCH) -lineinfo -Xcompiler -fPIC,-O3,-ffp-contract=off -fmad=false --expt-relaxed-constexpr -Xptxas -v $(EXTRA)
# Shading kernels (their own translation unit).  `make SHADE_APPROX=1` builds them with approximate division / square
# root: +4.2 % Mpaths/s on the bench scene, every tolerance of the north star still met, but 2-ulp quotients flip more
# of the paths' discrete choices against the reference's host build (deep-path agreement on real hair 0.91 -> 0.88,
# NRC training records 0.95 -> 0.92; profiles/r2x_*).  Default: IEEE.
SHADE_APPROX ?= 0
NVFLAGS_SHADE := $(NVFLAGS) $(if $(filter 1,$(SHADE_APPROX)),-prec-div=false -prec-sqrt=false,)
NVFLAGS_MLP := -O3 -std=c++17 $(ARCH) -lineinfo -Xcompiler -fPIC,-O3 --expt-relaxed-constexpr -Xptxas -v
CXXFLAGS := -O2 -std=c++17 -fPIC -ffp-contract=off -mfma -pthread -I/usr/local/cuda/include
# NCCL: the copy bundled with the image's PyTorch (2.28.9) when present, so a process that also imports torch
# (bench.py, the tests) maps ONE libnccl.so.2; else the system library.
NCCL_HOME ?= $(firstword $(wildcard /opt/prime-rl/.venv/lib/python3.12/site-packages/nvidia/nccl) /usr)
NCCL_INC := $(if $(filter /usr,$(NCCL_HOME)),,-I$(NCCL_HOME)/include)
NCCL_LIB := $(if $(filter /usr,$(NCCL_HOME)),-lnccl,-L$(NCCL_HOME)/lib -l:libnccl.so.2 -Xlinker -rpath,$(NCCL_HOME)/lib)
OBJ := $(BUILD)/hm_wavefront.o $(BUILD)/hm_shade_kernels.o $(BUILD)/hm_renderer.o $(BUILD)/hm_mlp.o $(BUILD)/hm_capi.o $(BUILD)/hm_io.o $(BUILD)/hm_piz.o $(BUILD)/hm_scene_util.o $(BUILD)/hm_bvh_build.o $(BUILD)/hm_comm.o
HDRS := $(wildcard $(CSRC)/*.h) $(wildcard $(CSRC)/*.cuh) include/hairmsnn.h

all: $(LIB) bin
lib: $(LIB)

$(BUILD)/hm_wavefront.o: $(CSRC)/hm_wavefront.cu $(HDRS)
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(BUILD)/hm_wavefront.ptxas.log || (cat $(BUILD)/hm_wavefront.ptxas.log; false)
$(BUILD)/hm_shade_kernels.o: $(CSRC)/hm_shade_kernels.cu $(HDRS)
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS_SHADE) -c $< -o $@ 2> $(BUILD)/hm_shade_kernels.ptxas.log || (cat $(BUILD)/hm_shade_kernels.ptxas.log; false)
$(BUILD)/hm_renderer.o: $(CSRC)/hm_renderer.cu $(HDRS)
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(BUILD)/hm_renderer.ptxas.log || (cat $(BUILD)/hm_renderer.ptxas.log; false)
$(BUILD)/hm_mlp.o: $(CSRC)/hm_mlp.cu $(HDRS)
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS_MLP) -c $< -o $@ 2> $(BUILD)/hm_mlp.ptxas.log || (cat $(BUILD)/hm_mlp.ptxas.log; false)
$(BUILD)/%.o: $(CSRC)/%.cpp $(HDRS)
	@mkdir -p $(BUILD)
	g++ $(CXXFLAGS) $(NCCL_INC) -c $< -o $@

$(LIB): $(OBJ)
	@mkdir -p hairmsnn_b200/lib
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -lz $(NCCL_LIB) -Xlinker -rpath,/usr/local/cuda/lib64

bin: hairmsnn_b200/bin/render_path_tracing hairmsnn_b200/bin/render_nrc hairmsnn_b200/bin/render_hair_msnn
hairmsnn_b200/bin/%: $(CSRC)/main_%.cpp $(LIB)
	@mkdir -p hairmsnn_b200/bin
	g++ -O2 -std=c++17 -Iinclude $< -o $@ -Lhairmsnn_b200/lib -lhairmsnn -Wl,-rpath,'$$ORIGIN/../lib'

clean:
	rm -rf build $(LIB) hairmsnn_b200/bin
