// ref_pt.cpp — TEST INFRASTRUCTURE (oracle/_ref).  Builds the reference's own
// path-tracing device program (cuda/path_tracing.cu + cuda_headers/*.cuh, included
// from /root/reference where they lie) for the host.  See ref_optix_emul.h.
#include "ref_optix_emul.h"

// common.cuh only declares the RadianceRay/ShadowRay typedefs for device passes
#define __CUDA_ARCH__ 860
#include "path_tracing.cuh"
#undef __CUDA_ARCH__
#include "utils.cuh"
#include "curve_utils.cuh"
#include "disney_hair.cuh"
#include "frostbite_anisotropic.cuh"
#include "ref_draw_order.h"
#define lcg_randomf(r) refemu::ordered_draw((r), __FILE__, __LINE__)
#include "optix_common.cuh"
#include "path_tracing.cu"

#include "ref_exports.inc"

extern "C" {

// Runs the reference ray-generation program for pixels [x0,x1) x [y0,y1) of a W x H
// frame at sample index accum_id.  accum/average are float4[W*H], fb is uint32[W*H].
void ref_render_pt(int accum_id, int x0, int y0, int x1, int y1, int W, int H,
                   float* accum, float* average, uint32_t* fb, int threads) {
    optixLaunchParams.accumId = accum_id;
    optixLaunchParams.accumBuffer = (float4*)accum;
    optixLaunchParams.averageBuffer = (float4*)average;
    g_raygen_data.frameBuffer = fb;
    g_raygen_data.frameBufferSize = vec2i(W, H);
    if (threads < 1) threads = 1;
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) {
        pool.emplace_back([=]() {
            for (int y = y0 + t; y < y1; y += threads)
                for (int x = x0; x < x1; ++x) {
                    refemu::g_ctx.launch_x = x; refemu::g_ctx.launch_y = y;
                    refemu::g_ctx.program_data = &g_raygen_data;
                    ref_raygen_rayGenCam();
                }
        });
    }
    for (auto& th : pool) th.join();
}

}  // extern "C"
