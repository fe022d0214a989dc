// ref_msnn.cpp — TEST INFRASTRUCTURE (oracle/_ref).  Builds the reference's own HairMSNN
// device program (cuda/hair_msnn.cu + cuda_headers/*.cuh, included from /root/reference
// where they lie) for the host.  See ref_optix_emul.h.
#include "ref_optix_emul.h"

#include <atomic>

#define __CUDA_ARCH__ 860
#include "hair_msnn.cuh"
#undef __CUDA_ARCH__
#include "utils.cuh"
#include "curve_utils.cuh"
#include "disney_hair.cuh"
#include "frostbite_anisotropic.cuh"
#include "ref_draw_order.h"
#define lcg_randomf(r) refemu::ordered_draw((r), __FILE__, __LINE__)
#include "optix_common.cuh"
#include "hair_msnn.cu"

#include "ref_exports.inc"

extern "C" {

int ref_msnn_gbuffer_stride() { return (int)sizeof(GBuffer); }

// G_BUFFER pass (cuda/hair_msnn.cu:187-312) for rows [y0,y1) of a W x H frame.
//   beta        : internal beta (CLI BETA - 1)
//   train_idxs  : int[records], every_nth as the host driver computes them
//   nn_frame_in : float[W*H*in_ch], nn_train_in: float[records*in_ch], nn_train_out: float[records*3]
//   gbuf_out    : float[W*H*8] = hit, isSurface, p[3], shortPathColor[3]
void ref_render_msnn_gbuffer(int accum_id, int y0, int y1, int W, int H, int beta, int every_nth, const int* train_idxs,
                             int in_ch, float* nn_frame_in, float* nn_train_in, float* nn_train_out, float* gbuf_out,
                             int threads) {
    LaunchParams& P = optixLaunchParams;
    P.accumId = accum_id;
    P.pass = G_BUFFER;
    P.beta = beta;
    P.everyNth = every_nth;
    P.trainIdxs = (int*)train_idxs;
    P.mlpInputCh = in_ch; P.mlpOutputCh = 3;
    P.nnFrameInput = nn_frame_in;
    P.nnTrainInput = nn_train_in; P.nnTrainOutput = nn_train_out;
    std::vector<GBuffer> gb((size_t)W * H);
    P.gBuffer = gb.data();
    g_raygen_data.frameBuffer = nullptr;
    g_raygen_data.frameBufferSize = vec2i(W, H);
    if (threads < 1) threads = 1;
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) {
        pool.emplace_back([=]() {
            for (int y = y0 + t; y < y1; y += threads)
                for (int x = 0; x < W; ++x) {
                    refemu::g_ctx.launch_x = x; refemu::g_ctx.launch_y = y;
                    refemu::g_ctx.program_data = &g_raygen_data;
                    ref_raygen_rayGenCam();
                }
        });
    }
    for (auto& th : pool) th.join();
    for (int y = y0; y < y1; ++y)
        for (int x = 0; x < W; ++x) {
            size_t i = (size_t)y * W + x;
            float* o = gbuf_out + 8 * i;
            o[0] = gb[i].hit; o[1] = gb[i].isSurface;
            o[2] = gb[i].p.x; o[3] = gb[i].p.y; o[4] = gb[i].p.z;
            o[5] = gb[i].shortPathColor.x; o[6] = gb[i].shortPathColor.y; o[7] = gb[i].shortPathColor.z;
        }
}

// Same pass for a LIST of rows (bench.py's CPU arm: rows stratified over the frame).  Pixels are handed to the
// threads in chunks of 32 through an atomic cursor (path lengths vary a lot between hair and background);
// the G-buffer is kept between calls.  gbuf_out: float[n_rows*W*8], compact in list order.
void ref_render_msnn_gbuffer_rows(int accum_id, const int* rows, int n_rows, int W, int H, int beta, int every_nth, const int* train_idxs,
                                  int in_ch, float* nn_frame_in, float* nn_train_in, float* nn_train_out, float* gbuf_out, int threads) {
    LaunchParams& P = optixLaunchParams;
    P.accumId = accum_id;
    P.pass = G_BUFFER;
    P.beta = beta;
    P.everyNth = every_nth;
    P.trainIdxs = (int*)train_idxs;
    P.mlpInputCh = in_ch; P.mlpOutputCh = 3;
    P.nnFrameInput = nn_frame_in;
    P.nnTrainInput = nn_train_in; P.nnTrainOutput = nn_train_out;
    static std::vector<GBuffer> gb;
    if (gb.size() != (size_t)W * H) gb.assign((size_t)W * H, GBuffer());
    P.gBuffer = gb.data();
    g_raygen_data.frameBuffer = nullptr;
    g_raygen_data.frameBufferSize = vec2i(W, H);
    if (threads < 1) threads = 1;
    const int total = n_rows * W;
    std::atomic<int> cursor{0};
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) {
        pool.emplace_back([&]() {
            for (;;) {
                const int k0 = cursor.fetch_add(32);
                if (k0 >= total) break;
                for (int k = k0; k < k0 + 32 && k < total; ++k) {
                    refemu::g_ctx.launch_x = k % W; refemu::g_ctx.launch_y = rows[k / W];
                    refemu::g_ctx.program_data = &g_raygen_data;
                    ref_raygen_rayGenCam();
                }
            }
        });
    }
    for (auto& th : pool) th.join();
    if (gbuf_out)
        for (int k = 0; k < total; ++k) {
            const size_t i = (size_t)rows[k / W] * W + (k % W);
            float* o = gbuf_out + 8 * (size_t)k;
            o[0] = gb[i].hit; o[1] = gb[i].isSurface;
            o[2] = gb[i].p.x; o[3] = gb[i].p.y; o[4] = gb[i].p.z;
            o[5] = gb[i].shortPathColor.x; o[6] = gb[i].shortPathColor.y; o[7] = gb[i].shortPathColor.z;
        }
}

// TRAIN_DATA_GEN pass (cuda/hair_msnn.cu:222-233 + the shared training-path block :234-277): a
// 128 x 128 launch inside a W x H frame; record i's ray runs from the camera position to
// sampled_points[scene_indices[i]].  nn_train_in: float[16384*in_ch], nn_train_out: float[16384*3].
void ref_render_msnn_train_data_gen(int accum_id, int W, int H, int beta, const int* scene_indices, const float* sampled_points3,
                                    int in_ch, float* nn_train_in, float* nn_train_out, int threads) {
    LaunchParams& P = optixLaunchParams;
    P.accumId = accum_id;
    P.pass = TRAIN_DATA_GEN;
    P.beta = beta;
    P.numTrainRecordsX = 128; P.numTrainRecordsY = 128;
    P.sceneIndices = (int*)scene_indices;
    P.sampledPoints = (float3*)sampled_points3;
    P.mlpInputCh = in_ch; P.mlpOutputCh = 3;
    P.nnTrainInput = nn_train_in; P.nnTrainOutput = nn_train_out;
    g_raygen_data.frameBuffer = nullptr;
    g_raygen_data.frameBufferSize = vec2i(W, H);
    if (threads < 1) threads = 1;
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) {
        pool.emplace_back([=]() {
            for (int y = t; y < 128; y += threads)
                for (int x = 0; x < 128; ++x) {
                    refemu::g_ctx.launch_x = x; refemu::g_ctx.launch_y = y;
                    refemu::g_ctx.program_data = &g_raygen_data;
                    ref_raygen_rayGenCam();
                }
        });
    }
    for (auto& th : pool) th.join();
}

// RENDER pass (cuda/hair_msnn.cu:314-356) over all pixels; buffers are float4[W*H].
void ref_render_msnn_composite(int accum_id, int W, int H, const float* gbuf8, const float* nn_out3,
                               float* pt_accum, float* nn_accum, float* final_accum,
                               float* pt_avg, float* nn_avg, float* final_avg, uint32_t* fb) {
    LaunchParams& P = optixLaunchParams;
    P.accumId = accum_id;
    P.pass = RENDER;
    P.mlpOutputCh = 3;
    std::vector<GBuffer> gb((size_t)W * H);
    for (size_t i = 0; i < gb.size(); ++i) {
        gb[i].hit = gbuf8[8 * i] != 0.f; gb[i].isSurface = gbuf8[8 * i + 1] != 0.f;
        gb[i].p = vec3f(gbuf8[8 * i + 2], gbuf8[8 * i + 3], gbuf8[8 * i + 4]);
        gb[i].shortPathColor = vec3f(gbuf8[8 * i + 5], gbuf8[8 * i + 6], gbuf8[8 * i + 7]);
    }
    P.gBuffer = gb.data();
    P.nnFrameOutput = (float*)nn_out3;
    P.ptAccumBuffer = (float4*)pt_accum; P.nnAccumBuffer = (float4*)nn_accum; P.finalAccumBuffer = (float4*)final_accum;
    P.ptAverageBuffer = (float4*)pt_avg; P.nnAverageBuffer = (float4*)nn_avg; P.finalAverageBuffer = (float4*)final_avg;
    g_raygen_data.frameBuffer = fb;
    g_raygen_data.frameBufferSize = vec2i(W, H);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            refemu::g_ctx.launch_x = x; refemu::g_ctx.launch_y = y;
            refemu::g_ctx.program_data = &g_raygen_data;
            ref_raygen_rayGenCam();
        }
}

}  // extern "C"
