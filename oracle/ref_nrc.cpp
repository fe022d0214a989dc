// ref_nrc.cpp — TEST INFRASTRUCTURE (oracle/_ref).  Builds the reference's own NRC device
// program (cuda/nrc.cu + cuda_headers/*.cuh, included from /root/reference where they lie)
// for the host.  See ref_optix_emul.h.
//
// The reference's G_BUFFER pass has a data race: EVERY pixel of a training group writes
// tBuffer[trOfs].bounces/.hit when its path leaves the scene (cuda/nrc.cu:210-211), not only
// the group's training pixel.  The host build serialises it as "non-training pixels first,
// training pixels last" — one of the orders the GPU can produce, and the one in which the
// training pixel's own record survives (the evident intent).
#include "ref_optix_emul.h"

#define __CUDA_ARCH__ 860
#include "nrc.cuh"
#undef __CUDA_ARCH__
#include "utils.cuh"
#include "curve_utils.cuh"
#include "disney_hair.cuh"
#include "frostbite_anisotropic.cuh"
#include "ref_draw_order.h"
#define lcg_randomf(r) refemu::ordered_draw((r), __FILE__, __LINE__)
#include "optix_common.cuh"
#include "nrc.cu"

#define REF_NO_PATH_LIMITS 1
#include "ref_exports.inc"

static std::vector<TrainBuffer> g_tbuffer;
static std::vector<GBuffer> g_gbuffer;

static void run_pixels(int W, int H, int threads, int every_nth, const int* train_idxs, int want_training) {
    if (threads < 1) threads = 1;
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) {
        pool.emplace_back([=]() {
            for (int y = t; y < H; y += threads)
                for (int x = 0; x < W; ++x) {
                    if (want_training >= 0) {
                        int fb = x + W * y;
                        int is_tr = fb % every_nth == train_idxs[fb / every_nth] % every_nth;
                        if (is_tr != want_training) continue;
                    }
                    refemu::g_ctx.launch_x = x; refemu::g_ctx.launch_y = y;
                    refemu::g_ctx.program_data = &g_raygen_data;
                    ref_raygen_rayGenCam();
                }
        });
    }
    for (auto& th : pool) th.join();
}

extern "C" {

int ref_nrc_max_bounces() { return MAX_BOUNCES; }

// G_BUFFER pass (cuda/nrc.cu:312-366) over a W x H frame.
//   train_idxs  : int[num_train_pixels + 1] (the reference reads one element past the end when
//                 W*H is not a multiple of everyNth; the caller supplies that element)
//   nn_frame_in : float[nn_frame_size * in_ch], zero-initialised by the caller (owlBufferClear)
//   gbuf_out    : float[W*H*8] = hit, pathRadiance[3], beta[3], bounces
void ref_render_nrc_gbuffer(int accum_id, int W, int H, int every_nth, const int* train_idxs, int num_train_pixels,
                            float c, int all_unbiased, int in_ch, float* nn_frame_in, float* gbuf_out, int threads) {
    LaunchParams& P = optixLaunchParams;
    P.accumId = accum_id;
    P.pass = G_BUFFER;
    P.everyNth = every_nth;
    P.trainIdxs = (int*)train_idxs;
    P.mlpInputCh = in_ch; P.mlpOutputCh = 3;
    P.nnFrameInput = nn_frame_in;
    P.numTrainingPixels = num_train_pixels;
    P.numTrainingRecords = num_train_pixels * MAX_BOUNCES;
    P.c = c;
    P.allUnbiased = all_unbiased != 0;
    P.showCache = false; P.showBounces = false;
    g_tbuffer.assign((size_t)num_train_pixels + 1, TrainBuffer());   // RESET pass state
    g_gbuffer.assign((size_t)W * H, GBuffer());
    P.tBuffer = g_tbuffer.data();
    P.gBuffer = g_gbuffer.data();
    g_raygen_data.frameBuffer = nullptr;
    g_raygen_data.frameBufferSize = vec2i(W, H);
    run_pixels(W, H, threads, every_nth, train_idxs, 0);
    run_pixels(W, H, 1, every_nth, train_idxs, 1);
    for (size_t i = 0; i < (size_t)W * H; ++i) {
        float* o = gbuf_out + 8 * i;
        const GBuffer& g = g_gbuffer[i];
        o[0] = g.hit; o[1] = g.pathRadiance.x; o[2] = g.pathRadiance.y; o[3] = g.pathRadiance.z;
        o[4] = g.beta.x; o[5] = g.beta.y; o[6] = g.beta.z; o[7] = (float)g.bounces;
    }
}

// training-path record of one training pixel after the G_BUFFER pass:
// out = bounces, hit, then per bounce b < MAX_BOUNCES: vert[3] wo[3] n[3] vertRadiance[3] vertBeta[3]
void ref_nrc_train_record(int tr_ofs, float* out) {
    const TrainBuffer& t = g_tbuffer[tr_ofs];
    out[0] = (float)t.bounces; out[1] = t.hit;
    for (int b = 0; b < MAX_BOUNCES; ++b) {
        float* o = out + 2 + 15 * b;
        o[0] = t.vert[b].x; o[1] = t.vert[b].y; o[2] = t.vert[b].z;
        o[3] = t.wo[b].x; o[4] = t.wo[b].y; o[5] = t.wo[b].z;
        o[6] = t.n[b].x; o[7] = t.n[b].y; o[8] = t.n[b].z;
        o[9] = t.vertRadiance[b].x; o[10] = t.vertRadiance[b].y; o[11] = t.vertRadiance[b].z;
        o[12] = t.vertBeta[b].x; o[13] = t.vertBeta[b].y; o[14] = t.vertBeta[b].z;
    }
}

// RENDER pass (cuda/nrc.cu:367-381) over all pixels, on the G_BUFFER state left by the call
// above.  nn_out: float[nn_frame_size*3]; train_in float[records*in_ch], train_gt float[records*3]
// (zero-initialised by the caller); accum/average float4[W*H]; fb uint32[W*H].
void ref_render_nrc_render(int accum_id, int W, int H, int every_nth, const int* train_idxs, int in_ch,
                           const float* nn_out, float* train_in, float* train_gt, float* accum, float* average,
                           uint32_t* fb) {
    LaunchParams& P = optixLaunchParams;
    P.accumId = accum_id;
    P.pass = RENDER;
    P.everyNth = every_nth;
    P.trainIdxs = (int*)train_idxs;
    P.mlpInputCh = in_ch; P.mlpOutputCh = 3;
    P.nnFrameOutput = (float3*)nn_out;
    P.trainInput = train_in; P.trainGT = (float3*)train_gt;
    P.accumBuffer = (float4*)accum; P.averageBuffer = (float4*)average;
    P.tBuffer = g_tbuffer.data();
    P.gBuffer = g_gbuffer.data();
    g_raygen_data.frameBuffer = fb;
    g_raygen_data.frameBufferSize = vec2i(W, H);
    run_pixels(W, H, 1, every_nth, train_idxs, -1);
}

}  // extern "C"
