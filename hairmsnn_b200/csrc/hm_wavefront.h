// hm_wavefront.h — device-side state of the wavefront path tracer and the launch
// entry points implemented in hm_wavefront.cu.
//
// One "slot" per pixel of the rank's tile rows (slot == pixel index in the FULL frame,
// so RNG streams and training-pixel selection are independent of the partition,
// SURVEY §8e).  Path state is SoA in HBM (float4 / u32 arrays, 128-bit accesses);
// queues hold slot ids (extend, shade) or packed occlusion rays (shadow) and are
// filled with warp-aggregated atomics.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "hm_shade.h"

namespace hm {

enum PathMode { MODE_PT = 0, MODE_MSNN = 1, MODE_NRC = 2 };
enum RendererKind { HM_KIND_PT = 0, HM_KIND_NRC = 1, HM_KIND_MSNN = 2 };

struct Camera {
    float pos[3], d00[3], du[3], dv[3];
};

struct PathBuffers {
    uint32_t* rng;
    float4* ray_o;
    float4* ray_d;
    float4* hit;         // t, prim (int bits), u, v
    float4* beta;        // rgb, w = bounce count (int bits)
    float4* color;       // rgb radiance so far
    float4* dl_beta;     // throughput multiplying the pending direct sample
    float4* dl_light;    // pending light-probe value, w = has-pending flag
    float4* dl_bsdf;     // pending bsdf-probe value
    uint32_t* vis;       // bit0 light probe unoccluded, bit1 bsdf probe unoccluded
    // MSNN training paths: short-path companion
    float4* beta_short;
    float4* color_short;
    float4* dl_beta_short;
    // NRC: spread heuristic state (cuda/nrc.cu:135-310)
    float4* nrc_state;   // spread, a0, c, flags (int bits: kNrc*)
    float4* nrc_prev;    // previous vertex position (before the spawn offset), w = its sampling pdf
};

// render_nrc constants (cuda_headers/nrc.cuh:13, headers/render_nrc.h:126,132)
constexpr int kNrcMaxBounces = 40;
constexpr int kNrcTrainRecords = 65536;
enum { kNrcSuffix = 1, kNrcRecalcA0 = 2, kNrcTerminated = 4 };

// One training pixel's path record (TrainBuffer, cuda_headers/nrc.cuh:29-40, without the
// `color` member nothing reads).
struct NrcTrainRec {
    float vert[kNrcMaxBounces][3];
    float wo[kNrcMaxBounces][3];
    float n[kNrcMaxBounces][3];
    float radiance[kNrcMaxBounces][3];
    float beta[kNrcMaxBounces][3];
    int bounces;
    int hit;
};

struct Queues {
    int* shade[2];       // double-buffered slot lists
    int* extend;
    float4* shadow;      // 2 x float4 per entry: (o.xyz, slot|bit<<31) (d.xyz, -)
    int* counts;         // [0],[1] shade ; [2] extend ; [3] shadow ; work cursors: [4] trace [6] primary [7] tail
    unsigned long long* trav;   // [0],[1] extend nodes/prims ; [2],[3] shadow nodes/prims ; [4],[5] primary ;
                                // [6] extend rays ; [7] shadow rays ; [8] shade items ; [9] primary rays ;
                                // tail-piece share of the above: [10] nodes [11] prims [12] rays
};

// Everything a frame's kernels need, passed by value (fits the 4 KB param space).
struct FrameParams {
    SceneView scene;
    Camera cam;
    PathBuffers paths;
    Queues q;
    int W, H;            // full frame
    int row0, row1;      // this rank's rows [row0,row1)
    int accum_id;        // accumulation index of this renderer (0 = first sample in its buffers)
    int frame_id;        // sample index that keys the RNG streams (== accum_id unless spp-sharded)
    int collect_stats;   // accumulate nodes/prims visited into q.trav
    int tail;            // this launch belongs to the frame's tail piece (long paths only): separate counters
    int tail_merged;     // tail piece of SEVERAL HairMSNN frames in one launch sequence (launch_merge_tail): slots run over the
                         // frames' consecutive path-state arrays; every live path is a training path (no per-frame lookups)
    int n_primary;       // primary work items (rows * W; numTrainRecords in the TRAIN_DATA_GEN pass)
    // HairMSNN TRAIN_DATA_GEN pass (cuda/hair_msnn.cu:222-233): work item i is training record i, its
    // ray runs from the camera position to sampled_points[scene_indices[i]]
    int pretrain;
    const float* sampled_points;   // [numSamples][3]
    const int* scene_indices;      // [numSamples] shuffled
    int v1_stop, v2_stop;
    int mode;
    // PT outputs
    float4* accum; float4* average; uint32_t* fb;
    // MSNN
    int msnn_beta;       // internal beta = CLI BETA - 1
    int every_nth;
    const int* train_idxs;
    int train_slot0, train_slots;    // this rank's range of training records
    int train_records;   // length of train_idxs (numTrainRecords); groups past it have no training pixel
    float* nn_frame_in;  // [W*H][12]
    float* nn_train_in;  // [records][12]
    float* nn_train_out; // [records][3]
    float4* gbuffer;     // rgb = short-path colour (NRC: pathRadiance), w = flags (bit0 hit, bit1 surface)
    int* query_tiles;    // [W*H / 128] set to 1 when any pixel of the 128-pixel tile reads its cache output
    int in_ch;
    // NRC
    float4* gbuffer_b;   // rgb = throughput at the cache query (GBuffer::beta), w = bounces (int bits)
    NrcTrainRec* tbuffer;        // [nrc_train_pixels]
    int nrc_train_pixels;
    int nrc_all_unbiased;
    float nrc_c;
};

// RENDER pass of the NRC program (cuda/nrc.cu:69-133,367-381)
struct NrcRender {
    float4* accum; float4* average; uint32_t* fb;
    const float4* gbuffer; const float4* gbuffer_b;
    const NrcTrainRec* tbuffer;
    const int* train_idxs;
    const float* nn_out;   // [nn_frame_size][3]
    float* train_in;       // [records][in_ch]
    float* train_gt;       // [records][3]
    int W, H, in_ch, every_nth, train_pixels, all_unbiased, accum_id;
};

struct MsnnComposite {
    float4* pt_accum; float4* nn_accum; float4* final_accum;
    float4* pt_avg; float4* nn_avg; float4* final_avg;
    uint32_t* fb;
    const float4* gbuffer;
    const float* nn_out;  // [W*H][3]
    int accum_id;
    int first, count;     // pixel range
};

// All launches are asynchronous on `stream`.
void launch_primary(const FrameParams& P, cudaStream_t stream);
void launch_shade(const FrameParams& P, int src_queue, cudaStream_t stream, long long max_items = 0);
// occlusion probes + continuation rays of one vertex in one launch
void launch_trace(const FrameParams& P, int dst_queue, cudaStream_t stream, long long max_items = 0);
// render_hair_msnn, tail pieces of n consecutive frames as one: the survivors of the frames' main pieces (shade queue
// `src` of each) are concatenated into the group's queue `out`, frame k's slots moved up by k * stride — the frames'
// path-state arrays lie `stride` elements apart, so the tail kernels address all of them from frame 0's pointers.
constexpr int kTailGroupMax = 8;
struct TailMerge {
    const int* counts[kTailGroupMax];
    const int* queue[kTailGroupMax];
    int n, src, stride, cap;
    int* out;
    int* out_counts;
};
void launch_merge_tail(const TailMerge& M, int max_items_per_frame, cudaStream_t stream);
void launch_finalize(const FrameParams& P, cudaStream_t stream);
void launch_msnn_composite(const MsnnComposite& C, cudaStream_t stream);
void launch_nrc_render(const NrcRender& R, cudaStream_t stream);
// multi-GPU: summed accumulation buffer -> average (in place) + optional 8-bit frame
void launch_resolve_sum(float4* img, uint32_t* fb, float inv_total, int n, cudaStream_t stream);
// test hook: closest-hit / any-hit for caller-supplied rays (device pointers)
void launch_trace_rays(const SceneView& S, const float* org, const float* dir, int n, int any,
                       float tmin, float tmax, float4* out_hit, int* out_stats, int* cursor, cudaStream_t stream);
// generateEnvSamplingTables (scene.cpp:349-425) on the device: env RGBA32F [H][W] -> cPdf/cCdf [H][W+1], mPdf/mCdf [H+1]
void launch_env_tables(const float* env_rgba, const float* sin_theta, int W, int H, float* cpdf, float* ccdf, float* mpdf, float* mcdf,
                       cudaStream_t stream);
// every 64th entry of each conditional-cdf row -> coarse [H][W / 64]
void launch_env_coarse(const float* ccdf, int W, int H, float* coarse, cudaStream_t stream);
// test hooks: hair_eval / hair_sample_dir + hair_eval (hm_bsdf.h) for caller-supplied local directions (device pointers)
void launch_bsdf_eval(const HairLobes& L, const float* wo, const float* wi, const float* h, int n, float* out_f, float* out_pdf, cudaStream_t stream);
void launch_bsdf_sample(const HairLobes& L, const float* wo, const float* h, const float* u, int n, float* out_wi, float* out_f, float* out_pdf,
                        cudaStream_t stream);
int wavefront_sm_count();
uint64_t wavefront_launch_count();

}  // namespace hm
