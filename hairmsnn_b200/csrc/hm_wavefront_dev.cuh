// hm_wavefront_dev.cuh — device helpers shared by the two translation units of the wavefront kernels:
//   hm_wavefront.cu     traversal, finalize, composite, table builds: IEEE division / square root and no FMA
//                       contraction — hit ids, t, u and the environment tables are compared BIT FOR BIT with the host build;
//   hm_shade_kernels.cu shading (k_shade, k_shade_nrc, the BSDF hooks): tolerance-tested against the reference's headers;
//                       IEEE by default, approximate division / square root with `make SHADE_APPROX=1`.
#pragma once
#include "hm_wavefront.h"

namespace hm {

void wavefront_count_launch();
int persistent_grid(int ctas_per_sm);

namespace {

constexpr int kBlock = 128;

__device__ __forceinline__ float4 f4(V3 v, float w) { return make_float4(v.x, v.y, v.z, w); }
__device__ __forceinline__ V3 v3(float4 a) { return V3(a.x, a.y, a.z); }

// Warp-aggregated append: one atomicAdd per warp, returns this lane's index (or -1).
__device__ __forceinline__ int queue_reserve(int* counter, bool want) {
    unsigned mask = __ballot_sync(0xffffffffu, want);
    if (mask == 0) return -1;
    int lane = threadIdx.x & 31;
    int leader = __ffs(mask) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    return want ? base + __popc(mask & ((1u << lane) - 1u)) : -1;
}

__device__ __forceinline__ bool is_training_pixel(const FrameParams& P, int fb_ofs, int& tr_ofs) {
    if (P.pretrain) { tr_ofs = fb_ofs; return true; }   // TRAIN_DATA_GEN: every work item is a training record
    tr_ofs = fb_ofs / P.every_nth;
    // W*H need not be a multiple of numTrainRecords (everyNth = floor(W*H / 16384)): the reference reads past
    // trainIdxs for the trailing groups (cuda/hair_msnn.cu:208); they have no training pixel here
    if (tr_ofs >= P.train_records) return false;
    int train_idx = __ldg(P.train_idxs + tr_ofs) % P.every_nth;
    return fb_ofs % P.every_nth == train_idx;
}

__device__ __forceinline__ void write_nn_input(float* dst, V3 p, V3 wo, V3 t, float scene_scale) {
    V3 point = p / scene_scale;
    float4* d4 = reinterpret_cast<float4*>(dst);
    d4[0] = make_float4(point.x, point.y, point.z, wo.x);
    d4[1] = make_float4(wo.y, wo.z, t.x, t.y);
    d4[2] = make_float4(t.z, 0.f, 0.f, 0.f);
}

__device__ __forceinline__ void fold_pending(const FrameParams& P, int slot, bool training, V3& color, V3& color_short) {
    float4 dl = P.paths.dl_light[slot];
    if (dl.w == 0.f) return;
    uint32_t vis = __ldcg(P.paths.vis + slot);   // cleared by whichever thread traced the probe: bypass L1
    V3 d = resolve_direct(v3(dl), (vis & 1u) != 0, v3(P.paths.dl_bsdf[slot]), (vis & 2u) != 0);
    color += v3(P.paths.dl_beta[slot]) * d;
    if (training) color_short += v3(P.paths.dl_beta_short[slot]) * d;
}

__device__ __forceinline__ bool nrc_training_pixel(const FrameParams& P, int fb_ofs, int& tr_ofs, bool& unbiased) {
    tr_ofs = fb_ofs / P.every_nth;
    unbiased = false;
    // the reference indexes one group past the end when W*H % everyNth != 0 (trOfs == numTrainingPixels,
    // SURVEY §8 a18): that group has no training pixel here
    if (tr_ofs >= P.nrc_train_pixels) return false;
    const int train_idx = __ldg(P.train_idxs + tr_ofs) % P.every_nth;
    const bool training = fb_ofs % P.every_nth == train_idx;
    unbiased = training && (tr_ofs % 16 == 0 || P.nrc_all_unbiased);
    return training;
}

__device__ __forceinline__ void write3(float* dst, V3 v) { dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; }

// max_items > 0: the caller knows an upper bound of the queue length (tail pieces carry the few long paths
// only) — the grid is sized for it instead of for the whole GPU, so these launches do not sweep every SM
// with CTAs that find no work while other frames' main pieces are running.
static int bounded_grid(int full, long long max_items, int items_per_cta) {
    if (max_items <= 0) return full;
    long long need = (max_items + items_per_cta - 1) / items_per_cta;
    if (need < 1) need = 1;
    return need < full ? (int)need : full;
}
}  // namespace

}  // namespace hm
