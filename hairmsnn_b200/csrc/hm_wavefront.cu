// hm_wavefront.cu — sm_100a kernels of the wavefront path tracer.
//
// Replaces the reference's OptiX megakernels rayGenCam (cuda/path_tracing.cu:19-67,
// cuda/hair_msnn.cu:187-357) and the hit/miss programs they invoke
// (cuda_headers/optix_common.cuh:485-616).  Stages per frame:
//
//   primary  : pixel -> RNG stream -> camera ray -> closest hit      (coherent rays)
//   shade    : hit -> vertex; direct-light probes (<=2 occlusion rays) + Russian
//              roulette + BSDF continuation ray                      (ALU bound)
//   shadow   : any-hit traversal of the occlusion queue              (L2/HBM latency)
//   extend   : closest-hit traversal of the continuation queue       (L2/HBM latency)
//   finalize : fold the last pending probes, write pixel / G-buffer  (HBM streaming)
//
// shade/shadow/extend are persistent grids (148 SMs x resident CTAs) that pull their
// item count from device memory, so the host never synchronises inside a bounce.
#include "hm_wavefront_dev.cuh"
#include "hm_trace_dev.cuh"

#include <atomic>

namespace hm {

static std::atomic<uint64_t> g_launches{0};
uint64_t wavefront_launch_count() { return g_launches.load(); }
void wavefront_count_launch() { g_launches++; }

int wavefront_sm_count() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

namespace {

#ifndef HM_TRACE_CTAS
#define HM_TRACE_CTAS 6
#endif
constexpr int kTraceCtasPerSm = HM_TRACE_CTAS;   // persistent traversal grid: resident CTAs per SM (launch bounds cap the registers)

__device__ __forceinline__ void flush_trav(unsigned long long* dst, const TraceStats& st) {
    unsigned n = (unsigned)st.nodes, p = (unsigned)st.prims;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { n += __shfl_xor_sync(0xffffffffu, n, o); p += __shfl_xor_sync(0xffffffffu, p, o); }
    if ((threadIdx.x & 31) == 0 && (n | p)) { atomicAdd(dst, (unsigned long long)n); atomicAdd(dst + 1, (unsigned long long)p); }
}

// ---------------------------------------------------------------------------------
// primary
// ---------------------------------------------------------------------------------
// measured on the real scenes/curly frame (profiles/r2d_sweep_knobs.txt): scanlines 1.27 ms, 8 x 4 tiles 1.35 ms
#ifndef HM_PRIMARY_TILES
#define HM_PRIMARY_TILES 0
#endif
struct PrimaryOps {
    const FrameParams& P;
    int first;
    bool tiled;
    // Work item -> pixel slot.  With HM_PRIMARY_TILES the 32 consecutive work items a warp pulls are an 8 x 4 pixel
    // tile instead of 32 pixels of a scanline: camera rays of a tile stay together deeper into the tree.  The slot
    // (= full-frame pixel index) keys the RNG and all path state, so results do not depend on the mapping.
    __device__ __forceinline__ int slot_of(int w) const {
        if (!tiled) return first + w;
        const int t = w >> 5, i = w & 31, tiles_x = P.W >> 3;
        const int tx = t % tiles_x, ty = t / tiles_x;
        return first + (ty * 4 + (i >> 3)) * P.W + tx * 8 + (i & 7);
    }
    __device__ __forceinline__ bool fetch(int w, V3& o, V3& d) const {
        const int slot = slot_of(w);
        o = V3(P.cam.pos[0], P.cam.pos[1], P.cam.pos[2]);
        Rng rng;
        if (P.pretrain) {
            // launch index = (w % numTrainRecordsX, w / numTrainRecordsX) in a frame W wide; no pixel jitter is drawn
            rng = rng_seed(P.frame_id + 10007, (uint32_t)(w % 128), (uint32_t)(w / 128), (uint32_t)P.W);
            const float* sp = P.sampled_points + 3 * (size_t)__ldg(P.scene_indices + w);
            d = normalize(V3(__ldg(sp), __ldg(sp + 1), __ldg(sp + 2)) - o);
        } else {
            const int px = slot % P.W, py = slot / P.W;
            rng = rng_seed(P.frame_id + 10007, (uint32_t)px, (uint32_t)py, (uint32_t)P.W);
            float ox = rng_next(rng);
            float oy = rng_next(rng);
            float su = ((float)px + ox) / (float)P.W;
            float sv = ((float)py + oy) / (float)P.H;
            d = normalize(V3(P.cam.d00[0], P.cam.d00[1], P.cam.d00[2]) +
                          su * V3(P.cam.du[0], P.cam.du[1], P.cam.du[2]) +
                          sv * V3(P.cam.dv[0], P.cam.dv[1], P.cam.dv[2]));
        }
        P.paths.rng[slot] = rng.state;
        P.paths.ray_o[slot] = f4(o, 0.f);
        P.paths.ray_d[slot] = f4(d, 0.f);
        return false;
    }
    __device__ __forceinline__ void commit(int w, const Hit& h, bool finished) const {
        const int slot = finished ? slot_of(w) : first;
        const bool hit_any = finished && h.prim >= 0;
        if (finished) {
            P.paths.hit[slot] = make_float4(h.t, __int_as_float(h.prim), h.u, h.v);
            P.paths.beta[slot] = make_float4(1.f, 1.f, 1.f, __int_as_float(0));
            P.paths.dl_light[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
            const V3 d = v3(P.paths.ray_d[slot]);
            V3 c(0.f);
            if (!hit_any && P.scene.lights.env.has_env) c = env_radiance(P.scene.lights.env, d);
            P.paths.color[slot] = f4(c, 0.f);
            if (P.mode == MODE_MSNN) {
                P.paths.beta_short[slot] = make_float4(1.f, 1.f, 1.f, 0.f);
                P.paths.color_short[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (!hit_any) {
                    // Interaction defaults (common.cuh:46-68): p = 0, t = 0; wo = -dir
                    write_nn_input(P.nn_frame_in + (size_t)slot * P.in_ch, V3(0.f), -1.f * d, V3(0.f), P.scene.scene_scale);
                    int tr;
                    if (is_training_pixel(P, slot, tr) && tr >= P.train_slot0 && tr < P.train_slot0 + P.train_slots) {
                        write_nn_input(P.nn_train_in + (size_t)tr * P.in_ch, V3(0.f), -1.f * d, V3(0.f), P.scene.scene_scale);
                        float* o3 = P.nn_train_out + (size_t)tr * 3;
                        o3[0] = 0.f; o3[1] = 0.f; o3[2] = 0.f;
                    }
                }
            }
        }
        int idx = queue_reserve(P.q.counts + 0, hit_any);
        if (idx >= 0) P.q.shade[0][idx] = slot;
    }
};

__global__ void __launch_bounds__(kBlock, kTraceCtasPerSm) k_primary(const __grid_constant__ FrameParams P) {
    const int n = P.n_primary;
    TraceStats st[2] = {{0, 0}, {0, 0}};
    const bool tiled = HM_PRIMARY_TILES && !P.pretrain && (P.W & 7) == 0 && ((P.row1 - P.row0) & 3) == 0;
    PrimaryOps ops{P, P.pretrain ? 0 : P.row0 * P.W, tiled};
    trace_queue(P.scene.geom, n, P.q.counts + 6, ops, 0.f, 1e30f, P.collect_stats ? st : nullptr);
    if (P.collect_stats) {
        flush_trav(P.q.trav + 4, st[0]);
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(P.q.trav + 9, (unsigned long long)n);
    }
}

// ---------------------------------------------------------------------------------
// extend / shadow
// ---------------------------------------------------------------------------------
// One launch traces a vertex's occlusion probes AND its continuation rays: work items
// [0, n_shadow) are shadow-queue entries (any-hit), [n_shadow, n_shadow + n_extend) extend-queue
// entries (closest-hit).  Fewer, fuller launches: the long-path tail is a chain of
// latency-bound launches, and the two ray kinds share warps.
struct TraceOps {
    const FrameParams& P;
    int n_shadow, dst;
    __device__ __forceinline__ bool fetch(int w, V3& o, V3& d) const {
        if (w < n_shadow) {
            float4 a = P.q.shadow[2 * (size_t)w + 0];
            float4 b = P.q.shadow[2 * (size_t)w + 1];
            o = V3(a.x, a.y, a.z);
            d = V3(b.x, b.y, b.z);
            return true;
        }
        const int slot = P.q.extend[w - n_shadow];
        o = v3(P.paths.ray_o[slot]);
        d = v3(P.paths.ray_d[slot]);
        return false;
    }
    __device__ __forceinline__ void commit(int w, const Hit& h, bool finished) const {
        const bool got = finished && h.prim >= 0;
        bool to_shade = false;
        int slot = 0;
        if (got) {
            if (w < n_shadow) {
                const int tag = __float_as_int(P.q.shadow[2 * (size_t)w].w);
                atomicAnd(P.paths.vis + (tag & 0x3fffffff), ~(1u << (tag >> 30)));
            } else {
                slot = P.q.extend[w - n_shadow];
                P.paths.hit[slot] = make_float4(h.t, __int_as_float(h.prim), h.u, h.v);
                to_shade = true;
            }
        }
        int idx = queue_reserve(P.q.counts + dst, to_shade);
        if (idx >= 0) P.q.shade[dst][idx] = slot;
    }
};

__global__ void __launch_bounds__(kBlock, kTraceCtasPerSm) k_trace(const __grid_constant__ FrameParams P, int dst) {
    const int n_extend = P.q.counts[2], n_shadow = P.q.counts[3];
    TraceStats st[2] = {{0, 0}, {0, 0}};
    TraceOps ops{P, n_shadow, dst};
    trace_queue(P.scene.geom, n_shadow + n_extend, P.q.counts + 4, ops, 0.f, 1e30f, P.collect_stats ? st : nullptr);
    if (P.collect_stats) {
        flush_trav(P.q.trav + 0, st[0]);
        flush_trav(P.q.trav + 2, st[1]);
        if (P.tail) {
            TraceStats both{st[0].nodes + st[1].nodes, st[0].prims + st[1].prims};
            flush_trav(P.q.trav + 10, both);
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            atomicAdd(P.q.trav + 6, (unsigned long long)n_extend);
            atomicAdd(P.q.trav + 7, (unsigned long long)n_shadow);
            if (P.tail) atomicAdd(P.q.trav + 12, (unsigned long long)(n_extend + n_shadow));
        }
    }
}

// ---------------------------------------------------------------------------------
// finalize
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_finalize(const __grid_constant__ FrameParams P) {
    const int n = P.n_primary;
    const int first = P.pretrain ? 0 : P.row0 * P.W;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int slot = first + i;
        V3 color = v3(P.paths.color[slot]);
        V3 color_short(0.f);
        int tr_ofs = 0;
        bool training = false;
        if (P.mode == MODE_MSNN) {
            training = is_training_pixel(P, slot, tr_ofs);
            if (training) color_short = v3(P.paths.color_short[slot]);
        }
        fold_pending(P, slot, training, color, color_short);

        if (P.mode == MODE_PT) {
            // writePixel (cuda_headers/utils.cuh:13-35)
            if (any_nan(color)) color = V3(0.f);
            if (P.accum_id > 0) color = color + v3(P.accum[slot]);
            P.accum[slot] = f4(color, 1.f);
            color = (1.f / (P.accum_id + 1)) * color;
            P.average[slot] = f4(color, 1.f);
            P.fb[slot] = pack_rgba8(V3(linear_to_srgb(color.x), linear_to_srgb(color.y), linear_to_srgb(color.z)));
        } else {
            float4 hr = P.paths.hit[slot];
            bool hit = __float_as_int(hr.y) >= 0;   // only a primary miss leaves prim < 0
            int flags = hit ? __float_as_int(P.gbuffer[slot].w) : 0;
            if (training) {
                if (hit) {
                    if (any_nan(color)) color = V3(0.f);
                    if (any_inf(color)) color = V3(1e5f);
                    if (any_nan(color_short)) color_short = V3(0.01f);
                    if (any_inf(color_short)) color_short = V3(1e5f);
                    if (tr_ofs >= P.train_slot0 && tr_ofs < P.train_slot0 + P.train_slots) {
                        float* o3 = P.nn_train_out + (size_t)tr_ofs * 3;
                        o3[0] = color.x - color_short.x;
                        o3[1] = color.y - color_short.y;
                        o3[2] = color.z - color_short.z;
                    }
                }
            }
            P.gbuffer[slot] = f4(color, __int_as_float(flags));
            // the RENDER pass reads the network output of hair hits only (cuda/hair_msnn.cu:325-340)
            if (P.query_tiles && (flags & 1) && !(flags & 2)) P.query_tiles[slot >> 7] = 1;
        }
    }
}

__global__ void __launch_bounds__(256) k_finalize_nrc(const __grid_constant__ FrameParams P) {
    const int n = (P.row1 - P.row0) * P.W;
    const int first = P.row0 * P.W;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int slot = first + i;
        V3 color = v3(P.paths.color[slot]);
        const bool hit = __float_as_int(P.paths.hit[slot].y) >= 0;   // only a primary miss leaves prim < 0
        if (!hit) {
            P.gbuffer[slot] = f4(color, __int_as_float(0));
            P.gbuffer_b[slot] = make_float4(0.f, 0.f, 0.f, __int_as_float(0));
            continue;
        }
        const int flags = __float_as_int(P.paths.nrc_state[slot].w);
        if (flags & kNrcTerminated) continue;
        const int b = __float_as_int(P.paths.beta[slot].w) - 1;     // last shaded vertex
        int tr_ofs = 0;
        bool unbiased = false;
        const bool training = nrc_training_pixel(P, slot, tr_ofs, unbiased);
        float4 dl = P.paths.dl_light[slot];
        if (dl.w != 0.f) {
            uint32_t vis = P.paths.vis[slot];
            V3 d = resolve_direct(v3(dl), (vis & 1u) != 0, v3(P.paths.dl_bsdf[slot]), (vis & 2u) != 0);
            color += v3(P.paths.dl_beta[slot]) * d;
            if (training) write3(P.tbuffer[tr_ofs].radiance[b], d);
        }
        if (!(flags & kNrcSuffix)) {
            P.gbuffer[slot] = f4(color, __int_as_float(1));
            P.gbuffer_b[slot] = make_float4(0.f, 0.f, 0.f, __int_as_float(b));
        }
        if (training) {
            // the reference lets every pixel of the group race on these two fields
            // (cuda/nrc.cu:210-211); only the training pixel's own values are meaningful
            P.tbuffer[tr_ofs].bounces = b;
            P.tbuffer[tr_ofs].hit = 0;
        }
    }
}

// RENDER pass: nrcGenerateTrainingData (cuda/nrc.cu:69-133) + composite (cuda/nrc.cu:367-381).
__global__ void __launch_bounds__(256) k_nrc_render(const NrcRender R) {
    const int n = R.W * R.H;
    for (int px = blockIdx.x * blockDim.x + threadIdx.x; px < n; px += gridDim.x * blockDim.x) {
        const float4 g = R.gbuffer[px];
        const float4 gb = R.gbuffer_b[px];
        const bool hit = (__float_as_int(g.w) & 1) != 0;
        const int tr_ofs = px / R.every_nth;
        bool training = false, unbiased = false;
        if (tr_ofs < R.train_pixels) {
            training = px % R.every_nth == __ldg(R.train_idxs + tr_ofs) % R.every_nth;
            unbiased = training && (tr_ofs % 16 == 0 || R.all_unbiased);
        }
        if (training) {
            float* tin = R.train_in + (size_t)tr_ofs * R.in_ch * kNrcMaxBounces;
            float* tgt = R.train_gt + (size_t)tr_ofs * 3 * kNrcMaxBounces;
            if (hit) {
                const NrcTrainRec& t = R.tbuffer[tr_ofs];
                V3 cache(0.f);
                if (t.hit && !unbiased) {
                    const float* o = R.nn_out + 3 * ((size_t)n + tr_ofs);
                    cache = V3(o[0], o[1], o[2]);
                }
                const int nb = t.bounces;
                for (int bounce = 0; bounce < nb; ++bounce) {
                    V3 color(0.f), beta(1.f);
                    for (int sub = bounce; sub < nb; ++sub) {
                        color = color + beta * V3(t.radiance[sub][0], t.radiance[sub][1], t.radiance[sub][2]);
                        beta = beta * V3(t.beta[sub][0], t.beta[sub][1], t.beta[sub][2]);
                    }
                    float* d = tin + bounce * R.in_ch;
                    d[0] = t.vert[bounce][0]; d[1] = t.vert[bounce][1]; d[2] = t.vert[bounce][2];
                    d[3] = t.wo[bounce][0]; d[4] = t.wo[bounce][1]; d[5] = t.wo[bounce][2];
                    d[6] = t.n[bounce][0]; d[7] = t.n[bounce][1]; d[8] = t.n[bounce][2];
                    V3 tc = color + beta * cache;
                    if (any_nan(tc)) tc = V3(0.f);
                    write3(tgt + 3 * bounce, tc);
                }
            } else {
                for (int k = 0; k < kNrcMaxBounces * R.in_ch; ++k) tin[k] = 0.f;
                for (int k = 0; k < kNrcMaxBounces * 3; ++k) tgt[k] = 0.f;
            }
        }
        V3 cache(R.nn_out[3 * (size_t)px + 0], R.nn_out[3 * (size_t)px + 1], R.nn_out[3 * (size_t)px + 2]);
        if (any_nan(cache)) cache = V3(0.f);
        V3 color = V3(g.x, g.y, g.z) + V3(gb.x, gb.y, gb.z) * V3(fmaxf(cache.x, 0.f), fmaxf(cache.y, 0.f), fmaxf(cache.z, 0.f));
        // writePixel (cuda_headers/utils.cuh:13-35)
        if (any_nan(color)) color = V3(0.f);
        if (R.accum_id > 0) color = color + v3(R.accum[px]);
        R.accum[px] = f4(color, 1.f);
        color = (1.f / (R.accum_id + 1)) * color;
        R.average[px] = f4(color, 1.f);
        R.fb[px] = pack_rgba8(V3(linear_to_srgb(color.x), linear_to_srgb(color.y), linear_to_srgb(color.z)));
    }
}

// RENDER pass of the HairMSNN program (cuda/hair_msnn.cu:314-356).
__global__ void __launch_bounds__(256) k_msnn_composite(const MsnnComposite C) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < C.count; i += gridDim.x * blockDim.x) {
        int px = C.first + i;
        float4 g = C.gbuffer[px];
        int flags = __float_as_int(g.w);
        V3 sp(g.x, g.y, g.z);
        V3 nn(C.nn_out[3 * (size_t)px + 0], C.nn_out[3 * (size_t)px + 1], C.nn_out[3 * (size_t)px + 2]);
        V3 color;
        if (!(flags & 1) || (flags & 2)) { color = sp; nn = color; }
        else color = sp + nn;
        if (C.accum_id > 0) {
            sp = sp + v3(C.pt_accum[px]);
            nn = nn + v3(C.nn_accum[px]);
            color = color + v3(C.final_accum[px]);
        }
        C.pt_accum[px] = f4(sp, 1.f);
        C.nn_accum[px] = f4(nn, 1.f);
        C.final_accum[px] = f4(color, 1.f);
        float inv = 1.f / (C.accum_id + 1);
        sp = inv * sp; nn = inv * nn; color = inv * color;
        C.pt_avg[px] = f4(sp, 1.f);
        C.nn_avg[px] = f4(nn, 1.f);
        C.final_avg[px] = f4(color, 1.f);
        C.fb[px] = pack_rgba8(V3(linear_to_srgb(color.x), linear_to_srgb(color.y), linear_to_srgb(color.z)));
    }
}

// Multi-GPU output resolve: `img` holds the all-reduced SUM of the ranks' accumulation buffers;
// turns it into the average over `total` samples in place and, for the final image, the 8-bit sRGB frame
// (writePixel's tail, utils.cuh:27-34).
__global__ void __launch_bounds__(256) k_resolve_sum(float4* img, uint32_t* fb, float inv_total, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 a = img[i];
        V3 c = inv_total * V3(a.x, a.y, a.z);
        img[i] = f4(c, 1.f);
        if (fb) fb[i] = pack_rgba8(V3(linear_to_srgb(c.x), linear_to_srgb(c.y), linear_to_srgb(c.z)));
    }
}

struct HookOps {
    const float* org; const float* dir; float4* out_hit; int any;
    __device__ __forceinline__ bool fetch(int w, V3& o, V3& d) const {
        o = V3(org[3 * w], org[3 * w + 1], org[3 * w + 2]);
        d = V3(dir[3 * w], dir[3 * w + 1], dir[3 * w + 2]);
        return any != 0;
    }
    __device__ __forceinline__ void commit(int w, const Hit& h, bool finished) const {
        if (finished) out_hit[w] = make_float4(h.t, __int_as_float(h.prim), h.u, h.v);
    }
};

// Test hook.  With out_stats == nullptr the rays go through the production warp-cooperative
// traversal; with per-ray statistics requested, through the portable one-thread-per-ray loop
// (the counters are per ray there).
__global__ void __launch_bounds__(kBlock) k_trace_rays(const SceneView S, const float* org, const float* dir, int n,
                                                        int any, float tmin, float tmax, float4* out_hit, int* out_stats,
                                                        int* cursor) {
    if (!out_stats) {
        HookOps ops{org, dir, out_hit, any};
        trace_queue(S.geom, n, cursor, ops, tmin, tmax, nullptr);
        return;
    }
    for (int i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) {
        V3 o(org[3 * i], org[3 * i + 1], org[3 * i + 2]), d(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]);
        TraceStats st; st.nodes = 0; st.prims = 0;
        Hit h = any ? trace_wide<true>(S.geom, o, d, tmin, tmax, &st) : trace_wide<false>(S.geom, o, d, tmin, tmax, &st);
        out_hit[i] = make_float4(h.t, __int_as_float(h.prim), h.u, h.v);
        out_stats[2 * i] = st.nodes; out_stats[2 * i + 1] = st.prims;
    }
}

}  // namespace

int persistent_grid(int ctas_per_sm) { return wavefront_sm_count() * ctas_per_sm; }

void launch_primary(const FrameParams& P, cudaStream_t stream) {
    int n = P.n_primary;
    int blocks = (n + kBlock - 1) / kBlock;
    int grid = blocks < persistent_grid(kTraceCtasPerSm) ? blocks : persistent_grid(kTraceCtasPerSm);
    if (grid < 1) grid = 1;
    k_primary<<<grid, kBlock, 0, stream>>>(P);
    g_launches++;
}
void launch_trace(const FrameParams& P, int dst, cudaStream_t stream, long long max_items) {
    // a vertex pushes at most 3 rays (2 probes + 1 continuation); 2 x 32 rays per warp keeps the refill loop busy
    k_trace<<<bounded_grid(persistent_grid(kTraceCtasPerSm), 3 * max_items, 2 * kBlock), kBlock, 0, stream>>>(P, dst);
    g_launches++;
}
__global__ void __launch_bounds__(256) k_merge_tail(const TailMerge M) {
    const int k = blockIdx.y;
    int ofs = 0;
    for (int j = 0; j < k; ++j) ofs += M.counts[j][M.src];
    int nk = M.counts[k][M.src];
    if (ofs + nk > M.cap) nk = M.cap > ofs ? M.cap - ofs : 0;   // cannot happen: a frame has at most `records` training paths
    const int* q = M.queue[k];
    const int shift = k * M.stride;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nk; i += gridDim.x * blockDim.x) M.out[ofs + i] = q[i] + shift;
    if (k == M.n - 1 && blockIdx.x == 0 && threadIdx.x == 0) M.out_counts[M.src] = ofs + nk;
}
void launch_merge_tail(const TailMerge& M, int max_items_per_frame, cudaStream_t stream) {
    int gx = (max_items_per_frame + 255) / 256;
    if (gx < 1) gx = 1;
    k_merge_tail<<<dim3(gx, M.n), 256, 0, stream>>>(M);
    g_launches++;
}
void launch_finalize(const FrameParams& P, cudaStream_t stream) {
    if (P.mode == MODE_NRC) k_finalize_nrc<<<persistent_grid(8), 256, 0, stream>>>(P);
    else k_finalize<<<persistent_grid(8), 256, 0, stream>>>(P);
    g_launches++;
}
void launch_msnn_composite(const MsnnComposite& C, cudaStream_t stream) {
    k_msnn_composite<<<persistent_grid(8), 256, 0, stream>>>(C);
    g_launches++;
}
void launch_resolve_sum(float4* img, uint32_t* fb, float inv_total, int n, cudaStream_t stream) {
    k_resolve_sum<<<persistent_grid(8), 256, 0, stream>>>(img, fb, inv_total, n);
    g_launches++;
}
void launch_nrc_render(const NrcRender& R, cudaStream_t stream) {
    k_nrc_render<<<persistent_grid(8), 256, 0, stream>>>(R);
    g_launches++;
}
// Environment importance tables on the device (generateEnvSamplingTables, scene.cpp:349-425): one thread per row
// keeps the reference's sequential float accumulation (bit-identical to the host recipe; this translation unit is
// compiled with -fmad=false), rows are independent; the marginal is one short sequential pass.  sin_theta comes
// from the host so that libdevice's sinf cannot move an entry by an ulp.
__global__ void __launch_bounds__(64) k_env_table_rows(const float4* __restrict__ env, const float* __restrict__ sin_theta, int W, int H,
                                                        float* __restrict__ cpdf, float* __restrict__ ccdf) {
    const int y = blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= H) return;
    const int cw = W + 1;
    const float st = sin_theta[y];
    const float4* row = env + (size_t)y * W;
    float* pdf = cpdf + (size_t)y * cw;
    float* cdf = ccdf + (size_t)y * cw;
    auto avg = [&](int x) { const float4 p = __ldg(row + x); return (p.x + p.y + p.z) * (1.0f / 3.0f); };
    float prev_pdf = avg(0) * st, prev_cdf = 0.f;
    pdf[0] = prev_pdf; cdf[0] = 0.f;
    for (int x = 1; x < W; ++x) {
        const float p = avg(x) * st;
        const float c = prev_cdf + prev_pdf / W;
        pdf[x] = p; cdf[x] = c;
        prev_pdf = p; prev_cdf = c;
    }
    const float total = prev_cdf + prev_pdf / W;
    pdf[W] = total;
    if (total > 0.f) {
        const float inv = 1.0f / total;
        for (int x = 1; x < W; ++x) cdf[x] *= inv;
    }
    cdf[W] = 1.0f;
}
__global__ void k_env_table_marginal(const float* __restrict__ cpdf, int W, int H, float* __restrict__ mpdf, float* __restrict__ mcdf) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    const int cw = W + 1;
    mpdf[0] = cpdf[W];
    mcdf[0] = 0.f;
    for (int i = 1; i < H; ++i) {
        mpdf[i] = cpdf[(size_t)i * cw + W];
        mcdf[i] = mcdf[i - 1] + mpdf[i - 1] / H;
    }
    const float total = mcdf[H - 1] + mpdf[H - 1] / H;
    mpdf[H] = total;
    if (total > 0.f)
        for (int i = 1; i < H; ++i) mcdf[i] /= total;
    mcdf[H] = 1.0f;
}
// every 64th entry of each conditional-cdf row (first level of cdf_lower_bound_two_level, hm_light.h)
__global__ void __launch_bounds__(256) k_env_table_coarse(const float* __restrict__ ccdf, int W, int H, float* __restrict__ coarse) {
    const int K = W >> 6;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K * H) return;
    const int y = i / K, k = i % K;
    coarse[i] = ccdf[(size_t)y * (W + 1) + 64 * k];
}
void launch_env_coarse(const float* ccdf, int W, int H, float* coarse, cudaStream_t stream) {
    const int n = (W >> 6) * H;
    k_env_table_coarse<<<(n + 255) / 256, 256, 0, stream>>>(ccdf, W, H, coarse);
    g_launches++;
}
void launch_env_tables(const float* env_rgba, const float* sin_theta, int W, int H, float* cpdf, float* ccdf, float* mpdf, float* mcdf,
                       cudaStream_t stream) {
    k_env_table_rows<<<(H + 63) / 64, 64, 0, stream>>>((const float4*)env_rgba, sin_theta, W, H, cpdf, ccdf);
    k_env_table_marginal<<<1, 32, 0, stream>>>(cpdf, W, H, mpdf, mcdf);
    g_launches += 2;
}

void launch_trace_rays(const SceneView& S, const float* org, const float* dir, int n, int any, float tmin, float tmax,
                       float4* out_hit, int* out_stats, int* cursor, cudaStream_t stream) {
    int blocks = (n + kBlock - 1) / kBlock;
    int grid = blocks < persistent_grid(kTraceCtasPerSm) ? blocks : persistent_grid(kTraceCtasPerSm);
    if (grid < 1) grid = 1;
    cudaMemsetAsync(cursor, 0, sizeof(int), stream);
    k_trace_rays<<<grid, kBlock, 0, stream>>>(S, org, dir, n, any, tmin, tmax, out_hit, out_stats, cursor);
    g_launches++;
}

}  // namespace hm
