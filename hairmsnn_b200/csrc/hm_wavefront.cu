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
#include "hm_wavefront.h"
#include "hm_trace_dev.cuh"

#include <atomic>

namespace hm {

static std::atomic<uint64_t> g_launches{0};
uint64_t wavefront_launch_count() { return g_launches.load(); }

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

constexpr int kBlock = 128;
#ifndef HM_TRACE_CTAS
#define HM_TRACE_CTAS 6
#endif
constexpr int kTraceCtasPerSm = HM_TRACE_CTAS;   // persistent traversal grid: resident CTAs per SM (launch bounds cap the registers)

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

__device__ __forceinline__ void flush_trav(unsigned long long* dst, const TraceStats& st) {
    unsigned n = (unsigned)st.nodes, p = (unsigned)st.prims;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { n += __shfl_xor_sync(0xffffffffu, n, o); p += __shfl_xor_sync(0xffffffffu, p, o); }
    if ((threadIdx.x & 31) == 0 && (n | p)) { atomicAdd(dst, (unsigned long long)n); atomicAdd(dst + 1, (unsigned long long)p); }
}

__device__ __forceinline__ void write_nn_input(float* dst, V3 p, V3 wo, V3 t, float scene_scale) {
    V3 point = p / scene_scale;
    float4* d4 = reinterpret_cast<float4*>(dst);
    d4[0] = make_float4(point.x, point.y, point.z, wo.x);
    d4[1] = make_float4(wo.y, wo.z, t.x, t.y);
    d4[2] = make_float4(t.z, 0.f, 0.f, 0.f);
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
// shade
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void push_probe(const FrameParams& P, const Probe& pr, int slot, int bit) {
    int idx = queue_reserve(P.q.counts + 3, pr.active);
    if (idx >= 0) {
        P.q.shadow[2 * (size_t)idx + 0] = make_float4(pr.o.x, pr.o.y, pr.o.z, __int_as_float(slot | (bit << 30)));
        P.q.shadow[2 * (size_t)idx + 1] = make_float4(pr.d.x, pr.d.y, pr.d.z, 0.f);
    }
}

__device__ __forceinline__ void fold_pending(const FrameParams& P, int slot, bool training, V3& color, V3& color_short) {
    float4 dl = P.paths.dl_light[slot];
    if (dl.w == 0.f) return;
    uint32_t vis = __ldcg(P.paths.vis + slot);   // cleared by whichever thread traced the probe: bypass L1
    V3 d = resolve_direct(v3(dl), (vis & 1u) != 0, v3(P.paths.dl_bsdf[slot]), (vis & 2u) != 0);
    color += v3(P.paths.dl_beta[slot]) * d;
    if (training) color_short += v3(P.paths.dl_beta_short[slot]) * d;
}

#ifndef HM_SHADE_GRID
#define HM_SHADE_GRID 8
#endif
// resident CTAs per SM = register cap of k_shade: 4 -> 128 regs 1.59 ms per frame, 6 -> 80 regs 1.48, 8 -> 64 regs 1.45
// (profiles/r2d_sweep_knobs.txt: the kernel stalls on instruction fetch, more warps hide it better than fewer spills)
#ifndef HM_SHADE_CTAS
#define HM_SHADE_CTAS 8
#endif
// One path vertex of render_path_tracing / render_hair_msnn: everything k_shade does for a live queue
// item except the queue pushes.  Reads and writes the slot's path state in HBM; the caller gets the two
// direct-light probes and whether a continuation ray was written to paths.ray_o / ray_d.
__device__ __forceinline__ void shade_item(const FrameParams& P, int slot, DirectSample& ds, bool& extend) {
    Rng rng; rng.state = P.paths.rng[slot];
    V3 ro = v3(P.paths.ray_o[slot]), rd = v3(P.paths.ray_d[slot]);
    float4 hr = __ldcg(P.paths.hit + slot);   // written by whichever thread traced the ray: bypass L1
    Hit hit; hit.t = hr.x; hit.prim = __float_as_int(hr.y); hit.u = hr.z; hit.v = hr.w;
    float4 b4 = P.paths.beta[slot];
    V3 beta = v3(b4);
    int bounces = __float_as_int(b4.w);
    V3 color = v3(P.paths.color[slot]);

    int tr_ofs = 0;
    bool training = false;
    V3 beta_short(1.f), color_short(0.f);
    if (P.mode == MODE_MSNN) {
        // merged tail pieces: only training paths outlive the main piece (bounces > beta ends the others), and
        // nothing past the first vertex needs the record index
        training = P.tail_merged ? true : is_training_pixel(P, slot, tr_ofs);
        if (training) {
            beta_short = v3(P.paths.beta_short[slot]);
            color_short = v3(P.paths.color_short[slot]);
        }
    }

    fold_pending(P, slot, training, color, color_short);

    Vertex v = vertex_from_hit(P.scene, hit, ro, rd);

    if (P.mode == MODE_MSNN && bounces == 0) {
        write_nn_input(P.nn_frame_in + (size_t)slot * P.in_ch, v.p, v.wo, v.t, P.scene.scene_scale);
        P.gbuffer[slot].w = __int_as_float(1 | (v.surface ? 2 : 0));
        if (training && tr_ofs >= P.train_slot0 && tr_ofs < P.train_slot0 + P.train_slots)
            write_nn_input(P.nn_train_in + (size_t)tr_ofs * P.in_ch, v.p, v.wo, v.t, P.scene.scene_scale);
    }

    // direct lighting
    const bool degenerate = P.mode == MODE_PT && P.v2_stop < P.v1_stop;   // pathTrace returns 0
    bool do_dl = (P.mode == MODE_PT) ? (bounces >= P.v1_stop && !degenerate) : true;
    if (do_dl) {
        sample_direct(P.scene, v, rng, ds);
        P.paths.dl_beta[slot] = f4(beta, 0.f);
        if (training) P.paths.dl_beta_short[slot] = f4(beta_short, 0.f);
        P.paths.dl_light[slot] = f4(ds.light.value, 1.f);
        P.paths.dl_bsdf[slot] = f4(ds.bsdf.value, 0.f);
        P.paths.vis[slot] = (ds.light.active ? 1u : 0u) | (ds.bsdf.active ? 2u : 0u);
    } else {
        P.paths.dl_light[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
    }

    // Russian roulette (after the direct sample of every vertex but the first)
    bool alive = !degenerate;
    if (bounces >= 1) {
        float q = fmaxf(0.05f, 1.f - luminance709(beta));
        if (training) {
            float qs = fmaxf(0.05f, 1.f - luminance709(beta_short));
            float eps = rng_next(rng);
            if (eps < qs || bounces > P.msnn_beta) beta_short = V3(0.f);
            if (eps < q) alive = false;
            else {
                beta = beta / (1.f - q);
                if (!(beta_short == V3(0.f))) beta_short = beta_short / (1.f - qs);
            }
        } else {
            float eps = rng_next(rng);
            if (eps < q) alive = false;
            else if (P.mode == MODE_MSNN && bounces > P.msnn_beta) alive = false;
            else beta = beta / (1.f - q);
        }
    }
    if (alive && bounces + 1 > P.v2_stop) alive = false;

    if (alive) {
        V3 no, nd;
        V3 mul = sample_continuation(P.scene, v, rng, no, nd);
        beta = beta * mul;
        if (training) beta_short = beta_short * mul;
        P.paths.ray_o[slot] = f4(no, 0.f);
        P.paths.ray_d[slot] = f4(nd, 0.f);
        extend = true;
    }
    P.paths.rng[slot] = rng.state;
    P.paths.beta[slot] = f4(beta, __int_as_float(bounces + 1));
    P.paths.color[slot] = f4(color, 0.f);
    if (training) {
        P.paths.beta_short[slot] = f4(beta_short, 0.f);
        P.paths.color_short[slot] = f4(color_short, 0.f);
    }
}

__global__ void __launch_bounds__(kBlock, HM_SHADE_CTAS) k_shade(const __grid_constant__ FrameParams P, int src) {
    const int n = P.q.counts[src];
    const int* queue = P.q.shade[src];
    const int rounds = (n + kBlock - 1) / kBlock;
    for (int r = blockIdx.x; r < rounds; r += gridDim.x) {
        int i = r * kBlock + threadIdx.x;
        bool live = i < n;
        int slot = live ? queue[i] : 0;

        DirectSample ds;
        ds.light.active = false; ds.bsdf.active = false;
        bool extend = false;
        if (live) shade_item(P, slot, ds, extend);
        push_probe(P, ds.light, slot, 0);
        push_probe(P, ds.bsdf, slot, 1);
        int idx = queue_reserve(P.q.counts + 2, extend);
        if (idx >= 0) P.q.extend[idx] = slot;
    }
    if (P.collect_stats && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(P.q.trav + 8, (unsigned long long)n);
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
// tail piece of a HairMSNN frame in ONE launch
// ---------------------------------------------------------------------------------
// Past vertex beta+1 only the <= 16384 training paths are alive, for up to 38 more vertices.  As
// (shade, trace) launch pairs that is 76 tiny latency-bound launches per frame, each sweeping the GPU
// with persistent CTAs that find almost no work while other frames' main pieces are running (measured:
// 0.73 ms of a 5.0 ms frame).  Here a warp keeps 32 such paths and walks them to the end: shade the 32
// vertices (shade_item, the code k_shade runs), put their <= 96 rays into a warp-private list in shared
// memory, trace the list with the warp-cooperative traversal, repeat while any path is alive.  Path
// state stays in the HBM arrays between steps exactly as the launch-per-vertex form leaves it, so
// k_finalize and the parity tests see no difference.
constexpr int kTailRaysPerWarp = 96;

struct TailOps {
    const FrameParams& P;
    float4* rays;        // warp-private: [kTailRaysPerWarp][2]  (o.xyz, slot | bit << 30) (d.xyz, kind | lane << 8)
    int* alive;          // warp-private [32]: set when a lane's continuation ray hit something
    __device__ __forceinline__ bool fetch(int w, V3& o, V3& d) const {
        const float4 a = rays[2 * w], b = rays[2 * w + 1];
        o = V3(a.x, a.y, a.z);
        d = V3(b.x, b.y, b.z);
        return (__float_as_int(b.w) & 0xff) == 0;      // kind 0: occlusion probe
    }
    __device__ __forceinline__ void commit(int w, const Hit& h, bool finished) const {
        if (!finished || h.prim < 0) return;
        const int tag = __float_as_int(rays[2 * w].w), info = __float_as_int(rays[2 * w + 1].w);
        const int slot = tag & 0x3fffffff;
        if ((info & 0xff) == 0) {
            atomicAnd(P.paths.vis + slot, ~(1u << (tag >> 30)));
        } else {
            P.paths.hit[slot] = make_float4(h.t, __int_as_float(h.prim), h.u, h.v);
            alive[info >> 8] = 1;
        }
    }
};

__global__ void __launch_bounds__(kBlock) k_tail_mega(const __grid_constant__ FrameParams P, int src) {
    __shared__ float4 s_rays[kBlock / 32][kTailRaysPerWarp][2];
    __shared__ int s_alive[kBlock / 32][32];
    __shared__ int s_cursor[kBlock / 32];
    const int n = P.q.counts[src];
    const int* queue = P.q.shade[src];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    TraceStats st[2] = {{0, 0}, {0, 0}};
    unsigned long long n_shade = 0, n_extend = 0, n_shadow = 0;

    // persistent warps: a lane whose path has ended takes the next queue entry, so few warps stay resident
    // (8 frames are in flight; idle lanes of long-lived warps would pin registers the main pieces need)
    int* cursor = P.q.counts + 7;
    int slot = -1;
    bool live = false, exhausted = false;
    while (true) {
        const unsigned want = __ballot_sync(0xffffffffu, !live);
        if (want && !exhausted) {
            const int cnt = __popc(want);
            int base = 0;
            if (lane == 0) base = atomicAdd(cursor, cnt);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base + cnt >= n) exhausted = true;
            if (!live) {
                const int i = base + __popc(want & ((1u << lane) - 1u));
                if (i < n) { slot = queue[i]; live = true; }
            }
        }
        if (!__any_sync(0xffffffffu, live)) break;
        {
            DirectSample ds;
            ds.light.active = false; ds.bsdf.active = false;
            bool extend = false;
            if (live) shade_item(P, slot, ds, extend);
            // warp-private ray list: probes first (they end early), then continuation rays
            const unsigned m0 = __ballot_sync(0xffffffffu, ds.light.active), m1 = __ballot_sync(0xffffffffu, ds.bsdf.active);
            const unsigned m2 = __ballot_sync(0xffffffffu, extend);
            const unsigned below = (1u << lane) - 1u;
            const int n0 = __popc(m0), n1 = __popc(m1), n2 = __popc(m2);
            if (ds.light.active) {
                const int w = __popc(m0 & below);
                s_rays[wib][w][0] = make_float4(ds.light.o.x, ds.light.o.y, ds.light.o.z, __int_as_float(slot));
                s_rays[wib][w][1] = make_float4(ds.light.d.x, ds.light.d.y, ds.light.d.z, __int_as_float(lane << 8));
            }
            if (ds.bsdf.active) {
                const int w = n0 + __popc(m1 & below);
                s_rays[wib][w][0] = make_float4(ds.bsdf.o.x, ds.bsdf.o.y, ds.bsdf.o.z, __int_as_float(slot | (1 << 30)));
                s_rays[wib][w][1] = make_float4(ds.bsdf.d.x, ds.bsdf.d.y, ds.bsdf.d.z, __int_as_float(lane << 8));
            }
            if (extend) {
                const int w = n0 + n1 + __popc(m2 & below);
                const float4 o = P.paths.ray_o[slot], d = P.paths.ray_d[slot];
                s_rays[wib][w][0] = make_float4(o.x, o.y, o.z, __int_as_float(slot));
                s_rays[wib][w][1] = make_float4(d.x, d.y, d.z, __int_as_float(1 | (lane << 8)));
            }
            s_alive[wib][lane] = 0;
            if (lane == 0) s_cursor[wib] = 0;
            __syncwarp();
            n_shade += __popc(__ballot_sync(0xffffffffu, live));
            n_shadow += n0 + n1; n_extend += n2;
            TailOps ops{P, &s_rays[wib][0][0], s_alive[wib]};
            trace_queue(P.scene.geom, n0 + n1 + n2, &s_cursor[wib], ops, 0.f, 1e30f, P.collect_stats ? st : nullptr);
            __syncwarp();
            live = s_alive[wib][lane] != 0;
            __syncwarp();
        }
    }
    if (P.collect_stats) {
        flush_trav(P.q.trav + 0, st[0]);
        flush_trav(P.q.trav + 2, st[1]);
        TraceStats both{st[0].nodes + st[1].nodes, st[0].prims + st[1].prims};
        flush_trav(P.q.trav + 10, both);
        if (lane == 0 && (n_shade | n_extend | n_shadow)) {
            atomicAdd(P.q.trav + 6, n_extend);
            atomicAdd(P.q.trav + 7, n_shadow);
            atomicAdd(P.q.trav + 8, n_shade);
            atomicAdd(P.q.trav + 12, n_extend + n_shadow);
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

// ---------------------------------------------------------------------------------
// render_nrc: nrcTracePaths in wavefront form (cuda/nrc.cu:135-310)
// ---------------------------------------------------------------------------------
// The reference loop body for bounce b is: direct light at vertex b -> sample + trace to
// vertex b+1 -> [training pixel: record vertex b] -> spread update with the NEW vertex ->
// terminate / query the cache / continue.  Here the shade of vertex b first finishes
// bounce b-1 (everything that needed the new hit), then starts bounce b.  No Russian
// roulette: paths end on the spread heuristic, on leaving the scene, or at kNrcMaxBounces.
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

__device__ __forceinline__ void write_nrc_query(float* dst, V3 p, V3 wo, V3 n, float scene_scale) {
    V3 point = p / scene_scale;
    dst[0] = point.x; dst[1] = point.y; dst[2] = point.z;
    dst[3] = wo.x; dst[4] = wo.y; dst[5] = wo.z;
    dst[6] = n.x; dst[7] = n.y; dst[8] = n.z;
}

// pow(length(d), 2) / (4 Pi) / |cos|   (cuda/nrc.cu:153,186)
__device__ __forceinline__ float nrc_area(V3 a, V3 b, float abscos) {
    float l = length(a - b);
    return l * l / (4.f * kPi) / abscos;
}

__global__ void __launch_bounds__(kBlock) k_shade_nrc(const __grid_constant__ FrameParams P, int src) {
    const int n = P.q.counts[src];
    const int* queue = P.q.shade[src];
    const int rounds = (n + kBlock - 1) / kBlock;
    const size_t frame_size = (size_t)P.W * P.H;
    for (int r = blockIdx.x; r < rounds; r += gridDim.x) {
        int i = r * kBlock + threadIdx.x;
        bool live = i < n;
        int slot = live ? queue[i] : 0;

        DirectSample ds;
        ds.light.active = false; ds.bsdf.active = false;
        bool extend = false;

        if (live) {
            Rng rng; rng.state = P.paths.rng[slot];
            V3 ro = v3(P.paths.ray_o[slot]), rd = v3(P.paths.ray_d[slot]);
            float4 hr = P.paths.hit[slot];
            Hit hit; hit.t = hr.x; hit.prim = __float_as_int(hr.y); hit.u = hr.z; hit.v = hr.w;
            float4 b4 = P.paths.beta[slot];
            V3 beta = v3(b4);
            const int b = __float_as_int(b4.w);       // index of this vertex == bounce about to start
            V3 color = v3(P.paths.color[slot]);

            int tr_ofs = 0;
            bool unbiased = false;
            const bool training = nrc_training_pixel(P, slot, tr_ofs, unbiased);
            NrcTrainRec* rec = P.tbuffer + tr_ofs;

            // direct light of vertex b-1, now that its probes are back
            {
                float4 dl = P.paths.dl_light[slot];
                if (dl.w != 0.f) {
                    uint32_t vis = P.paths.vis[slot];
                    V3 d = resolve_direct(v3(dl), (vis & 1u) != 0, v3(P.paths.dl_bsdf[slot]), (vis & 2u) != 0);
                    color += v3(P.paths.dl_beta[slot]) * d;
                    if (training) write3(rec->radiance[b - 1], d);
                }
            }

            Vertex v = vertex_from_hit(P.scene, hit, ro, rd);
            const float abscos = fabsf(v.wo_local.z);

            float spread = 0.f, a0 = 0.f, c = P.nrc_c;
            int flags = 0;
            bool terminated = false;
            if (b == 0) {
                if (v.surface && v.wo_local.z < 0.f) {
                    // a head triangle seen from behind counts as a miss (cuda/nrc.cu:353-360); si.Le is 0 on a hit
                    P.gbuffer[slot] = make_float4(0.f, 0.f, 0.f, __int_as_float(0));
                    P.gbuffer_b[slot] = make_float4(0.f, 0.f, 0.f, __int_as_float(0));
                    terminated = true;
                } else {
                    a0 = nrc_area(v.p, V3(P.cam.pos[0], P.cam.pos[1], P.cam.pos[2]), abscos);
                }
            } else {
                float4 st = P.paths.nrc_state[slot];
                spread = st.x; a0 = st.y; c = st.z; flags = __float_as_int(st.w);
                float4 pv = P.paths.nrc_prev[slot];
                const V3 prev_point = v3(pv);
                if (training && (flags & kNrcRecalcA0)) {
                    a0 = nrc_area(v.p, prev_point, abscos);
                    flags &= ~kNrcRecalcA0;
                }
                {   // nrcSpread (cuda_headers/utils.cuh:86-90)
                    float l = length(v.p - prev_point);
                    spread = spread + sqrtf(l * l / pv.w / abscos);
                }
                const bool cond = spread * spread > c * a0;
                if (cond && !(flags & kNrcSuffix)) {
                    P.gbuffer[slot] = f4(color, __int_as_float(1));
                    P.gbuffer_b[slot] = f4(beta, __int_as_float(b - 1));
                    write_nrc_query(P.nn_frame_in + (size_t)slot * P.in_ch, v.p, v.wo, v.n, P.scene.scene_scale);
                    if (training) { flags |= kNrcSuffix | kNrcRecalcA0; spread = 0.f; }
                    else terminated = true;
                } else if (cond) {
                    write_nrc_query(P.nn_frame_in + (frame_size + tr_ofs) * P.in_ch, v.p, v.wo, v.n, P.scene.scene_scale);
                    rec->bounces = b - 1;
                    rec->hit = 1;
                    if (!unbiased) terminated = true;
                    else c = 1e30f;
                }
            }

            if (terminated) {
                flags |= kNrcTerminated;
                P.paths.dl_light[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
                sample_direct(P.scene, v, rng, ds);
                P.paths.dl_beta[slot] = f4(beta, 0.f);
                P.paths.dl_light[slot] = f4(ds.light.value, 1.f);
                P.paths.dl_bsdf[slot] = f4(ds.bsdf.value, 0.f);
                P.paths.vis[slot] = (ds.light.active ? 1u : 0u) | (ds.bsdf.active ? 2u : 0u);

                V3 no, nd;
                float pdf = 1.f;
                V3 mul = sample_continuation(P.scene, v, rng, no, nd, &pdf);
                beta = beta * mul;
                P.paths.nrc_prev[slot] = f4(v.p, pdf);
                if (training) {
                    write3(rec->vert[b], v.p / P.scene.scene_scale);
                    write3(rec->wo[b], v.wo);
                    write3(rec->n[b], v.n);
                    write3(rec->beta[b], mul);
                }
                if (b < kNrcMaxBounces - 1) {   // the last bounce's ray cannot change anything (cuda/nrc.cu:200)
                    P.paths.ray_o[slot] = f4(no, 0.f);
                    P.paths.ray_d[slot] = f4(nd, 0.f);
                    extend = true;
                }
            }
            P.paths.rng[slot] = rng.state;
            P.paths.beta[slot] = f4(beta, __int_as_float(b + 1));
            P.paths.color[slot] = f4(color, 0.f);
            P.paths.nrc_state[slot] = make_float4(spread, a0, c, __int_as_float(flags));
        }
        push_probe(P, ds.light, slot, 0);
        push_probe(P, ds.bsdf, slot, 1);
        int idx = queue_reserve(P.q.counts + 2, extend);
        if (idx >= 0) P.q.extend[idx] = slot;
    }
    if (P.collect_stats && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(P.q.trav + 8, (unsigned long long)n);
}

// End of the G_BUFFER pass: paths that left the scene (or reached the bounce cap) fold their last
// direct sample and write their G-buffer entry; primary misses show the environment.
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

int persistent_grid(int ctas_per_sm) { return wavefront_sm_count() * ctas_per_sm; }

}  // namespace

void launch_primary(const FrameParams& P, cudaStream_t stream) {
    int n = P.n_primary;
    int blocks = (n + kBlock - 1) / kBlock;
    int grid = blocks < persistent_grid(kTraceCtasPerSm) ? blocks : persistent_grid(kTraceCtasPerSm);
    if (grid < 1) grid = 1;
    k_primary<<<grid, kBlock, 0, stream>>>(P);
    g_launches++;
}
// max_items > 0: the caller knows an upper bound of the queue length (tail pieces carry the few long paths
// only) — the grid is sized for it instead of for the whole GPU, so these launches do not sweep every SM
// with CTAs that find no work while other frames' main pieces are running.
static int bounded_grid(int full, long long max_items, int items_per_cta) {
    if (max_items <= 0) return full;
    long long need = (max_items + items_per_cta - 1) / items_per_cta;
    if (need < 1) need = 1;
    return need < full ? (int)need : full;
}
void launch_shade(const FrameParams& P, int src, cudaStream_t stream, long long max_items) {
    if (P.mode == MODE_NRC) k_shade_nrc<<<bounded_grid(persistent_grid(8), max_items, kBlock), kBlock, 0, stream>>>(P, src);
    else k_shade<<<bounded_grid(persistent_grid(HM_SHADE_GRID), max_items, kBlock), kBlock, 0, stream>>>(P, src);
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
#ifndef HM_TAIL_CTAS
#define HM_TAIL_CTAS 37      // x 4 warps x 32 lanes = 4736 paths in flight per frame
#endif
void launch_tail_mega(const FrameParams& P, int src, int max_paths, cudaStream_t stream) {
    int grid = (max_paths + kBlock - 1) / kBlock;
    if (grid < 1) grid = 1;
    if (grid > HM_TAIL_CTAS) grid = HM_TAIL_CTAS;
    k_tail_mega<<<grid, kBlock, 0, stream>>>(P, src);
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

// Test hooks: the fibre scattering model for caller-supplied local directions (device pointers).
__global__ void __launch_bounds__(256) k_bsdf_eval(const HairLobes L, const float* wo, const float* wi, const float* h, int n,
                                                   float* out_f, float* out_pdf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float pdf;
    const V3 f = hair_eval(L, V3(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]), V3(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]), h[i], &pdf);
    out_f[3 * i] = f.x; out_f[3 * i + 1] = f.y; out_f[3 * i + 2] = f.z; out_pdf[i] = pdf;
}
__global__ void __launch_bounds__(256) k_bsdf_sample(const HairLobes L, const float* wo, const float* h, const float* u, int n,
                                                     float* out_wi, float* out_f, float* out_pdf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const V3 o(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]);
    const V3 w = hair_sample_dir(L, o, h[i], u[4 * i], u[4 * i + 1], u[4 * i + 2], u[4 * i + 3]);
    float pdf;
    const V3 f = hair_eval(L, o, w, h[i], &pdf);
    out_wi[3 * i] = w.x; out_wi[3 * i + 1] = w.y; out_wi[3 * i + 2] = w.z;
    out_f[3 * i] = f.x; out_f[3 * i + 1] = f.y; out_f[3 * i + 2] = f.z; out_pdf[i] = pdf;
}
void launch_bsdf_eval(const HairLobes& L, const float* wo, const float* wi, const float* h, int n, float* out_f, float* out_pdf, cudaStream_t stream) {
    k_bsdf_eval<<<(n + 255) / 256, 256, 0, stream>>>(L, wo, wi, h, n, out_f, out_pdf);
    g_launches++;
}
void launch_bsdf_sample(const HairLobes& L, const float* wo, const float* h, const float* u, int n, float* out_wi, float* out_f, float* out_pdf,
                        cudaStream_t stream) {
    k_bsdf_sample<<<(n + 255) / 256, 256, 0, stream>>>(L, wo, h, u, n, out_wi, out_f, out_pdf);
    g_launches++;
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
