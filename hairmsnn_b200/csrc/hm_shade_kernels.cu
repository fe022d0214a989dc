// hm_shade_kernels.cu — the shading kernels of the wavefront path tracer (one path vertex per queue item).
//
// A translation unit of its own so that it can be built with other floating-point flags than the traversal:
// `make SHADE_APPROX=1` compiles it with -prec-div=false -prec-sqrt=false.  On the real scene k_shade spends a fifth of
// its 7600 instructions in the IEEE division / square-root sequences of normalisations and vector quotients (hm_math.h)
// and a third of its stall samples waiting for instructions (119 KB of code); approximate (2 ulp) forms make the frame
// 4.2 % faster (185.3 -> 193.0 Mpaths/s, k_shade 1.16 -> 0.92 ms) and stay inside every stated tolerance (the two
// quotients of the hair model that feed an exponential at magnitudes of several hundred are div_exact, hm_bsdf.h) — but
// each path vertex then differs from the reference's host build by 2 ulp instead of half an ulp in ~46 quotients, and
// more paths take a different discrete branch somewhere along their 40 vertices.  The default build keeps IEEE
// arithmetic here; traversal and everything compared bit for bit with the host build lives in hm_wavefront.cu.
#include "hm_wavefront_dev.cuh"

namespace hm {

namespace {

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
// render_nrc: nrcTracePaths in wavefront form (cuda/nrc.cu:135-310)
// ---------------------------------------------------------------------------------
// The reference loop body for bounce b is: direct light at vertex b -> sample + trace to
// vertex b+1 -> [training pixel: record vertex b] -> spread update with the NEW vertex ->
// terminate / query the cache / continue.  Here the shade of vertex b first finishes
// bounce b-1 (everything that needed the new hit), then starts bounce b.  No Russian
// roulette: paths end on the spread heuristic, on leaving the scene, or at kNrcMaxBounces.
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
}  // namespace

void launch_shade(const FrameParams& P, int src, cudaStream_t stream, long long max_items) {
    if (P.mode == MODE_NRC) k_shade_nrc<<<bounded_grid(persistent_grid(8), max_items, kBlock), kBlock, 0, stream>>>(P, src);
    else k_shade<<<bounded_grid(persistent_grid(HM_SHADE_GRID), max_items, kBlock), kBlock, 0, stream>>>(P, src);
    wavefront_count_launch();
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
    wavefront_count_launch();
}
void launch_bsdf_sample(const HairLobes& L, const float* wo, const float* h, const float* u, int n, float* out_wi, float* out_f, float* out_pdf,
                        cudaStream_t stream) {
    k_bsdf_sample<<<(n + 255) / 256, 256, 0, stream>>>(L, wo, h, u, n, out_wi, out_f, out_pdf);
    wavefront_count_launch();
}
}  // namespace hm
