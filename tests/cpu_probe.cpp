// cpu_probe.cpp — test-only host build of the product's shared host/device headers
// (hm_rng.h, hm_bsdf.h, hm_curve.h, hm_bvh.h).  Lets the CPU test-suite check the
// arithmetic the CUDA kernels run against the oracle without a GPU.
#include <algorithm>
#include <cstring>
#include <vector>

#include "../hairmsnn_b200/csrc/hm_bsdf.h"
#include "../hairmsnn_b200/csrc/hm_host.h"
#include "../hairmsnn_b200/csrc/hm_light.h"
#include "../hairmsnn_b200/csrc/hm_rng.h"

using namespace hm;

extern "C" {

uint32_t probe_rng_seed(int frame_id, uint32_t px, uint32_t py, uint32_t w) { return rng_seed(frame_id, px, py, w).state; }
void probe_rng_draws(uint32_t state, int n, uint32_t* states, float* floats) {
    Rng r; r.state = state;
    for (int i = 0; i < n; ++i) { floats[i] = rng_next(r); states[i] = r.state; }
}

void probe_hair_setup(float bm, float bn, float alpha, float* out) {
    HairLobes L; L.setup(bm, bn, alpha);
    for (int i = 0; i < 3; ++i) { out[i] = L.v[i]; out[4 + i] = L.sin2k[i]; out[7 + i] = L.cos2k[i]; }
    out[3] = L.s;
}
static HairLobes make_lobes(const float* sigma_a, float bm, float bn, float alpha) {
    HairLobes L; L.setup(bm, bn, alpha);
    L.sigma_a = V3(sigma_a[0], sigma_a[1], sigma_a[2]);
    for (int i = 0; i < 4; ++i) L.gain[i] = 1.f;
    return L;
}
void probe_hair_eval(int n, const float* wo, const float* wi, const float* h, const float* sigma_a,
                     float bm, float bn, float alpha, float* out_f, float* out_pdf) {
    HairLobes L = make_lobes(sigma_a, bm, bn, alpha);
    for (int i = 0; i < n; ++i) {
        float pdf;
        V3 f = hair_eval(L, V3(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]), V3(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]), h[i], &pdf);
        out_f[3 * i] = f.x; out_f[3 * i + 1] = f.y; out_f[3 * i + 2] = f.z; out_pdf[i] = pdf;
    }
}
void probe_hair_sample(int n, const float* wo, const float* h, const float* u, const float* sigma_a,
                       float bm, float bn, float alpha, float* out_wi, float* out_f, float* out_pdf) {
    HairLobes L = make_lobes(sigma_a, bm, bn, alpha);
    for (int i = 0; i < n; ++i) {
        V3 o(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]);
        V3 w = hair_sample_dir(L, o, h[i], u[4 * i], u[4 * i + 1], u[4 * i + 2], u[4 * i + 3]);
        float pdf;
        V3 f = hair_eval(L, o, w, h[i], &pdf);
        out_wi[3 * i] = w.x; out_wi[3 * i + 1] = w.y; out_wi[3 * i + 2] = w.z;
        out_f[3 * i] = f.x; out_f[3 * i + 1] = f.y; out_f[3 * i + 2] = f.z; out_pdf[i] = pdf;
    }
}
void probe_surf_eval(int n, const float* wo, const float* wi, const float* kd, float alpha, float* out_f, float* out_pdf) {
    for (int i = 0; i < n; ++i) {
        V3 o(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]), w(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]);
        V3 f = surf_eval(o, w, V3(kd[0], kd[1], kd[2]), alpha);
        out_f[3 * i] = f.x; out_f[3 * i + 1] = f.y; out_f[3 * i + 2] = f.z;
        out_pdf[i] = surf_pdf(alpha, o, normalize(o + w));
    }
}
void probe_surf_sample(int n, const float* wo, const float* u, float alpha, float* out_wi, float* out_pdf) {
    for (int i = 0; i < n; ++i) {
        float pdf;
        V3 w = surf_sample(u[2 * i], u[2 * i + 1], alpha, V3(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]), &pdf);
        out_wi[3 * i] = w.x; out_wi[3 * i + 1] = w.y; out_wi[3 * i + 2] = w.z; out_pdf[i] = pdf;
    }
}
void probe_curve_geometry(const float* cps16, const float* org, const float* dir, float t, float u, float* out13) {
    CubicSeg s;
    s.from_catmull_rom(V4(cps16[0], cps16[1], cps16[2], cps16[3]), V4(cps16[4], cps16[5], cps16[6], cps16[7]),
                       V4(cps16[8], cps16[9], cps16[10], cps16[11]), V4(cps16[12], cps16[13], cps16[14], cps16[15]));
    V3 o(org[0], org[1], org[2]), d(dir[0], dir[1], dir[2]);
    FibreHit h = fibre_hit_geometry(s, u, o + t * d);
    out13[0] = h.p.x; out13[1] = h.p.y; out13[2] = h.p.z; out13[3] = h.n.x; out13[4] = h.n.y; out13[5] = h.n.z;
    out13[6] = h.t.x; out13[7] = h.t.y; out13[8] = h.t.z; out13[9] = h.centre.x; out13[10] = h.centre.y; out13[11] = h.centre.z;
    out13[12] = h.radius;
}
// single-segment intersection; returns 1 on hit
int probe_intersect_fibre(const float* cps16, const float* org, const float* dir, float tmin, float tmax, float* t, float* u) {
    V3 o(org[0], org[1], org[2]), d(dir[0], dir[1], dir[2]);
    RayFrame rf = make_ray_frame(o, d);
    SegHit sh;
    bool ok = intersect_fibre(rf, tmin, tmax, V4(cps16[0], cps16[1], cps16[2], cps16[3]), V4(cps16[4], cps16[5], cps16[6], cps16[7]),
                              V4(cps16[8], cps16[9], cps16[10], cps16[11]), V4(cps16[12], cps16[13], cps16[14], cps16[15]), sh);
    if (ok) { *t = sh.t; *u = sh.u; }
    return ok ? 1 : 0;
}

// environment CDF search: binary (std::lower_bound order) and 4-ary variants on the same table
void probe_cdf_search(const float* table, int w, int h, int n, const float* u, const float* yn, float size, int* out2) {
    for (int i = 0; i < n; ++i) {
        out2[2 * i + 0] = cdf_lower_bound(u[i], table, w, h, yn[i], size);
        out2[2 * i + 1] = cdf_search(u[i], table, w, h, yn[i], size);
    }
}

// two-level search (power-of-two rows) vs the reference-order search, row by row
void probe_cdf_two_level(const float* table, int W, int H, int n, const float* u, const int* rows, int* out2) {
    const int cw = W + 1, K = W >> 6;
    std::vector<float> coarse((size_t)K * H);
    for (int y = 0; y < H; ++y)
        for (int k = 0; k < K; ++k) coarse[(size_t)y * K + k] = table[(size_t)y * cw + 64 * k];
    for (int i = 0; i < n; ++i) {
        const int y = rows[i];
        out2[2 * i + 0] = cdf_lower_bound(u[i], table, cw, H, (y + 0.5f) / H, (float)W);
        out2[2 * i + 1] = cdf_lower_bound_two_level(u[i], table + (size_t)y * cw, coarse.data() + (size_t)y * K, W);
    }
}

// BVH build + trace on the host
struct ProbeScene { HostGeometry geo; HostBvh bvh; };
void* probe_scene_create(const float* cps, int ncps, const int* seg_cp, int nseg, const float* tri_verts, int ntri, int threads) {
    ProbeScene* s = new ProbeScene;
    s->geo.cps.resize(ncps);
    memcpy(s->geo.cps.data(), cps, sizeof(float) * 4 * ncps);
    s->geo.seg_cp.assign(seg_cp, seg_cp + nseg);
    s->geo.tri_verts.resize(3 * (size_t)ntri);
    if (ntri) memcpy(s->geo.tri_verts.data(), tri_verts, sizeof(float) * 12 * ntri);
    build_bvh(s->geo, s->bvh, threads);
    return s;
}
void probe_scene_destroy(void* p) { delete (ProbeScene*)p; }
int probe_scene_num_nodes(void* p) { return (int)(((ProbeScene*)p)->bvh.nodes.size() / 4); }
void probe_scene_arrays(void* p, const float** nodes, const int** leaf_code, const int** leaf_prim) {
    ProbeScene* s = (ProbeScene*)p;
    *nodes = (const float*)s->bvh.nodes.data(); *leaf_code = s->bvh.leaf_code.data(); *leaf_prim = s->bvh.leaf_prim.data();
}
// out per ray: t, prim (as float bits int), u, v ; stats: nodes, prims
void probe_trace(void* p, int n, const float* org, const float* dir, float tmin, float tmax, int any,
                 float* out_t, int* out_prim, float* out_u, float* out_v, int* out_nodes, int* out_prims) {
    ProbeScene* s = (ProbeScene*)p;
    GeomView g = make_view(s->geo, s->bvh);
    for (int i = 0; i < n; ++i) {
        TraceStats st{0, 0};
        V3 o(org[3 * i], org[3 * i + 1], org[3 * i + 2]), d(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]);
        Hit h = any ? trace<true>(g, o, d, tmin, tmax, &st) : trace<false>(g, o, d, tmin, tmax, &st);
        out_t[i] = h.t; out_prim[i] = h.prim; out_u[i] = h.u; out_v[i] = h.v;
        if (out_nodes) { out_nodes[i] = st.nodes; out_prims[i] = st.prims; }
    }
}
// same through the 8-wide quantised tree (what the CUDA kernels traverse)
int probe_scene_num_wide_nodes(void* p) { return (int)(((ProbeScene*)p)->bvh.wnodes.size() / 5); }
int probe_scene_wide_depth(void* p) { return ((ProbeScene*)p)->bvh.wide_depth; }
void probe_trace_wide(void* p, int n, const float* org, const float* dir, float tmin, float tmax, int any,
                      float* out_t, int* out_prim, float* out_u, float* out_v, int* out_nodes, int* out_prims) {
    ProbeScene* s = (ProbeScene*)p;
    GeomView g = make_view(s->geo, s->bvh);
    for (int i = 0; i < n; ++i) {
        TraceStats st{0, 0};
        V3 o(org[3 * i], org[3 * i + 1], org[3 * i + 2]), d(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]);
        Hit h = any ? trace_wide<true>(g, o, d, tmin, tmax, &st) : trace_wide<false>(g, o, d, tmin, tmax, &st);
        out_t[i] = h.t; out_prim[i] = h.prim; out_u[i] = h.u; out_v[i] = h.v;
        if (out_nodes) { out_nodes[i] = st.nodes; out_prims[i] = st.prims; }
    }
}
// experiment: wide-tree traversal visiting hit children in true entry-distance order (upper bound on what
// a better child order could save over the octant order); statistics only
struct DistEntry { float t; int code; float tg; };   // code >= 0 inner node, < 0 leaf ref (~ref)
void probe_trace_wide_sorted(void* p, int n, const float* org, const float* dir, float tmin, float tmax,
                             int* out_prim, int* out_nodes, int* out_prims, int mode) {
    // mode bit 0: sort children by entry distance (else octant order); bit 1: cull stale entries when popped;
    // bit 2: octant order but the nearest hit child first; bit 3: cull by the node's minimum entry distance only
    ProbeScene* s = (ProbeScene*)p;
    GeomView g = make_view(s->geo, s->bvh);
    for (int i = 0; i < n; ++i) {
        V3 o(org[3 * i], org[3 * i + 1], org[3 * i + 2]), d(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]);
        WideRay wr = make_wide_ray(o, d);
        RayFrame rf = make_ray_frame(o, d);
        Hit best; best.t = tmax; best.prim = -1; best.u = 0; best.v = 0;
        std::vector<DistEntry> stack;
        stack.push_back({0.f, 0, 0.f});
        int nodes = 0, prims = 0;
        while (!stack.empty()) {
            DistEntry e = stack.back(); stack.pop_back();
            if ((mode & 2) && e.t > best.t) continue;
            if ((mode & 8) && e.tg > best.t) continue;
            if (e.code < 0) { prims++; test_wide_leaf(g, ~e.code, o, d, rf, tmin, best); continue; }
            const F4* nd = g.wnodes + 5 * (size_t)e.code;
            nodes++;
            unsigned imask = f_as_u(nd[0].w) >> 24, lmask = f_as_u(nd[1].z) & 0xff;
            unsigned h = wide_node_hits(nd[0], nd[2], nd[3], nd[4], wr, tmin, best.t) & (imask | lmask);
            // entry distance per hit child (recomputed the slow way)
            const unsigned em = f_as_u(nd[0].w);
            float cell[3] = {u_as_f((em & 0xff) << 23), u_as_f(((em >> 8) & 0xff) << 23), u_as_f(((em >> 16) & 0xff) << 23)};
            float org3[3] = {nd[0].x, nd[0].y, nd[0].z};
            float oo[3] = {o.x, o.y, o.z}, id[3] = {wr.idir.x, wr.idir.y, wr.idir.z};
            const unsigned words[12] = {f_as_u(nd[2].x), f_as_u(nd[2].y), f_as_u(nd[2].z), f_as_u(nd[2].w), f_as_u(nd[3].x), f_as_u(nd[3].y),
                                        f_as_u(nd[3].z), f_as_u(nd[3].w), f_as_u(nd[4].x), f_as_u(nd[4].y), f_as_u(nd[4].z), f_as_u(nd[4].w)};
            DistEntry found[8]; int nf = 0;
            for (int sl = 0; sl < 8; ++sl) {
                if (!((h >> sl) & 1)) continue;
                float tn = tmin;
                for (int a = 0; a < 3; ++a) {
                    unsigned qlo = (words[2 * a + (sl >> 2)] >> (8 * (sl & 3))) & 0xff, qhi = (words[6 + 2 * a + (sl >> 2)] >> (8 * (sl & 3))) & 0xff;
                    float lo = org3[a] + qlo * cell[a], hi = org3[a] + qhi * cell[a];
                    float t0 = (lo - oo[a]) * id[a], t1 = (hi - oo[a]) * id[a];
                    tn = fmaxf(tn, fminf(t0, t1));
                }
                unsigned below = (1u << sl) - 1u;
                int code = ((imask >> sl) & 1) ? f_as_i(nd[1].x) + popc_u(imask & below) : ~(f_as_i(nd[1].y) + popc_u(lmask & below));
                found[nf++] = {tn, code, 0.f};
            }
            if (mode & 1) std::sort(found, found + nf, [](const DistEntry& a, const DistEntry& b) { return a.t > b.t; });   // far first: near popped first
            else {
                // octant order: priority of slot sl = sl ^ octinv, highest first -> push lowest priority first
                DistEntry tmp[8]; int order[8]; int k = 0;
                int slots[8]; { int q = 0; for (int sl = 0; sl < 8; ++sl) if ((h >> sl) & 1) slots[q++] = sl; }
                for (int pr = 0; pr < 8; ++pr) for (int q = 0; q < nf; ++q) if ((slots[q] ^ wr.octinv) == pr) order[k++] = q;
                for (int q = 0; q < nf; ++q) tmp[q] = found[order[q]];
                for (int q = 0; q < nf; ++q) found[q] = tmp[q];
            }
            if ((mode & 4) && nf > 1) {   // nearest to the end of the list (= popped first), the rest keeps its order
                int best_k = 0;
                for (int k = 1; k < nf; ++k) if (found[k].t < found[best_k].t) best_k = k;
                DistEntry near = found[best_k];
                for (int k = best_k; k + 1 < nf; ++k) found[k] = found[k + 1];
                found[nf - 1] = near;
            }
            float tg = 3e38f;
            // bit 4: the group minimum excludes the nearest child (which is visited right away)
            for (int k = 0; k < nf - ((mode & 16) ? 1 : 0); ++k) tg = fminf(tg, found[k].t);
            for (int k = 0; k < nf; ++k) { found[k].tg = ((mode & 16) && k == nf - 1) ? found[k].t : tg; stack.push_back(found[k]); }
        }
        out_prim[i] = best.prim; out_nodes[i] = nodes; out_prims[i] = prims;
    }
}
// exhaustive reference: test every primitive, no BVH
void probe_trace_brute(void* p, int n, const float* org, const float* dir, float tmin, float tmax,
                       float* out_t, int* out_prim, float* out_u) {
    ProbeScene* s = (ProbeScene*)p;
    int ns = (int)s->geo.seg_cp.size(), nt = (int)(s->geo.tri_verts.size() / 3);
    for (int i = 0; i < n; ++i) {
        V3 o(org[3 * i], org[3 * i + 1], org[3 * i + 2]), d(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]);
        RayFrame rf = make_ray_frame(o, d);
        float best = tmax; int bp = -1; float bu = 0;
        for (int k = 0; k < ns; ++k) {
            const F4* c = s->geo.cps.data() + s->geo.seg_cp[k];
            SegHit sh;
            if (intersect_fibre(rf, tmin, best, f4_to_v4(c[0]), f4_to_v4(c[1]), f4_to_v4(c[2]), f4_to_v4(c[3]), sh)) { best = sh.t; bp = k; bu = sh.u; }
        }
        for (int k = 0; k < nt; ++k) {
            const F4* v = s->geo.tri_verts.data() + 3 * k;
            float t, b1, b2;
            if (intersect_triangle(o, d, tmin, best, V3(v[0].x, v[0].y, v[0].z), V3(v[1].x, v[1].y, v[1].z), V3(v[2].x, v[2].y, v[2].z), t, b1, b2)) { best = t; bp = ns + k; bu = b1; }
        }
        out_t[i] = best; out_prim[i] = bp; out_u[i] = bu;
    }
}

}  // extern "C"

// ---------------------------------------------------------------------------------
// Host megakernel over the product's shading header (hm_shade.h): the same vertex /
// direct-light / continuation code the CUDA shade kernel runs, driven depth-first
// with inline traversal.  Lets the CPU suite compare the shading logic with the
// reference host build pixel by pixel, independent of the wavefront plumbing.
// ---------------------------------------------------------------------------------
#include <thread>
#include "../hairmsnn_b200/csrc/hm_shade.h"

extern "C" {

struct ProbeSceneDesc {
    const float* nodes; int num_nodes; const int* leaf_code; const int* leaf_prim; const float* leaf_data;
    const float* cps; const float* tri_verts; const float* tri_normals; const int* seg_cp;
    int num_segments, num_tris;
    const float* env; const float* cpdf; const float* ccdf; const float* mpdf; const float* mcdf;
    int env_w, env_h; float env_scale, env_rot; int has_env, env_pdf;
    int num_dlights; const float* dl_from; const float* dl_emit;
    float sigma_a[3]; float beta_m, beta_n, alpha; float gains[4];
    float kd[3]; float surf_alpha; float scene_scale; int mis;
    float cam_pos[3], cam_d00[3], cam_du[3], cam_dv[3];
};

static SceneView make_scene_view(const ProbeSceneDesc& d) {
    SceneView S;
    memset(&S, 0, sizeof(S));
    S.geom.nodes = (const F4*)d.nodes; S.geom.num_nodes = d.num_nodes;
    S.geom.leaf_code = d.leaf_code; S.geom.leaf_prim = d.leaf_prim; S.geom.leaf_data = (const F4*)d.leaf_data;
    S.geom.cps = (const F4*)d.cps; S.geom.tri_verts = (const F4*)d.tri_verts;
    S.geom.num_segments = d.num_segments; S.geom.num_tris = d.num_tris;
    S.seg_cp = d.seg_cp; S.tri_normals = (const F4*)d.tri_normals;
    S.lights.env.env = d.env; S.lights.env.cpdf = d.cpdf; S.lights.env.ccdf = d.ccdf;
    S.lights.env.mpdf = d.mpdf; S.lights.env.mcdf = d.mcdf;
    S.lights.env.W = d.env_w; S.lights.env.H = d.env_h; S.lights.env.scale = d.env_scale; S.lights.env.rot_phi = d.env_rot;
    S.lights.env.has_env = d.has_env; S.lights.env.pdf_sampling = d.env_pdf;
    S.lights.num_dlights = d.num_dlights; S.lights.num_total = d.num_dlights + (d.has_env ? 1 : 0);
    for (int i = 0; i < d.num_dlights; ++i)
        for (int k = 0; k < 3; ++k) { S.lights.dl_from[i][k] = d.dl_from[3 * i + k]; S.lights.dl_emit[i][k] = d.dl_emit[3 * i + k]; }
    S.lobes.setup(d.beta_m, d.beta_n, d.alpha);
    S.lobes.sigma_a = V3(d.sigma_a[0], d.sigma_a[1], d.sigma_a[2]);
    for (int i = 0; i < 4; ++i) S.lobes.gain[i] = d.gains[i];
    for (int k = 0; k < 3; ++k) S.kd[k] = d.kd[k];
    S.surf_alpha = d.surf_alpha; S.scene_scale = d.scene_scale; S.mis = d.mis;
    return S;
}

static V3 probe_direct(const SceneView& S, const Vertex& v, Rng& rng) {
    DirectSample ds;
    sample_direct(S, v, rng, ds);
    bool va = false, vb = false;
    if (ds.light.active) va = trace<true>(S.geom, ds.light.o, ds.light.d, 0.f, 1e30f).prim < 0;
    if (ds.bsdf.active) vb = trace<true>(S.geom, ds.bsdf.o, ds.bsdf.d, 0.f, 1e30f).prim < 0;
    return resolve_direct(ds.light.value, va, ds.bsdf.value, vb);
}

// out: float[W*H*4] radiance of ONE sample per pixel (not accumulated); rows [y0,y1)
void probe_render_pt(const ProbeSceneDesc* d, int accum_id, int W, int H, int y0, int y1, int v1_stop, int v2_stop,
                     float* out, int threads) {
    SceneView S = make_scene_view(*d);
    if (threads < 1) threads = 1;
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) {
        pool.emplace_back([=]() {
            for (int y = y0 + t; y < y1; y += threads)
                for (int x = 0; x < W; ++x) {
                    Rng rng = rng_seed(accum_id + 10007, (uint32_t)x, (uint32_t)y, (uint32_t)W);
                    float ox = rng_next(rng), oy = rng_next(rng);
                    float su = ((float)x + ox) / (float)W, sv = ((float)y + oy) / (float)H;
                    V3 o(d->cam_pos[0], d->cam_pos[1], d->cam_pos[2]);
                    V3 dir = normalize(V3(d->cam_d00[0], d->cam_d00[1], d->cam_d00[2]) + su * V3(d->cam_du[0], d->cam_du[1], d->cam_du[2]) +
                                       sv * V3(d->cam_dv[0], d->cam_dv[1], d->cam_dv[2]));
                    Hit h = trace<false>(S.geom, o, dir, 0.f, 1e30f);
                    V3 color(0.f);
                    if (h.prim < 0) {
                        if (S.lights.env.has_env) color = env_radiance(S.lights.env, dir);
                    } else if (v2_stop >= v1_stop) {
                        V3 beta(1.f);
                        Vertex v = vertex_from_hit(S, h, o, dir);
                        if (v1_stop == 0) color = probe_direct(S, v, rng);
                        for (int b = 1; b <= v2_stop; ++b) {
                            V3 no, nd;
                            V3 mul = sample_continuation(S, v, rng, no, nd);
                            beta = beta * mul;
                            Hit nh = trace<false>(S.geom, no, nd, 0.f, 1e30f);
                            if (nh.prim < 0) break;
                            v = vertex_from_hit(S, nh, no, nd);
                            if (b >= v1_stop) color += beta * probe_direct(S, v, rng);
                            float q = fmaxf(0.05f, 1.f - luminance709(beta));
                            float eps = rng_next(rng);
                            if (eps < q) break;
                            beta = beta / (1.f - q);
                        }
                    }
                    if (any_nan(color)) color = V3(0.f);
                    float* px = out + 4 * ((size_t)y * W + x);
                    px[0] = color.x; px[1] = color.y; px[2] = color.z; px[3] = 1.f;
                }
        });
    }
    for (auto& th : pool) th.join();
}

}  // extern "C"

// debug aid: walk the wide tree for references of `prim`, report per level whether the ray hits the child's box
extern "C" void probe_wide_debug(void* p, const float* org, const float* dir, int prim) {
    ProbeScene* s = (ProbeScene*)p;
    GeomView g = make_view(s->geo, s->bvh);
    V3 o(org[0], org[1], org[2]), d(dir[0], dir[1], dir[2]);
    WideRay wr = make_wide_ray(o, d);
    struct E { int ni; int depth; };
    std::vector<E> st; st.push_back({0, 0});
    std::vector<int> parent(g.num_wnodes, -1), pslot(g.num_wnodes, -1);
    while (!st.empty()) {
        E e = st.back(); st.pop_back();
        const F4* n = g.wnodes + 5 * (size_t)e.ni;
        unsigned imask = f_as_u(n[0].w) >> 24, lmask = f_as_u(n[1].z) & 0xff;
        unsigned h = wide_node_hits(n[0], n[2], n[3], n[4], wr, 0.f, 1e30f);
        for (int sl = 0; sl < 8; ++sl) {
            unsigned below = (1u << sl) - 1u;
            if ((imask >> sl) & 1) { int c = f_as_i(n[1].x) + popc_u(imask & below); parent[c] = e.ni; pslot[c] = sl; st.push_back({c, e.depth + 1}); }
            if ((lmask >> sl) & 1) {
                int ref = f_as_i(n[1].y) + popc_u(lmask & below);
                const F4* q = g.wleaf_data + 4 * (size_t)ref;
                int id = q[3].w < 0.f ? f_as_i(q[3].x) : f_as_i(q[0].w);
                if (id == prim) {
                    printf("ref %d of prim %d in node %d slot %d (depth %d) box hit=%d\n", ref, prim, e.ni, sl, e.depth, (h >> sl) & 1);
                    int c = e.ni;
                    while (parent[c] >= 0) {
                        const F4* pn = g.wnodes + 5 * (size_t)parent[c];
                        unsigned ph = wide_node_hits(pn[0], pn[2], pn[3], pn[4], wr, 0.f, 1e30f);
                        printf("   node %d is slot %d of node %d: box hit=%d\n", c, pslot[c], parent[c], (ph >> pslot[c]) & 1);
                        c = parent[c];
                    }
                }
            }
        }
    }
}
