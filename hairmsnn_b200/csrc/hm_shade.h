// hm_shade.h — one path vertex: hit record -> shading frame, direct lighting with MIS,
// BSDF continuation.  Shared by the sm_100a wavefront kernels (hm_wavefront.cu) and
// the host probe used by the CPU test-suite.
//
// Behavioural contract:
//   vertex_from_hit      == hairCH / triangleMeshCH      (cuda_headers/optix_common.cuh:527-599)
//   sample_direct        == directLighting = sampleLights + sampleBSDF with MIS
//                                                        (optix_common.cuh:151-309)
//   sample_continuation  == nextPathVertex / msnnNextPathVertex up to (not including)
//                           the radiance trace           (optix_common.cuh:311-360,
//                                                         cuda/hair_msnn.cu:17-65)
//
// The reference runs these inside one OptiX megakernel and traces shadow/radiance rays
// inline.  Here nothing in a vertex's sampling depends on a trace result, so a vertex
// emits up to two occlusion rays and one continuation ray in a single pass and the
// results are folded in afterwards (wavefront form).  The random-number draw order per
// path is the reference's:  light pick (1) -> env sample (2) -> [BSDF sample (2 + 2 for
// fibres) if the light is not a delta light]  -> Russian roulette (1, bounces >= 1)
// -> continuation (2 + 2 for fibres).
#pragma once
#include "hm_bsdf.h"
#include "hm_bvh.h"
#include "hm_light.h"
#include "hm_rng.h"

namespace hm {

struct SceneView {
    GeomView geom;
    const int* seg_cp;       // segment id -> first control point
    const F4* tri_normals;   // 3 per triangle
    LightSet lights;
    HairLobes lobes;
    float kd[3];
    float surf_alpha;        // raw material alpha; squared after clamping at the hit
    float scene_scale;
    int mis;
};

struct Vertex {
    V3 p, n, t, wo;
    V3 wo_local;
    M3 to_local;     // rows X, Y, Z
    float radius;    // fibres only
    float h;         // fibres: dot(Y, n)
    float alpha;     // surfaces: squared roughness
    bool surface;
    hairdetail::FibreGeom geom;   // fibres: the wo/h-dependent part of the scattering model, shared by
                                  // every evaluation at this vertex (same arithmetic, computed once)
};

// orthonormalBasis (cuda_headers/utils.cuh:232-259) — note the double-precision
// literals: 1./(1.+n.z) and 1.-x are evaluated in double and rounded once.
HM_HD M3 surface_frame(V3 n) {
    V3 c1, c2;
    if (n.z < -0.999999f) {
        c1 = V3(0.f, -1.f, 0.f);
        c2 = V3(-1.f, 0.f, 0.f);
    } else {
        float a = (float)(1. / (1. + (double)n.z));
        float b = -n.x * n.y * a;
        c1 = normalize(V3((float)(1. - (double)(n.x * n.x * a)), b, -n.x));
        c2 = normalize(V3(b, (float)(1. - (double)(n.y * n.y * a)), -n.y));
    }
    M3 m;
    m.r0 = c1; m.r1 = c2; m.r2 = n;
    return m;
}

// ray_o + t * ray_d is formed exactly as getHitPoint() does (curve_utils.cuh:133-140).
HM_HD Vertex vertex_from_hit(const SceneView& S, const Hit& hit, V3 ray_o, V3 ray_d) {
    Vertex v;
    v.wo = -1.f * ray_d;
    const int ns = S.geom.num_segments;
    if (hit.prim < ns) {
        const F4* cp = S.geom.cps + load_i(S.seg_cp + hit.prim);
        CubicSeg seg;
        seg.from_catmull_rom(f4_to_v4(load_f4(cp + 0)), f4_to_v4(load_f4(cp + 1)),
                             f4_to_v4(load_f4(cp + 2)), f4_to_v4(load_f4(cp + 3)));
        FibreHit fh = fibre_hit_geometry(seg, hit.u, ray_o + hit.t * ray_d);
        v.p = fh.p; v.n = fh.n; v.t = fh.t; v.radius = fh.radius;
        V3 X = v.t;
        V3 Y = normalize(cross(v.wo, X));
        V3 Z = normalize(cross(X, Y));
        v.to_local.r0 = X; v.to_local.r1 = Y; v.to_local.r2 = Z;
        v.wo_local = normalize(v.to_local.apply(v.wo));
        v.h = dot(Y, v.n);
        v.alpha = 0.f;
        v.surface = false;
        hairdetail::fibre_geom(S.lobes, v.wo_local, v.h, v.geom);
    } else {
        const int ti = hit.prim - ns;
        const float bu = hit.u, bv = hit.v;
        const float bw = 1.f - bu - bv;
        F4 a = load_f4(S.geom.tri_verts + 3 * (size_t)ti + 0);
        F4 b = load_f4(S.geom.tri_verts + 3 * (size_t)ti + 1);
        F4 c = load_f4(S.geom.tri_verts + 3 * (size_t)ti + 2);
        v.p = bw * V3(a.x, a.y, a.z) + bu * V3(b.x, b.y, b.z) + bv * V3(c.x, c.y, c.z);
        F4 na = load_f4(S.tri_normals + 3 * (size_t)ti + 0);
        F4 nb = load_f4(S.tri_normals + 3 * (size_t)ti + 1);
        F4 nc = load_f4(S.tri_normals + 3 * (size_t)ti + 2);
        v.n = normalize(bw * V3(na.x, na.y, na.z) + bu * V3(nb.x, nb.y, nb.z) + bv * V3(nc.x, nc.y, nc.z));
        v.to_local = surface_frame(v.n);
        v.t = v.to_local.r0;
        v.wo_local = normalize(v.to_local.apply(v.wo));
        float al = fminf(1.f, fmaxf(0.01f, S.surf_alpha));
        v.alpha = al * al;
        v.radius = 0.f;
        v.h = 0.f;
        v.surface = true;
    }
    return v;
}

// An occlusion query paired with the radiance it unlocks when unoccluded.
struct Probe {
    V3 o, d;
    V3 value;
    bool active;
};

struct DirectSample {
    Probe light, bsdf;
};

HM_HD V3 eval_bsdf(const SceneView& S, const Vertex& v, V3 wi_local, float* pdf) {
    if (!v.surface) return hair_eval_geom(S.lobes, v.geom, wi_local, pdf);
    V3 f = surf_eval(v.wo_local, wi_local, V3(S.kd[0], S.kd[1], S.kd[2]), v.alpha);
    *pdf = surf_pdf(v.alpha, v.wo_local, normalize(v.wo_local + wi_local));
    return f;
}

// Draws wi for the vertex's BSDF; returns f*cos, pdf, world + local directions.
HM_HD V3 sample_bsdf_dir(const SceneView& S, const Vertex& v, Rng& rng, V3& wi, float& pdf) {
    float r0 = rng_next(rng);
    float r1 = rng_next(rng);
    if (!v.surface) {
        float r2 = rng_next(rng);
        float r3 = rng_next(rng);
        V3 wl = hair_sample_dir_geom(S.lobes, v.geom, r0, r1, r2, r3);
        wi = normalize(v.to_local.transposed().apply(wl));
        return hair_eval_geom(S.lobes, v.geom, wl, &pdf);
    }
    V3 wl = surf_sample(r0, r1, v.alpha, v.wo_local, &pdf);
    wi = normalize(v.to_local.transposed().apply(wl));
    return surf_eval(v.wo_local, wl, V3(S.kd[0], S.kd[1], S.kd[2]), v.alpha);
}

// Offsets the spawn point as the reference does: surfaces never spawn into the lower
// hemisphere (returns false); fibres hop to the far side of the tube.
HM_HD bool spawn_point(const Vertex& v, V3 wi, V3& p, V3& origin) {
    bool lower = dot(wi, v.n) < 0.f;
    V3 nd = v.n;
    if (v.surface && lower) return false;
    if (!v.surface && lower) {
        nd = -v.n;
        p = p + 2.f * v.radius * nd;
    }
    origin = p + 1e-3f * nd;
    return true;
}

HM_HD void sample_direct(const SceneView& S, const Vertex& v, Rng& rng, DirectSample& out) {
    out.light.active = false; out.bsdf.active = false;
    out.light.value = V3(0.f); out.bsdf.value = V3(0.f);
    const LightSet& L = S.lights;
    V3 p = v.p;   // the light pass's offset carries into the BSDF pass (si is shared by value)

    // --- light sampling ---
    bool is_delta = false;
    {
        float r = rng_next(rng);
        int sel = (int)floorf(r * L.num_total);
        float light_pdf = 1.f;
        V3 emit(0.f), wi(0.f);
        if (sel >= L.num_dlights) {
            light_pdf = light_pdf * 1.f / L.num_total;
            float env_p = 1.f;
            float u0 = rng_next(rng);
            float u1 = rng_next(rng);
            emit = env_sample(L.env, u0, u1, wi, env_p);
            light_pdf = light_pdf * env_p;
        } else {
            light_pdf = light_pdf * 1.f / L.num_total;
            emit = V3(L.dl_emit[sel][0], L.dl_emit[sel][1], L.dl_emit[sel][2]);
            wi = normalize(V3(L.dl_from[sel][0], L.dl_from[sel][1], L.dl_from[sel][2]));
            is_delta = true;
        }
        V3 wi_local = normalize(v.to_local.apply(wi));
        V3 origin;
        if (spawn_point(v, wi, p, origin)) {
            float bsdf_pdf = 1.f;
            V3 f = eval_bsdf(S, v, wi_local, &bsdf_pdf);
            if (S.mis && !is_delta)
                out.light.value = f * emit * power_heuristic(light_pdf, bsdf_pdf) / light_pdf;
            else
                out.light.value = f * emit / light_pdf;
            out.light.o = origin; out.light.d = wi; out.light.active = true;
        }
    }
    if (!S.mis || is_delta) return;

    // --- BSDF sampling (environment only) ---
    {
        V3 wi; float bsdf_pdf = 1.f;
        V3 f = sample_bsdf_dir(S, v, rng, wi, bsdf_pdf);
        V3 origin;
        if (spawn_point(v, wi, p, origin)) {
            V3 emit = env_radiance(L.env, wi);
            float light_pdf = env_pdf(L.env, wi) * 1.f / L.num_total;
            out.bsdf.value = f * emit * power_heuristic(bsdf_pdf, light_pdf) / bsdf_pdf;
            out.bsdf.o = origin; out.bsdf.d = wi; out.bsdf.active = true;
        }
    }
}

// Folds the two occlusion results into the vertex's direct radiance
// (directLighting's NaN scrub applies to the sum).
HM_HD V3 resolve_direct(V3 light_value, bool light_visible, V3 bsdf_value, bool bsdf_visible) {
    V3 s = (light_visible ? light_value : V3(0.f)) + (bsdf_visible ? bsdf_value : V3(0.f));
    if (any_nan(s)) s = V3(0.f);
    return s;
}

// Continuation ray and throughput factor.  Always produces a ray (the reference traces
// even when a surface sample points below the horizon).
// pdf_out (optional) receives the sampling density (nextPathVertexNRC's return value,
// cuda/nrc.cu:18-67).
HM_HD V3 sample_continuation(const SceneView& S, const Vertex& v, Rng& rng, V3& ray_o, V3& ray_d, float* pdf_out = nullptr) {
    V3 wi; float pdf = 1.f;
    V3 f = sample_bsdf_dir(S, v, rng, wi, pdf);
    if (pdf_out) *pdf_out = pdf;
    V3 mul = (pdf == 0.f) ? f : f / pdf;
    if (any_nan(mul)) mul = V3(1.f);
    V3 wo_next = -wi;
    bool lower = dot(wi, v.n) < 0.f;
    V3 nd = v.n;
    V3 p = v.p;
    if (lower && !v.surface) {
        nd = -v.n;
        p = p + 2.f * v.radius * nd;
    }
    ray_o = p + 1e-3f * nd;
    ray_d = -wo_next;
    return mul;
}

}  // namespace hm
