// hm_curve.h — round Catmull-Rom fibre segments: polynomial form, ray intersection,
// and post-hit surface geometry.
//
// Two different contracts live here:
//
//  (1) Post-hit geometry (row a4): given (segment control points, u, t) produce the
//      refined hit point, normal, tangent, centre and radius exactly as the
//      reference derives them from OptiX's hit attributes
//      (cuda_headers/curve_utils.cuh:57-129 cubic interpolator, :159-199 surface
//      normal with projection of the hit point onto the tube, :210-215 tangent,
//      :218-245 computeCurveIntersection).  fibre_hit_geometry() follows that
//      arithmetic operation for operation.
//
//  (2) The ray/segment intersection itself (row a3).  The reference delegates it to
//      OptiX's closed-source OPTIX_PRIMITIVE_TYPE_ROUND_CATMULLROM intersector
//      (extern/owl/owl/CurvesGeomGroup.cpp:162) — there is no source to follow, so
//      this is new arithmetic: a ray-centric iterative tangent-cylinder solver in
//      the spirit of Reshetov & Luebke's phantom intersector (HPG 2018).  It only
//      uses + - * / sqrt and explicit fmaf, so the SAME source compiled for the host
//      (-ffp-contract=off, -mfma) and for sm_100a (-fmad=false) yields bit-identical
//      (t, u) — that is the hit-ID parity contract.
//      Semantics: front-face entry hits of the lateral tube surface for u in [0,1];
//      no end caps (the reference's NRC/HairMSNN setting, render_hair_msnn.cu:346;
//      the path tracer's flat caps, render_path_tracing.cu:223, only differ at the
//      50k strand tips and are not modelled).
#pragma once
#include "hm_math.h"

namespace hm {

// P(u) = ((a*u + b)*u + c)*u + d, w = radius
struct CubicSeg {
    V4 a, b, c, d;

    HM_HD void from_catmull_rom(V4 q0, V4 q1, V4 q2, V4 q3) {
        a = (-1.0f * q0 + (3.0f) * q1 + (-3.0f) * q2 + (1.0f) * q3) / 2.0f;
        b = (2.0f * q0 + (-5.0f) * q1 + (4.0f) * q2 + (-1.0f) * q3) / 2.0f;
        c = (-1.0f * q0 + (1.0f) * q2) / 2.0f;
        d = ((2.0f) * q1) / 2.0f;
    }
    HM_HD V4 pos4(float u) const { return (((a * u) + b) * u + c) * u + d; }
    HM_HD V4 vel4(float u) const {
        if (u == 0) u = 0.000001f;
        if (u == 1) u = 0.999999f;
        return ((3.0f * a * u) + 2.0f * b) * u + c;
    }
    HM_HD V3 acc3(float u) const { return (6.0f * a * u + 2.0f * b).xyz(); }
};

struct FibreHit {
    V3 p;       // hit point dropped onto the tube surface
    V3 n;       // outward normal
    V3 t;       // unit tangent
    V3 centre;  // curve point at u
    float radius;
};

// ray_point = origin + t_hit * direction
HM_HD FibreHit fibre_hit_geometry(const CubicSeg& s, float u, V3 ray_point) {
    FibreHit h;
    V4 p4 = s.pos4(u);
    h.centre = p4.xyz();
    h.radius = p4.w;
    h.p = ray_point;

    V3 normal;
    if (u == 0.0f) {
        normal = -s.vel4(0).xyz();
    } else if (u == 1.0f) {
        normal = s.vel4(1).xyz();
    } else {
        V3 c = p4.xyz();
        float r = p4.w;
        V4 d4 = s.vel4(u);
        V3 d = d4.xyz();
        float dr = d4.w;
        float dd = dot(d, d);

        V3 o1 = h.p - c;
        o1 -= (dot(o1, d) / dd) * d;
        o1 = o1 * (r / length(o1));
        h.p = c + o1;

        dd -= dot(s.acc3(u), o1);
        normal = dd * o1 - (dr * r) * d;
    }
    h.n = normalize(normalize(normal));
    h.t = normalize(s.vel4(u).xyz());
    return h;
}

// ---------------------------------------------------------------------------
// Intersection
// ---------------------------------------------------------------------------

// Orthonormal frame with z = unit ray direction (branch-free Frisvad/Duff variant,
// spelled with explicit fmaf for host/device bit equality).
struct RayFrame {
    V3 o, ex, ey, ez;
};

HM_HD RayFrame make_ray_frame(V3 o, V3 d) {
    RayFrame f;
    f.o = o;
    f.ez = d;
    float sign = d.z >= 0.f ? 1.f : -1.f;
    float a = -1.f / (sign + d.z);
    float b = d.x * d.y * a;
    f.ex = V3(fmaf(sign * d.x * d.x, a, 1.f), sign * b, -sign * d.x);
    f.ey = V3(b, fmaf(d.y * d.y, a, sign), -d.y);
    return f;
}

HM_HD float fdot(V3 a, V3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }

HM_HD V3 to_ray_space(const RayFrame& f, V3 p) {
    V3 q = p - f.o;
    return V3(fdot(q, f.ex), fdot(q, f.ey), fdot(q, f.ez));
}
HM_HD V3 dir_to_ray_space(const RayFrame& f, V3 v) {
    return V3(fdot(v, f.ex), fdot(v, f.ey), fdot(v, f.ez));
}

#ifndef HM_FIBRE_REASON
#define HM_FIBRE_REASON(code)
#endif

struct SegHit {
    float t;   // ray parameter
    float u;   // curve parameter
};

// Catmull-Rom segment in ray space as a cubic with Horner coefficients.
struct RaySpaceCubic {
    V3 a, b, c, d;
    HM_HD V3 pos(float u) const {
        return V3(fmaf(fmaf(fmaf(a.x, u, b.x), u, c.x), u, d.x),
                  fmaf(fmaf(fmaf(a.y, u, b.y), u, c.y), u, d.y),
                  fmaf(fmaf(fmaf(a.z, u, b.z), u, c.z), u, d.z));
    }
    HM_HD V3 vel(float u) const {
        return V3(fmaf(fmaf(3.f * a.x, u, 2.f * b.x), u, c.x),
                  fmaf(fmaf(3.f * a.y, u, 2.f * b.y), u, c.y),
                  fmaf(fmaf(3.f * a.z, u, 2.f * b.z), u, c.z));
    }
};

// Ray vs. the swept tube of one Catmull-Rom span, control points q0..q3 (xyz + radius in w;
// radius taken at q1, constant along the span — the .hair pipeline produces one width per
// file, scene.cpp:52-57).  Split in two so that the GPU traversal can run the cheap part for
// every candidate a leaf offers and batch the expensive part across the lanes of a warp:
//
//   fibre_candidate : conservative rejects
//       1. depth range of the Bezier hull vs (tmin, tmax);
//       2. "fat line" in the ray's projection plane: the projected curve lies in the convex
//          hull of its Bezier points, i.e. inside a slab around the projected chord; the ray
//          axis (the origin of that plane) must lie within the slab grown by r, and within
//          the hull's extent along the chord.
//       Also yields the Newton start: the chord point nearest to the ray axis.
//   fibre_solve     : Newton iteration on the curve parameter — ray vs. the tangent cylinder
//       at u, step u by the axial offset of the hit.  Accepts hits with t in (tmin, tmax).
//
// intersect_fibre = candidate && solve.  Only + - * / sqrt and explicit fmaf are used.
struct FibreCandidate {
    V3 k0, k1, k2, k3;   // control points in ray space
    float r, u_start;
};

HM_HD bool fibre_candidate(const RayFrame& rf, float tmin, float tmax, V4 q0, V4 q1, V4 q2, V4 q3, FibreCandidate& c) {
    c.k0 = to_ray_space(rf, q0.xyz());
    c.k1 = to_ray_space(rf, q1.xyz());
    c.k2 = to_ray_space(rf, q2.xyz());
    c.k3 = to_ray_space(rf, q3.xyz());
    const V3 k0 = c.k0, k1 = c.k1, k2 = c.k2, k3 = c.k3;
    const float r = q1.w;
    c.r = r;

    // Bezier hull of the span: b0=k1, b1=k1+(k2-k0)/6, b2=k2-(k3-k1)/6, b3=k2
    const float sixth = 1.f / 6.f;
    V3 b1 = V3(fmaf(k2.x - k0.x, sixth, k1.x), fmaf(k2.y - k0.y, sixth, k1.y), fmaf(k2.z - k0.z, sixth, k1.z));
    V3 b2 = V3(fmaf(k1.x - k3.x, sixth, k2.x), fmaf(k1.y - k3.y, sixth, k2.y), fmaf(k1.z - k3.z, sixth, k2.z));

    float zmin = fminf(fminf(k1.z, b1.z), fminf(b2.z, k2.z));
    float zmax = fmaxf(fmaxf(k1.z, b1.z), fmaxf(b2.z, k2.z));
    if (zmin - r > tmax || zmax + r < tmin) { HM_FIBRE_REASON(1); return false; }

    // projected chord e = k2 - k1; offsets of the axis and of the inner hull points from the
    // chord line, and their positions along it (both scaled by |e|)
    const float ex = k2.x - k1.x, ey = k2.y - k1.y;
    const float len2 = fmaf(ex, ex, ey * ey);
    const float elen = sqrtf(len2);
    const float re = fmaf(r, elen, 1e-6f * elen + 1e-12f) * 1.0001f;
    const float c0 = fmaf(ey, k1.x, -(ex * k1.y));
    const float c1 = fmaf(ex, b1.y - k1.y, -(ey * (b1.x - k1.x)));
    const float c2 = fmaf(ex, b2.y - k1.y, -(ey * (b2.x - k1.x)));
    if (c0 > fmaxf(fmaxf(c1, c2), 0.f) + re || c0 < fminf(fminf(c1, c2), 0.f) - re) { HM_FIBRE_REASON(2); return false; }
    const float s0 = -fmaf(ex, k1.x, ey * k1.y);
    const float s1 = fmaf(ex, b1.x - k1.x, ey * (b1.y - k1.y));
    const float s2 = fmaf(ex, b2.x - k1.x, ey * (b2.y - k1.y));
    if (s0 > fmaxf(fmaxf(s1, s2), len2) + re || s0 < fminf(fminf(s1, s2), 0.f) - re) { HM_FIBRE_REASON(3); return false; }

    // start at the chord point nearest to the ray axis (mid-span when seen end-on)
    c.u_start = len2 > 1e-12f ? fminf(fmaxf(s0 / len2, 0.f), 1.f) : 0.5f;
    return true;
}

HM_HD bool fibre_solve(const FibreCandidate& c, float tmin, float tmax, SegHit& hit) {
    const V3 k0 = c.k0, k1 = c.k1, k2 = c.k2, k3 = c.k3;
    RaySpaceCubic cu;
    cu.a = V3(0.5f * (-k0.x + 3.f * k1.x - 3.f * k2.x + k3.x),
              0.5f * (-k0.y + 3.f * k1.y - 3.f * k2.y + k3.y),
              0.5f * (-k0.z + 3.f * k1.z - 3.f * k2.z + k3.z));
    cu.b = V3(0.5f * (2.f * k0.x - 5.f * k1.x + 4.f * k2.x - k3.x),
              0.5f * (2.f * k0.y - 5.f * k1.y + 4.f * k2.y - k3.y),
              0.5f * (2.f * k0.z - 5.f * k1.z + 4.f * k2.z - k3.z));
    cu.c = V3(0.5f * (k2.x - k0.x), 0.5f * (k2.y - k0.y), 0.5f * (k2.z - k0.z));
    cu.d = k1;

    const float r2 = c.r * c.r;
    const float kConv = 5e-5f;
    float u = c.u_start;
    float uold = 0.f, dt1 = 0.f, dt2 = 0.f;
    for (int it = 0; it < 16; ++it) {
        V3 c0p = cu.pos(u);
        V3 cd = cu.vel(u);
        // ray (0,0,s) vs infinite cylinder through c0p along cd, radius r
        float cxy = fmaf(cd.y, cd.y, cd.x * cd.x);
        float dp = fmaf(c0p.y, c0p.y, c0p.x * c0p.x);
        float cdd = fmaf(c0p.y, cd.y, c0p.x * cd.x);
        float cxd = fmaf(c0p.x, cd.y, -(c0p.y * cd.x));
        float cz2 = cd.z * cd.z;
        float dd = cxy + cz2;
        float bq = -(cd.z * cdd);
        float aq = fmaf(cxd, cxd, fmaf(dp, cz2, -(dd * r2)));
        float det = fmaf(bq, bq, -(aq * cxy));
        bool real_hit = det > 0.f;
        // guard against a ray (numerically) parallel to the tangent
        float cq = fmaxf(cxy, 1e-12f * dd);
        float s = (bq - (real_hit ? sqrtf(det) : 0.f)) / cq;
        float dt = fmaf(s, cd.z, -cdd) / dd;

        if (fabsf(dt) < kConv) {
            if (!real_hit) { HM_FIBRE_REASON(4); return false; }
            float t = s + c0p.z;
            if (!(t > tmin && t < tmax)) { HM_FIBRE_REASON(5); return false; }
            hit.t = t;
            hit.u = u;
            return true;
        }
        dt = fminf(dt, 0.5f);
        dt = fmaxf(dt, -0.5f);
        dt1 = dt2;
        dt2 = dt;
        float unext;
        if (it > 0 && dt1 * dt2 < 0.f) {
            // bracketed: regula falsi with a periodic bisection safeguard
            if ((it & 3) == 0) unext = 0.5f * (uold + u);
            else unext = (dt2 * uold - dt1 * u) / (dt2 - dt1);
        } else {
            unext = u + dt;
        }
        uold = u;
        u = unext;
        // A step may overshoot the span (rays nearly parallel to the fibre): retry once from
        // the end it left through.  Leaving through the same end again means the surface
        // point nearest to this ray lies beyond the span — the neighbouring segment of the
        // strand owns it (no end caps).
        if (u < 0.f) {
            if (uold == 0.f) { HM_FIBRE_REASON(6); return false; }
            u = 0.f;
        } else if (u > 1.f) {
            if (uold == 1.f) { HM_FIBRE_REASON(6); return false; }
            u = 1.f;
        }
    }
    HM_FIBRE_REASON(7);
    return false;
}

HM_HD bool intersect_fibre(const RayFrame& rf, float tmin, float tmax,
                           V4 q0, V4 q1, V4 q2, V4 q3, SegHit& hit) {
    FibreCandidate c;
    if (!fibre_candidate(rf, tmin, tmax, q0, q1, q2, q3, c)) return false;
    return fibre_solve(c, tmin, tmax, hit);
}

// Watertight-enough Moeller-Trumbore for the head mesh (closed-source in the
// reference: OptiX built-in triangles).  Returns barycentrics (b1, b2) of v1, v2.
HM_HD bool intersect_triangle(V3 o, V3 d, float tmin, float tmax, V3 v0, V3 v1, V3 v2,
                              float& t, float& b1, float& b2) {
    V3 e1 = v1 - v0, e2 = v2 - v0;
    V3 pv = cross(d, e2);
    float det = fdot(e1, pv);
    if (det == 0.f) return false;
    float inv = 1.f / det;
    V3 tv = o - v0;
    float u = fdot(tv, pv) * inv;
    if (u < 0.f || u > 1.f) return false;
    V3 qv = cross(tv, e1);
    float v = fdot(d, qv) * inv;
    if (v < 0.f || u + v > 1.f) return false;
    float tt = fdot(e2, qv) * inv;
    if (!(tt > tmin && tt < tmax)) return false;
    t = tt; b1 = u; b2 = v;
    return true;
}

}  // namespace hm
