// hm_bvh.h — software acceleration structure over fibre segments + head triangles.
//
// Replaces the reference's OptiX IAS{triangle GAS, curve GAS} + RT-core traversal
// (owl::traceRay call sites cuda/path_tracing.cu:50,
// cuda/hair_msnn.cu:62,240, cuda_headers/optix_common.cuh:189,257,357).  B200 has no
// RT cores, so this is a plain binary BVH laid out for 128-bit loads:
//
//   node = 4 x float4 (64 B, one coalesced 64-byte read per visit)
//     q0 = (lo0.x lo0.y lo0.z hi0.x)
//     q1 = (hi0.y hi0.z lo1.x lo1.y)
//     q2 = (lo1.z hi1.x hi1.y hi1.z)
//     q3 = (child0, child1, -, -) as int bits
//   child >= 0 : inner node index
//   child <  0 : leaf holding ONE primitive reference, ~child = its leaf slot
//
//   A fibre segment enters the tree as several references (one per sub-span of its
//   parameter range, each with the tight box of its piece of the curve — a thin diagonal
//   tube fills a tiny fraction of its own box); all references of a segment name the same
//   slot, so the primitive data exists once.
//
//   leaf slot s : leaf_data[4s..4s+3] = the primitive itself, 64 B, so a leaf test is ONE
//                                 dependent fetch after the node (HBM capacity is cheap
//                                 on B200; an index -> control-point indirection is not):
//                                   fibre   : 4 control points (xyz, w); w of point 1 is the
//                                             radius, w of point 0 carries the primitive id
//                                             (int bits), w of point 3 is >= 0
//                                   triangle: v0, v1, v2, (id bits, -, -, -1)   (w < 0 tags it)
//                 leaf_prim[s]  = primitive id (segment id, or num_segments + triangle id)
//                 leaf_code[s]  = host-side bookkeeping only (first control-point index,
//                                 or triangle index | kTriTag)
//
// The traversal routine is shared by the CUDA kernels and the host build (the
// latter only serves the CPU oracle / baseline); it uses only IEEE + - * / sqrt and
// explicit fmaf so both builds return bit-identical (t, prim, u).
#pragma once
#include "hm_curve.h"

namespace hm {

static constexpr int kTriTag = 0x40000000;
static constexpr int kStackDepth = 64;

struct F4 {
    float x, y, z, w;
};

struct GeomView {
    const F4* nodes;      // 4 per node
    const F4* leaf_data;  // 4 per leaf slot
    const int* leaf_code;
    const int* leaf_prim;
    const F4* cps;        // xyz + radius
    const F4* tri_verts;  // 3 per triangle (w unused)
    int num_segments;
    int num_tris;
    int num_nodes;
};

struct Hit {
    float t;
    int prim;   // -1 = miss
    float u;    // curve parameter, or barycentric of v1
    float v;    // barycentric of v2 (triangles)
};

#if defined(__CUDA_ARCH__)
#define HM_LDG4(p) __ldg(reinterpret_cast<const float4*>(p))
HM_D F4 load_f4(const F4* p) {
    float4 v = HM_LDG4(p);
    F4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w;
    return r;
}
HM_D int load_i(const int* p) { return __ldg(p); }
#else
inline F4 load_f4(const F4* p) { return *p; }
inline int load_i(const int* p) { return *p; }
#endif

HM_HD int f_as_i(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_int(f);
#else
    union { float f; int i; } c; c.f = f; return c.i;
#endif
}

HM_HD V4 f4_to_v4(F4 a) { return V4(a.x, a.y, a.z, a.w); }

// Slab test against [lo,hi] with precomputed 1/d and o/d; returns entry distance or
// a value > tmax on miss.
HM_HD float slab(float lox, float loy, float loz, float hix, float hiy, float hiz,
                 V3 idir, V3 ood, float tmin, float tmax) {
    float x0 = fmaf(lox, idir.x, -ood.x), x1 = fmaf(hix, idir.x, -ood.x);
    float y0 = fmaf(loy, idir.y, -ood.y), y1 = fmaf(hiy, idir.y, -ood.y);
    float z0 = fmaf(loz, idir.z, -ood.z), z1 = fmaf(hiz, idir.z, -ood.z);
    float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), tmin));
    float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), tmax));
    return tn <= tf ? tn : 3.0e38f;
}

struct TraceStats {
    int nodes;
    int prims;
};

// One primitive of a leaf slot against the ray; updates `best` on an accepted hit.  A
// segment is reachable through several references: once it holds the current best hit the
// others are skipped (they would find the same (t, u) again).
HM_HD bool test_slot(const GeomView& g, int slot, V3 o, V3 d, const RayFrame& rf, float tmin, Hit& best) {
    const F4* p = g.leaf_data + 4 * (size_t)slot;
    F4 a = load_f4(p + 0), b = load_f4(p + 1), c = load_f4(p + 2), e = load_f4(p + 3);
    if (e.w < 0.f) {
        float t, b1, b2;
        if (intersect_triangle(o, d, tmin, best.t, V3(a.x, a.y, a.z), V3(b.x, b.y, b.z), V3(c.x, c.y, c.z), t, b1, b2)) {
            best.t = t; best.u = b1; best.v = b2; best.prim = f_as_i(e.x);
            return true;
        }
    } else if (f_as_i(a.w) != best.prim) {
        SegHit sh;
        if (intersect_fibre(rf, tmin, best.t, f4_to_v4(a), f4_to_v4(b), f4_to_v4(c), f4_to_v4(e), sh)) {
            best.t = sh.t; best.u = sh.u; best.v = 0.f; best.prim = f_as_i(a.w);
            return true;
        }
    }
    return false;
}

// ANY = true: occlusion query, returns at the first accepted hit.  Portable one-ray loop
// (host oracle/baseline, and the per-ray statistics hook); the production kernels run the
// warp-cooperative schedule of hm_trace_dev.cuh over the same per-primitive arithmetic.
template <bool ANY>
HM_HD Hit trace(const GeomView& g, V3 o, V3 d, float tmin, float tmax, TraceStats* stats = nullptr) {
    Hit best;
    best.t = tmax; best.prim = -1; best.u = 0.f; best.v = 0.f;
    if (g.num_nodes == 0) return best;

    const float eps = 1e-20f;
    V3 dd = V3(fabsf(d.x) > eps ? d.x : (d.x < 0.f ? -eps : eps),
               fabsf(d.y) > eps ? d.y : (d.y < 0.f ? -eps : eps),
               fabsf(d.z) > eps ? d.z : (d.z < 0.f ? -eps : eps));
    V3 idir = V3(1.f / dd.x, 1.f / dd.y, 1.f / dd.z);
    V3 ood = V3(o.x * idir.x, o.y * idir.y, o.z * idir.z);
    RayFrame rf = make_ray_frame(o, d);

    int stack[kStackDepth];
    int sp = 0;
    int cur = 0;  // root
    const int kDone = 0x7fffffff;

    while (cur != kDone) {
        if (cur >= 0) {
            const F4* n = g.nodes + 4 * (size_t)cur;
            F4 q0 = load_f4(n + 0), q1 = load_f4(n + 1), q2 = load_f4(n + 2), q3 = load_f4(n + 3);
            if (stats) stats->nodes++;
            float t0 = slab(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, idir, ood, tmin, best.t);
            float t1 = slab(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, idir, ood, tmin, best.t);
            int c0 = f_as_i(q3.x), c1 = f_as_i(q3.y);
            bool h0 = t0 < 2.9e38f, h1 = t1 < 2.9e38f;
            if (h0 && h1) {
                if (t1 < t0) { int tmp = c0; c0 = c1; c1 = tmp; }
                stack[sp++] = c1;
                cur = c0;
            } else if (h0) {
                cur = c0;
            } else if (h1) {
                cur = c1;
            } else {
                cur = sp > 0 ? stack[--sp] : kDone;
            }
        } else {
            if (stats) stats->prims++;
            if (test_slot(g, ~cur, o, d, rf, tmin, best) && ANY) return best;
            cur = sp > 0 ? stack[--sp] : kDone;
        }
    }
    return best;
}

}  // namespace hm
