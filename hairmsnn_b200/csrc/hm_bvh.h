// hm_bvh.h — software acceleration structure over fibre segments + head triangles.
//
// Replaces the reference's OptiX IAS{triangle GAS, curve GAS} + RT-core traversal
// (owl::traceRay call sites cuda/path_tracing.cu:50,
// cuda/hair_msnn.cu:62,240, cuda_headers/optix_common.cuh:189,257,357).  B200 has no
// RT cores.  Two trees over the same primitive references live here:
//   * the binary SAH tree described next — what the builder produces first, what the host oracle /
//     CPU baseline traverses (trace<>), and the input of the collapse;
//   * the 8-wide quantised tree derived from it (second half of this file: wide_node_hits,
//     trace_wide<>) — what every CUDA kernel traverses (hm_trace_dev.cuh).
// Both return bit-identical closest hits (same references, same primitive tests).
//
// Binary tree, laid out for 128-bit loads:
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
// The primitive tests and both traversal routines compile for host and device; they use only
// IEEE + - * / sqrt and explicit fmaf so both builds return bit-identical (t, prim, u).
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
    // 8-wide quantised tree (the structure the sm_100a kernels traverse; see WideNode below)
    const F4* wnodes;     // 5 per wide node (80 B)
    const F4* wleaf_data; // 4 per wide leaf reference (64 B, same content as a leaf slot)
    int num_wnodes;
    unsigned k47;         // 0x47000000 as a RUN-TIME value (see byte_biased_t): set by whoever fills the view
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
// leaf primitives are touched once per test and there are 2.8 GB of them: HM_LEAF_LOAD picks the cache
// operator (0 = ld.global.nc like the nodes, 1 = .cs streaming, 2 = .L2::evict_first) so that they displace
// fewer node lines
#ifndef HM_LEAF_LOAD
#define HM_LEAF_LOAD 0
#endif
HM_D F4 load_f4_leaf(const F4* p) {
#if HM_LEAF_LOAD == 1
    float4 v = __ldcs(reinterpret_cast<const float4*>(p));
#else
    float4 v = HM_LDG4(p);
#endif
    F4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w;
    return r;
}
// the whole 64-byte primitive; HM_LEAF_LOAD == 2: two 256-bit loads marked evict-first in L2 (sm_100 accepts
// the L2 eviction hint on the 256-bit forms only)
HM_D void load_leaf64(const F4* p, F4& a, F4& b, F4& c, F4& e) {
#if HM_LEAF_LOAD == 2
    unsigned r[16];
    asm volatile("ld.global.L2::evict_first.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
    asm volatile("ld.global.L2::evict_first.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "l"(p + 2));
    a.x = __uint_as_float(r[0]); a.y = __uint_as_float(r[1]); a.z = __uint_as_float(r[2]); a.w = __uint_as_float(r[3]);
    b.x = __uint_as_float(r[4]); b.y = __uint_as_float(r[5]); b.z = __uint_as_float(r[6]); b.w = __uint_as_float(r[7]);
    c.x = __uint_as_float(r[8]); c.y = __uint_as_float(r[9]); c.z = __uint_as_float(r[10]); c.w = __uint_as_float(r[11]);
    e.x = __uint_as_float(r[12]); e.y = __uint_as_float(r[13]); e.z = __uint_as_float(r[14]); e.w = __uint_as_float(r[15]);
#else
    a = load_f4_leaf(p + 0); b = load_f4_leaf(p + 1); c = load_f4_leaf(p + 2); e = load_f4_leaf(p + 3);
#endif
}
#else
inline F4 load_f4(const F4* p) { return *p; }
inline int load_i(const int* p) { return *p; }
inline F4 load_f4_leaf(const F4* p) { return *p; }
inline void load_leaf64(const F4* p, F4& a, F4& b, F4& c, F4& e) { a = p[0]; b = p[1]; c = p[2]; e = p[3]; }
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

// ---------------------------------------------------------------------------------
// 8-wide tree with boxes quantised to 8 bits against the node's own bounds (after Ylitie,
// Karras, Laine 2017, "Efficient incoherent ray traversal on GPUs through compressed wide
// BVHs" — their idea, this layout and code).  Why: the binary tree above is 215 MB of nodes
// for the curly scene and misses the 126 MB L2 on nearly every warp-wide step; this one is
// ~5x smaller (80 B per ~6 children), stays L2-resident, and a ray visits ~3x fewer nodes.
//
//   wide node = 5 x 16 B
//     w0 : origin.x, origin.y, origin.z, (ex | ey << 8 | ez << 16 | imask << 24)
//          ex/ey/ez = IEEE-biased exponents: cell size on an axis = 2^(e-127)
//          imask    = bit i set: child slot i is an inner node
//     w1 : child_base, leaf_base, lmask (bit i: child slot i is a leaf), unused
//          inner child of slot i = child_base + popc(imask & ((1 << i) - 1))
//          leaf reference of slot i = leaf_base + popc(lmask & ((1 << i) - 1))
//     w2 : qlo_x[8], qlo_y[8]      one byte per child slot; box = origin + q * cell
//     w3 : qlo_z[8], qhi_x[8]
//     w4 : qhi_y[8], qhi_z[8]
//   Children sit in the slot whose bits name their octant relative to the node centre
//   (bit k set = towards +axis k), so visiting slots in the order of (slot ^ octinv),
//   highest first, walks them roughly front to back for any ray direction without sorting
//   (octinv: bit k set when the ray travels towards +axis k).
//   Every leaf reference owns a 64-byte copy of its primitive (wleaf_data): the references of
//   one node are contiguous, so a leaf test is one dependent fetch with no index array.
// The wide tree is derived from the binary one (hm_bvh_build.cpp: collapse), holds the same
// references with the same (or looser, by < 1 cell) boxes, and the primitive tests are
// shared — both trees return bit-identical closest hits.
static constexpr int kWideStack = 32;

HM_HD unsigned f_as_u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    union { float f; unsigned u; } c; c.f = f; return c.u;
#endif
}
HM_HD float u_as_f(unsigned u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    union { float f; unsigned u; } c; c.u = u; return c.f;
#endif
}
// 32768 + byte k of w, as a float (exact).  On the device one PRMT drops the byte into mantissa
// bits 8..15 of 2^15 (LSB weight 1 there); the bias is folded into the constant term of the
// slab FMA, so a quantised plane costs PRMT + FFMA.  Folding rounds the constant once more:
// at most 1/512 of a cell, which the builder's 1/64-cell margin covers.
static constexpr float kQBias = 32768.f;
HM_HD float byte_biased(unsigned w, int k) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(__byte_perm(w, 0x47000000u, 0x7504 | (k << 4)));
#else
    return kQBias + (float)((w >> (8 * k)) & 0xffu);
#endif
}
// Same value with the byte index as a template constant and the 0x47000000 operand in a REGISTER (k47, a runtime value
// the caller reads from GeomView): ptxas then encodes the selector as PRMT's immediate.  With the literal __byte_perm
// above it makes 0x47000000 the immediate and re-materialises the selector in a register before every PRMT (the
// destination overwrites it): 48 extra MOVs per node test, a tenth of the traversal's instructions
// (profiles/r2b_k_trace_lines.txt: 80.8 instructions per warp-visit on the byte_biased line instead of 48).
template <int K>
HM_HD float byte_biased_t(unsigned w, unsigned k47) {
#if defined(__CUDA_ARCH__)
    unsigned r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(w), "r"(k47), "n"(0x7504 | (K << 4)));
    return __uint_as_float(r);
#else
    (void)k47;
    return kQBias + (float)((w >> (8 * K)) & 0xffu);
#endif
}
HM_HD int popc_u(unsigned v) {
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}
HM_HD int top_bit(unsigned v) {   // index of the highest set bit, v != 0
#if defined(__CUDA_ARCH__)
    return 31 - __clz((int)v);
#else
    return 31 - __builtin_clz(v);
#endif
}
// bit i of m -> bit (i ^ x), x in [0,7], m 8 bits wide
HM_HD unsigned xor_permute8(unsigned m, int x) {
    if (x & 1) m = ((m & 0x55u) << 1) | ((m & 0xAAu) >> 1);
    if (x & 2) m = ((m & 0x33u) << 2) | ((m & 0xCCu) >> 2);
    if (x & 4) m = ((m & 0x0Fu) << 4) | ((m & 0xF0u) >> 4);
    return m;
}

struct WideRay {
    V3 idir, ood;       // 1/d and o/d
    int octinv;         // bit k: d[k] >= 0
};
HM_HD WideRay make_wide_ray(V3 o, V3 d) {
    const float eps = 1e-20f;
    V3 dd = V3(fabsf(d.x) > eps ? d.x : (d.x < 0.f ? -eps : eps),
               fabsf(d.y) > eps ? d.y : (d.y < 0.f ? -eps : eps),
               fabsf(d.z) > eps ? d.z : (d.z < 0.f ? -eps : eps));
    WideRay r;
    r.idir = V3(1.f / dd.x, 1.f / dd.y, 1.f / dd.z);
    r.ood = V3(o.x * r.idir.x, o.y * r.idir.y, o.z * r.idir.z);
    r.octinv = (dd.x >= 0.f ? 1 : 0) | (dd.y >= 0.f ? 2 : 0) | (dd.z >= 0.f ? 4 : 0);
    return r;
}

// Slab test of the 8 child boxes of one wide node; returns the slots hit (bit i = slot i).
// near_key (optional): (bits of the smallest entry distance among the hit children, low 3 bits replaced by
// that child's slot), 0xffffffff when nothing is hit — entry distances are >= tmin >= 0, so their bit patterns
// order like unsigned integers.
HM_HD unsigned wide_node_hits(F4 w0, F4 w2, F4 w3, F4 w4, const WideRay& r, float tmin, float tmax, unsigned* near_key = nullptr,
                              unsigned k47 = 0x47000000u) {
    const unsigned em = f_as_u(w0.w);
    // t(q) = (q + bias) * (cell * idir) + (origin * idir - o * idir - bias * cell * idir)
    const float ax = u_as_f((em & 0xffu) << 23) * r.idir.x, bx = fmaf(-kQBias, ax, fmaf(w0.x, r.idir.x, -r.ood.x));
    const float ay = u_as_f(((em >> 8) & 0xffu) << 23) * r.idir.y, by = fmaf(-kQBias, ay, fmaf(w0.y, r.idir.y, -r.ood.y));
    const float az = u_as_f(((em >> 16) & 0xffu) << 23) * r.idir.z, bz = fmaf(-kQBias, az, fmaf(w0.z, r.idir.z, -r.ood.z));
    // near / far planes per axis by ray direction
    const bool px = (r.octinv & 1) != 0, py = (r.octinv & 2) != 0, pz = (r.octinv & 4) != 0;
    const unsigned lox[2] = {f_as_u(w2.x), f_as_u(w2.y)}, loy[2] = {f_as_u(w2.z), f_as_u(w2.w)};
    const unsigned loz[2] = {f_as_u(w3.x), f_as_u(w3.y)}, hix[2] = {f_as_u(w3.z), f_as_u(w3.w)};
    const unsigned hiy[2] = {f_as_u(w4.x), f_as_u(w4.y)}, hiz[2] = {f_as_u(w4.z), f_as_u(w4.w)};
    unsigned hits = 0;
    unsigned nearest = 0xffffffffu;
#define HM_WIDE_CHILD(K)                                                                                                  \
    {                                                                                                                     \
        const float tnx = fmaf(byte_biased_t<K>(nx, k47), ax, bx), tfx = fmaf(byte_biased_t<K>(fx, k47), ax, bx);         \
        const float tny = fmaf(byte_biased_t<K>(ny, k47), ay, by), tfy = fmaf(byte_biased_t<K>(fy, k47), ay, by);         \
        const float tnz = fmaf(byte_biased_t<K>(nz, k47), az, bz), tfz = fmaf(byte_biased_t<K>(fz, k47), az, bz);         \
        const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, tmin));                                                        \
        const float tf = fminf(fminf(tfx, tfy), fminf(tfz, tmax));                                                        \
        if (tn <= tf) {                                                                                                   \
            hits |= 1u << (4 * h + K);                                                                                    \
            const unsigned key = (f_as_u(tn) & ~7u) | (unsigned)(4 * h + K);                                              \
            nearest = key < nearest ? key : nearest;                                                                      \
        }                                                                                                                 \
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int h = 0; h < 2; ++h) {
        const unsigned nx = px ? lox[h] : hix[h], fx = px ? hix[h] : lox[h];
        const unsigned ny = py ? loy[h] : hiy[h], fy = py ? hiy[h] : loy[h];
        const unsigned nz = pz ? loz[h] : hiz[h], fz = pz ? hiz[h] : loz[h];
        HM_WIDE_CHILD(0) HM_WIDE_CHILD(1) HM_WIDE_CHILD(2) HM_WIDE_CHILD(3)
    }
#undef HM_WIDE_CHILD
    if (near_key) *near_key = nearest;
    return hits;
}

// Primitive of one wide leaf reference against the ray (same tests as test_slot).
HM_HD bool test_wide_leaf(const GeomView& g, int ref, V3 o, V3 d, const RayFrame& rf, float tmin, Hit& best) {
    const F4* p = g.wleaf_data + 4 * (size_t)ref;
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

// Portable one-ray traversal of the wide tree (host tests, per-ray statistics hook); the
// production kernels run the warp-cooperative schedule of hm_trace_dev.cuh over the same
// node and primitive arithmetic.
template <bool ANY>
HM_HD Hit trace_wide(const GeomView& g, V3 o, V3 d, float tmin, float tmax, TraceStats* stats = nullptr) {
    Hit best;
    best.t = tmax; best.prim = -1; best.u = 0.f; best.v = 0.f;
    if (g.num_wnodes == 0) return best;
    const WideRay wr = make_wide_ray(o, d);
    const RayFrame rf = make_ray_frame(o, d);

    // a "group" = the not-yet-visited hit children of one node: base indices + masks
    struct Group { int child_base, leaf_base; unsigned bits; };   // bits = imask | lmask << 8 | hits(permuted) << 16
    Group stack[kWideStack];
    int sp = 0;
    Group cur;
    cur.child_base = 0; cur.leaf_base = 0; cur.bits = 1u | (1u << (16 + (0 ^ wr.octinv)));   // pseudo-node whose slot 0 is the root

    while (true) {
        unsigned hits = cur.bits >> 16;
        if (hits == 0) {
            if (sp == 0) break;
            cur = stack[--sp];
            continue;
        }
        const int b = top_bit(hits);
        cur.bits &= ~(1u << (16 + b));
        const int slot = b ^ wr.octinv;
        const unsigned below = (1u << slot) - 1u;
        if ((cur.bits >> slot) & 1u) {
            const int ni = cur.child_base + popc_u(cur.bits & 0xffu & below);
            const F4* n = g.wnodes + 5 * (size_t)ni;
            F4 w0 = load_f4(n + 0), w1 = load_f4(n + 1), w2 = load_f4(n + 2), w3 = load_f4(n + 3), w4 = load_f4(n + 4);
            if (stats) stats->nodes++;
            const unsigned imask = f_as_u(w0.w) >> 24, lmask = f_as_u(w1.z) & 0xffu;
            unsigned h = wide_node_hits(w0, w2, w3, w4, wr, tmin, best.t) & (imask | lmask);
            if (h) {
                if (cur.bits >> 16) stack[sp++] = cur;
                cur.child_base = f_as_i(w1.x); cur.leaf_base = f_as_i(w1.y);
                cur.bits = imask | (lmask << 8) | (xor_permute8(h, wr.octinv) << 16);
            }
        } else {
            const int ref = cur.leaf_base + popc_u((cur.bits >> 8) & 0xffu & below);
            if (stats) stats->prims++;
            if (test_wide_leaf(g, ref, o, d, rf, tmin, best) && ANY) return best;
        }
    }
    return best;
}

}  // namespace hm
