// hm_trace_dev.cuh — warp-cooperative BVH traversal for the sm_100a kernels.
//
// Same per-(ray, primitive) arithmetic as the portable trace<>() in hm_bvh.h (slab test,
// fibre_candidate + fibre_solve, intersect_triangle), so results are bit-identical to the
// host build; what differs is the schedule, which is built around what ncu showed for the
// thread-per-ray loop (profiles/r1_k_shadow_v1: 3.9 of 32 lanes active per instruction,
// issue-bound, DRAM at 1 %):
//
//   * persistent warps pull rays from the queue through one atomic cursor and REFILL idle
//     lanes as soon as a quarter of the warp has finished, instead of waiting for the
//     slowest ray of a fixed 32-ray batch;
//   * each outer iteration is phase-structured and warp-synchronous: inner nodes -> leaf
//     (cheap conservative rejects only) -> curve solver.  Span candidates that survive the
//     rejects are parked in a small per-lane list; the Newton solver runs when at least
//     kSolveLanes lanes have one (or a lane cannot go on without it), so its long,
//     variable-length loop executes with many lanes instead of one or two;
//   * commits happen at the top of the loop with the whole warp present, so queue appends
//     stay warp-aggregated (one atomic per warp).
#pragma once
#include "hm_bvh.h"

namespace hm {

constexpr int kPendMax = 12;       // parked candidates per lane
constexpr int kRefillLanes = 8;    // refill when this many lanes are idle
constexpr int kSolveLanes = 12;    // run the solver when this many lanes have a candidate

// Ops must provide, all __device__:
//   bool fetch(int work, V3& o, V3& d)                 — ray of work item `work`; returns true for
//                                                         an occlusion (any-hit) query
//   void commit(int work, const Hit& h, bool finished) — called by ALL 32 lanes together;
//                                                         `finished` marks lanes with a result
// Closest-hit and any-hit rays share one launch (and one warp): `any` is per lane.
// stats[0] counts closest-hit rays' nodes/prims, stats[1] any-hit rays'.
template <class Ops>
__device__ __forceinline__ void trace_queue(const GeomView& g, int n, int* cursor, Ops& ops, float tmin, float tmax,
                                            TraceStats* stats) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int kDone = 0x7fffffff;

    int id = -1;
    V3 o, d, idir, ood;
    RayFrame rf;
    Hit best;
    best.t = tmax; best.prim = -1; best.u = 0.f; best.v = 0.f;
    int best_slot = -1;
    int cur = kDone, sp = 0, npend = 0;
    int stack[kStackDepth];
    int pend[kPendMax];
    bool exhausted = false;
    bool any = false;

    while (true) {
        // ---- (A) commit finished rays, refill idle lanes ----
        const bool finished = id >= 0 && cur == kDone && npend == 0;
        if (__any_sync(FULL, finished)) {
            if (finished && best_slot >= 0) best.prim = load_i(g.leaf_prim + best_slot);
            ops.commit(id, best, finished);
            if (finished) id = -1;
        }
        const unsigned idle = __ballot_sync(FULL, id < 0);
        if (idle && !exhausted && (idle == FULL || __popc(idle) >= kRefillLanes)) {
            const int cnt = __popc(idle);
            int base = 0;
            if (lane == 0) base = atomicAdd(cursor, cnt);
            base = __shfl_sync(FULL, base, 0);
            if (base + cnt >= n) exhausted = true;
            if (id < 0) {
                const int w = base + __popc(idle & ((1u << lane) - 1u));
                if (w < n) {
                    id = w;
                    any = ops.fetch(w, o, d);
                    const float eps = 1e-20f;
                    V3 dd = V3(fabsf(d.x) > eps ? d.x : (d.x < 0.f ? -eps : eps),
                               fabsf(d.y) > eps ? d.y : (d.y < 0.f ? -eps : eps),
                               fabsf(d.z) > eps ? d.z : (d.z < 0.f ? -eps : eps));
                    idir = V3(1.f / dd.x, 1.f / dd.y, 1.f / dd.z);
                    ood = V3(o.x * idir.x, o.y * idir.y, o.z * idir.z);
                    rf = make_ray_frame(o, d);
                    best.t = tmax; best.prim = -1; best.u = 0.f; best.v = 0.f;
                    best_slot = -1;
                    cur = g.num_nodes > 0 ? 0 : kDone;
                    sp = 0; npend = 0;
                }
            }
        }
        if (__all_sync(FULL, id < 0)) break;

        // ---- (B) inner nodes ----
        while (cur >= 0 && cur != kDone) {
            const F4* nd = g.nodes + 4 * (size_t)cur;
            F4 q0 = load_f4(nd + 0), q1 = load_f4(nd + 1), q2 = load_f4(nd + 2), q3 = load_f4(nd + 3);
            if (stats) stats[any ? 1 : 0].nodes++;
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
        }

        // ---- (C) leaf: triangles now, fibre spans through the cheap rejects only ----
        if (cur < 0 && npend <= kPendMax - kMaxLeaf) {   // room for a whole leaf's spans
            const int code = ~cur;
            const int first = code >> 3;
            const int count = (code & 7) + 1;
            cur = sp > 0 ? stack[--sp] : kDone;
            for (int i = 0; i < count; ++i) {
                const F4* p = g.leaf_data + 4 * (size_t)(first + i);
                F4 a = load_f4(p + 0), b = load_f4(p + 1), c = load_f4(p + 2), e = load_f4(p + 3);
                if (stats) stats[any ? 1 : 0].prims++;
                if (e.w < 0.f) {
                    float t, b1, b2;
                    if (intersect_triangle(o, d, tmin, best.t, V3(a.x, a.y, a.z), V3(b.x, b.y, b.z), V3(c.x, c.y, c.z), t, b1, b2)) {
                        best.t = t; best.u = b1; best.v = b2; best_slot = first + i;
                        if (any) { cur = kDone; sp = 0; npend = 0; break; }
                    }
                } else {
                    FibreCandidate fc;
                    if (fibre_candidate(rf, tmin, best.t, f4_to_v4(a), f4_to_v4(b), f4_to_v4(c), f4_to_v4(e), fc)) pend[npend++] = first + i;
                }
            }
        }

        // ---- (D) curve solver, batched across the warp ----
        const bool has = npend > 0;
        const bool must = has && (npend > kPendMax - kMaxLeaf || cur == kDone);
        const unsigned have = __ballot_sync(FULL, has);
        if (__any_sync(FULL, must) || __popc(have) >= kSolveLanes) {
            if (has) {
                const int slot = pend[--npend];
                const F4* p = g.leaf_data + 4 * (size_t)slot;
                F4 a = load_f4(p + 0), b = load_f4(p + 1), c = load_f4(p + 2), e = load_f4(p + 3);
                FibreCandidate fc;
                // re-run the rejects: best.t may have shrunk since the span was parked
                if (fibre_candidate(rf, tmin, best.t, f4_to_v4(a), f4_to_v4(b), f4_to_v4(c), f4_to_v4(e), fc)) {
                    SegHit sh;
                    if (fibre_solve(fc, tmin, best.t, sh)) {
                        best.t = sh.t; best.u = sh.u; best.v = 0.f; best_slot = slot;
                        if (any) { cur = kDone; sp = 0; npend = 0; }
                    }
                }
            }
        }
    }
}

}  // namespace hm
