// hm_trace_dev.cuh — warp-cooperative BVH traversal for the sm_100a kernels.
//
// Same per-(ray, primitive) arithmetic as the portable trace<>() in hm_bvh.h (slab test,
// fibre_candidate + fibre_solve, intersect_triangle), so results are bit-identical to the
// host build; what differs is the schedule.  ncu on the earlier "descend, then leaf, then
// solve" loop (profiles/r1c_k_trace) showed the kernel latency-bound on its warp-level
// iteration count: one dependent 64-byte fetch per iteration, with 5.2 of 32 lanes active
// in the node loop because every lane waited for the slowest descent.  This schedule makes
// every warp iteration ONE kind of unit step, taken by as many lanes as possible:
//
//   node step  : fetch one inner node, test both child boxes; inner children go to the
//                lane's stack, leaf children (one primitive reference each) are PARKED in a
//                small per-lane list in shared memory and the lane keeps descending;
//   prim step  : pop one parked reference, fetch its 64-byte primitive; triangles are
//                intersected directly, fibre spans go through the cheap conservative
//                rejects and survivors are parked for the solver;
//   solve step : pop one solver candidate, run the Newton iteration.
//
// The warp votes each iteration: a prim (solve) step runs once enough lanes hold a parked
// reference (candidate), or when no lane can take a node step; otherwise a node step runs.
// Parking defers a leaf by a few node visits, which only costs when that leaf would have
// shortened the ray — rare (a ray tests ~7 primitives for at most a couple of accepted hits).
// Persistent warps pull rays from the queue through one atomic cursor and refill idle lanes
// as soon as a quarter of the warp has finished; commits happen at the top of the loop with
// the whole warp present, so queue appends stay warp-aggregated (one atomic per warp).
#pragma once
#include "hm_bvh.h"

namespace hm {

constexpr int kTraceBlock = 128;   // threads per CTA of every kernel that calls trace_queue
constexpr int kLeafCap = 8;        // parked primitive references per lane
constexpr int kSolveCap = 4;       // parked solver candidates per lane
constexpr int kRefillLanes = 8;    // refill when this many lanes are idle
#ifndef HM_TRACE_PRIM_LANES
#define HM_TRACE_PRIM_LANES 20     // prim step when this many lanes hold a parked reference
#endif
#ifndef HM_TRACE_SOLVE_LANES
#define HM_TRACE_SOLVE_LANES 12    // solve step when this many lanes hold a candidate
#endif
#ifndef HM_TRACE_NODE_LANES
#define HM_TRACE_NODE_LANES 10     // below this many node-ready lanes, parked work goes first
#endif

#ifndef HM_TRACE_PREFETCH
#define HM_TRACE_PREFETCH 0        // 1: prefetch parked primitives and pushed far children into L2/L1
#endif
__device__ __forceinline__ void prefetch_line(const void* p) {
#if HM_TRACE_PREFETCH
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#endif
}

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
    __shared__ int s_leaf[kLeafCap][kTraceBlock];
    __shared__ int s_solve[kSolveCap][kTraceBlock];
    const int tx = threadIdx.x;

    int id = -1;
    V3 o, d, idir, ood;
    RayFrame rf;
    Hit best;
    best.t = tmax; best.prim = -1; best.u = 0.f; best.v = 0.f;
    int cur = kDone, sp = 0, nleaf = 0, nsolve = 0;
    int stack[kStackDepth];
    bool exhausted = false;
    bool any = false;

    while (true) {
        // ---- commit finished rays, refill idle lanes ----
        const bool finished = id >= 0 && cur == kDone && nleaf == 0 && nsolve == 0;
        if (__any_sync(FULL, finished)) {
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
                    cur = g.num_nodes > 0 ? 0 : kDone;
                    sp = 0; nleaf = 0; nsolve = 0;
                }
            }
        }
        if (__all_sync(FULL, id < 0)) break;

        // ---- vote on the step kind ----
        const bool can_node = cur != kDone && nleaf <= kLeafCap - 2;          // a node parks at most 2 references
        const bool can_prim = nleaf > 0 && nsolve < kSolveCap;
        const bool can_solve = nsolve > 0;
        const int n_node = __popc(__ballot_sync(FULL, can_node));
        const int n_prim = __popc(__ballot_sync(FULL, can_prim));
        const int n_solve = __popc(__ballot_sync(FULL, can_solve));
        const bool few_nodes = n_node < HM_TRACE_NODE_LANES;

        if (n_solve >= HM_TRACE_SOLVE_LANES || (n_solve > 0 && few_nodes && n_solve >= n_prim)) {
            // ---- solve step ----
            if (can_solve) {
                const int slot = s_solve[--nsolve][tx];
                const F4* p = g.leaf_data + 4 * (size_t)slot;
                F4 a = load_f4(p + 0), b = load_f4(p + 1), c = load_f4(p + 2), e = load_f4(p + 3);
                FibreCandidate fc;
                // re-run the rejects: best.t may have shrunk since the span was parked
                if (f_as_i(a.w) != best.prim &&
                    fibre_candidate(rf, tmin, best.t, f4_to_v4(a), f4_to_v4(b), f4_to_v4(c), f4_to_v4(e), fc)) {
                    SegHit sh;
                    if (fibre_solve(fc, tmin, best.t, sh)) {
                        best.t = sh.t; best.u = sh.u; best.v = 0.f; best.prim = f_as_i(a.w);
                        if (any) { cur = kDone; sp = 0; nleaf = 0; nsolve = 0; }
                    }
                }
            }
        } else if (n_prim >= HM_TRACE_PRIM_LANES || (n_prim > 0 && few_nodes)) {
            // ---- prim step ----
            if (can_prim) {
                const int slot = s_leaf[--nleaf][tx];
                const F4* p = g.leaf_data + 4 * (size_t)slot;
                F4 a = load_f4(p + 0), b = load_f4(p + 1), c = load_f4(p + 2), e = load_f4(p + 3);
                if (stats) stats[any ? 1 : 0].prims++;
                if (e.w < 0.f) {
                    float t, b1, b2;
                    if (intersect_triangle(o, d, tmin, best.t, V3(a.x, a.y, a.z), V3(b.x, b.y, b.z), V3(c.x, c.y, c.z), t, b1, b2)) {
                        best.t = t; best.u = b1; best.v = b2; best.prim = f_as_i(e.x);
                        if (any) { cur = kDone; sp = 0; nleaf = 0; nsolve = 0; }
                    }
                } else if (f_as_i(a.w) != best.prim) {
                    FibreCandidate fc;
                    if (fibre_candidate(rf, tmin, best.t, f4_to_v4(a), f4_to_v4(b), f4_to_v4(c), f4_to_v4(e), fc)) s_solve[nsolve++][tx] = slot;
                }
            }
        } else {
            // ---- node step ----
            if (can_node) {
                const F4* nd = g.nodes + 4 * (size_t)cur;
                F4 q0 = load_f4(nd + 0), q1 = load_f4(nd + 1), q2 = load_f4(nd + 2), q3 = load_f4(nd + 3);
                if (stats) stats[any ? 1 : 0].nodes++;
                float t0 = slab(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, idir, ood, tmin, best.t);
                float t1 = slab(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, idir, ood, tmin, best.t);
                int c0 = f_as_i(q3.x), c1 = f_as_i(q3.y);
                bool h0 = t0 < 2.9e38f, h1 = t1 < 2.9e38f;
                if (h1 && (!h0 || t1 < t0)) { int tmp = c0; c0 = c1; c1 = tmp; bool th = h0; h0 = h1; h1 = th; }
                // c0 = nearer hit child (if any), c1 = the other hit child (if any)
                if (h1) {
                    if (c1 < 0) { s_leaf[nleaf++][tx] = ~c1; prefetch_line(g.leaf_data + 4 * (size_t)(~c1)); }
                    else { stack[sp++] = c1; prefetch_line(g.nodes + 4 * (size_t)c1); }
                }
                if (h0 && c0 >= 0) {
                    cur = c0;
                } else {
                    if (h0) { s_leaf[nleaf++][tx] = ~c0; prefetch_line(g.leaf_data + 4 * (size_t)(~c0)); }     // parked last: popped first
                    cur = sp > 0 ? stack[--sp] : kDone;
                }
            }
        }
    }
}

}  // namespace hm
