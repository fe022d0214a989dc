// hm_trace_dev.cuh — warp-cooperative BVH traversal for the sm_100a kernels.
//
// Same node and primitive arithmetic as the portable trace_wide<>() in hm_bvh.h (wide_node_hits,
// fibre_candidate + fibre_solve, intersect_triangle), so results are bit-identical to the
// host build; what differs is the schedule.  ncu on the earlier "descend, then leaf, then
// solve" loop (profiles/r1c_k_trace) showed the kernel latency-bound on its warp-level
// iteration count: one dependent 64-byte fetch per iteration, with 5.2 of 32 lanes active
// in the node loop because every lane waited for the slowest descent.  This schedule makes
// every warp iteration ONE kind of unit step, taken by as many lanes as possible:
//
//   node step  : fetch one 8-wide quantised node (80 B, hm_bvh.h), test its 8 child boxes; hit
//                inner children become the lane's current group (older groups go to its
//                stack), hit leaf children (one primitive reference each) are PARKED in a
//                small per-lane list in shared memory and the lane keeps descending;
//   prim step  : take parked references, fetch their 64-byte primitives; triangles are
//                intersected directly, fibre spans go through the cheap conservative
//                rejects and survivors are kept for the solver;
//   solve step : run the Newton iteration of the kept candidates.
//
// The warp votes each iteration: a prim (solve) step runs once enough parked references
// (candidates) have collected, or when few lanes can take a node step; otherwise a node step runs.
// Parking defers a leaf by a few node visits; measured against the host's test-at-once order it costs
// 17-23 % more node visits (scripts/trav_split.py vs scripts/bvh_stats.py), but making lanes wait for
// the prim step instead costs more in idle lanes than it saves (HM_TRACE_*_WAIT,
// profiles/r1m_sweep_wait_for_prim_step.txt).  Within a node the child with the smallest entry
// distance goes first, the others in octant order.
// Primitive work is POOLED per warp (HM_TRACE_POOL = 2, the default): a prim step collects up to two parked references
// from every lane into a list of (ray, primitive) pairs and tests 32 pairs at once — the tester lane reads the owning
// lane's ray from shared memory; fibre spans that survive the rejects go, as ray-space control points, into a warp-level
// pool that the solve step hands out one candidate per lane.  Hits return to the owning lane through a shared-memory
// atomicMin on the bits of t.  Which lane tests a pair changes nothing in the arithmetic: results stay bit-identical to
// the host build and deterministic from run to run.  Measured on the bench scene (profiles/r2v, r2ab, r2ac): per-lane
// prim and solve steps 183.2, pooled solver 185.2, + retuned votes 188.0, + pooled prim step 193.1 Mpaths/s.
// Persistent warps pull rays from the queue through one atomic cursor and refill idle lanes
// as soon as a quarter of the warp has finished; commits happen at the top of the loop with
// the whole warp present, so queue appends stay warp-aggregated (one atomic per warp).
#pragma once
#include "hm_bvh.h"

namespace hm {

constexpr int kTraceBlock = 128;   // threads per CTA of every kernel that calls trace_queue
// 15, not 16: with the pools the CTA's shared memory is 26.0 KB, and 6 CTAs (+ 1 KB each of system use) stay inside the
// 164 KB carve-out — one more reference per lane pushes the SM to the 196 KB one and costs 32 KB of L1 (194.7 vs 192.9)
#ifndef HM_LEAF_CAP
#define HM_LEAF_CAP 15
#endif
constexpr int kLeafCap = HM_LEAF_CAP;   // parked primitive references per lane (a wide node can park 8)
constexpr int kSolveCap = 4;       // parked solver candidates per lane (HM_TRACE_POOL == 0)
// HM_TRACE_POOL == 1: fibre spans that survive the conservative rejects go into a WARP-level pool in shared memory
// (their ray-space control points, 16 words each), and the solve step hands the pool out one candidate per lane:
// the Newton iteration runs at up to 32 lanes instead of the ~7 that hold a candidate of their own, and nothing is
// fetched or projected twice.  Hits go back to the owning lane through a shared-memory atomicMin on the bits of t.
// HM_TRACE_POOL == 2: the prim step is pooled too (see the file header).  0 = per-lane prim and solve steps (round 1).
#ifndef HM_TRACE_POOL
#define HM_TRACE_POOL 2
#endif
#ifndef HM_POOL_CAP
#define HM_POOL_CAP 48
#endif
constexpr int kPoolCap = HM_POOL_CAP;   // pool entries per warp: a prim step adds at most 32 (HM_POOL_SOLVE <= cap - 32)
#ifndef HM_POOL_SOLVE
#define HM_POOL_SOLVE 16           // solve step once the pool holds this many candidates (<= kPoolCap - 32)
#endif
#ifndef HM_TRACE_REFILL
#define HM_TRACE_REFILL 8
#endif
constexpr int kRefillLanes = HM_TRACE_REFILL;    // refill when this many lanes are idle
#ifndef HM_TRACE_PRIM_LANES
#define HM_TRACE_PRIM_LANES 16     // prim step when the warp holds this many (ray, primitive) pairs, at most two per lane
                                   // (HM_TRACE_POOL < 2: when this many lanes hold a parked reference; best there: 12)
#endif
#ifndef HM_TRACE_SOLVE_LANES
#define HM_TRACE_SOLVE_LANES 10    // solve step when this many lanes hold a candidate
#endif
#ifndef HM_TRACE_NODE_LANES
#define HM_TRACE_NODE_LANES 10     // below this many node-ready lanes, parked work goes first
#endif

#ifndef HM_TRACE_ANY_WAIT
#define HM_TRACE_ANY_WAIT 0
#endif
#ifndef HM_TRACE_CLOSEST_WAIT
#define HM_TRACE_CLOSEST_WAIT 0
#endif
#ifndef HM_TRACE_NODE_REPEAT
#define HM_TRACE_NODE_REPEAT 3     // node steps per vote
#endif
#ifndef HM_TRACE_PREFETCH
#define HM_TRACE_PREFETCH 0        // 1: prefetch parked primitives and pushed far children into L2/L1
#endif
__device__ __forceinline__ void prefetch_line(const void* p) {
#if HM_TRACE_PREFETCH == 1
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#endif
}
// HM_TRACE_PREFETCH == 2: the first two 128-byte lines of a pushed group's child nodes, into L2 (they are visited a few
// steps later, if at all)
__device__ __forceinline__ void prefetch_group(const void* p) {
#if HM_TRACE_PREFETCH == 2
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
    asm volatile("prefetch.global.L2 [%0];" ::"l"((const char*)p + 128));
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
    __shared__ int s_leaf[kLeafCap][kTraceBlock];
#if HM_TRACE_POOL
    __shared__ float s_pool[kTraceBlock / 32][16][kPoolCap];   // [warp][word][entry]
    __shared__ unsigned s_bt[kTraceBlock];                     // per lane: bits of its best t, lowered by the solvers
    __shared__ float s_bu[kTraceBlock];
    __shared__ int s_bp[kTraceBlock];
    float (*pool)[kPoolCap] = s_pool[threadIdx.x >> 5];
    const int wbase = threadIdx.x & ~31;
    int wpool = 0;           // candidates in the pool (warp-uniform)
#if HM_TRACE_POOL == 2
    // The prim step is pooled as well: lanes hand in up to two parked references each, the warp tests 32 (ray, primitive)
    // pairs at a time.  A tester reads the owner's ray from shared memory and rebuilds its frame (same arithmetic, same
    // bits); triangle hits return through the same atomicMin as the solver's.
    __shared__ float s_ray[6][kTraceBlock];
    __shared__ int s_pair[kTraceBlock];
    __shared__ float s_bv[kTraceBlock];
    __shared__ int s_cur[kTraceBlock];     // the owner's current best primitive (its other references are skipped)
    __shared__ int s_pend[kTraceBlock];    // the owner's candidates in the pool
#endif
#else
    __shared__ int s_solve[kSolveCap][kTraceBlock];
#endif
    const int tx = threadIdx.x;

    int id = -1;
    V3 o, d;
    WideRay wr;
#if HM_TRACE_POOL != 2
    RayFrame rf;
#endif
    Hit best;
    best.t = tmax; best.prim = -1; best.u = 0.f; best.v = 0.f;
    // current group of pending inner children: base index + (imask | permuted hits << 8); older groups on the stack
    int g_base = 0;
    unsigned g_bits = 0;
    int sp = 0, nleaf = 0, nsolve = 0;
    int next_node = -1;      // the nearest hit inner child of the node just tested: visited before the rest of its group
    int2 stack[kWideStack];
    bool exhausted = false;
    bool any = false;

    while (true) {
        // ---- commit finished rays, refill idle lanes ----
#if HM_TRACE_POOL == 2
        if (id >= 0 && nleaf == 0) nsolve = s_pend[tx];
#endif
        const bool finished = id >= 0 && next_node < 0 && (g_bits >> 8) == 0 && sp == 0 && nleaf == 0 && nsolve == 0;
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
                    wr = make_wide_ray(o, d);
#if HM_TRACE_POOL == 2
                    s_ray[0][tx] = o.x; s_ray[1][tx] = o.y; s_ray[2][tx] = o.z;
                    s_ray[3][tx] = d.x; s_ray[4][tx] = d.y; s_ray[5][tx] = d.z;
                    s_pend[tx] = 0;
#else
                    rf = make_ray_frame(o, d);
#endif
                    best.t = tmax; best.prim = -1; best.u = 0.f; best.v = 0.f;
                    // pseudo-group whose slot 0 is the root
                    g_base = 0;
                    g_bits = g.num_wnodes > 0 ? (1u | (1u << (8 + wr.octinv))) : 0u;
                    sp = 0; nleaf = 0; nsolve = 0; next_node = -1;
                }
            }
        }
        if (__all_sync(FULL, id < 0)) break;

        // ---- vote on the step kind ----
        // An occlusion ray ends at its first accepted primitive (72 % of them do on the bench scene): once it holds
        // a parked reference it stops descending and waits for the prim step (HM_TRACE_ANY_WAIT parked references;
        // 0 = never wait).  Closest-hit rays keep descending (HM_TRACE_CLOSEST_WAIT).
        const int wait_at = any ? HM_TRACE_ANY_WAIT : HM_TRACE_CLOSEST_WAIT;
        const bool waits = wait_at > 0 && nleaf >= wait_at;
        const bool can_node = id >= 0 && !waits && (next_node >= 0 || (g_bits >> 8) != 0 || sp > 0) && nleaf <= kLeafCap - 8;   // a node parks at most 8 references
#if HM_TRACE_POOL
        const bool can_prim = nleaf > 0;
        const int n_node = __popc(__ballot_sync(FULL, can_node));
#if HM_TRACE_POOL == 2
        // (ray, primitive) pairs a prim step could test: up to two per lane
        const int n_prim = __popc(__ballot_sync(FULL, can_prim)) + __popc(__ballot_sync(FULL, nleaf >= 2));
#else
        const int n_prim = __popc(__ballot_sync(FULL, can_prim));
#endif
        const bool few_nodes = n_node < HM_TRACE_NODE_LANES;

        if (wpool >= HM_POOL_SOLVE || (wpool > 0 && few_nodes && wpool >= n_prim)) {
            // ---- solve step: the whole pool, 32 candidates per pass ----
            s_bt[tx] = f_as_u(best.t);
            __syncwarp();
            for (int base = 0; base < wpool; base += 32) {
                const int e = base + lane;
                bool won = false;
                int owner = 0, prim = 0;
                SegHit sh;
                sh.t = 0.f; sh.u = 0.f;
                if (e < wpool) {
                    const int tag = __float_as_int(pool[15][e]);
                    owner = wbase + (tag & 31);
                    const float tmax_o = u_as_f(s_bt[owner]);   // the owner's best so far (may be a pass old: conservative)
                    if (pool[14][e] <= tmax_o) {                // depth reject again: tmax may have shrunk since the park
                        FibreCandidate fc;
                        fc.k0 = V3(pool[0][e], pool[1][e], pool[2][e]);
                        fc.k1 = V3(pool[3][e], pool[4][e], pool[5][e]);
                        fc.k2 = V3(pool[6][e], pool[7][e], pool[8][e]);
                        fc.k3 = V3(pool[9][e], pool[10][e], pool[11][e]);
                        fc.r = pool[12][e];
                        fc.u_start = pool[13][e];
                        if (fibre_solve(fc, tmin, tmax_o, sh)) {
                            atomicMin(&s_bt[owner], f_as_u(sh.t));   // t > tmin >= 0: the bit patterns order like the values
                            won = true;
                            prim = tag >> 5;
                        }
                    }
                }
                __syncwarp();
                if (won && s_bt[owner] == f_as_u(sh.t)) { s_bu[owner] = sh.u; s_bp[owner] = prim; }
                __syncwarp();
            }
            wpool = 0;
            nsolve = 0;
#if HM_TRACE_POOL == 2
            s_pend[tx] = 0;
#endif
            const unsigned nt = s_bt[tx];
            if (nt < f_as_u(best.t)) {
                best.t = u_as_f(nt); best.u = s_bu[tx]; best.v = 0.f; best.prim = s_bp[tx];
                if (any) { g_bits = 0; sp = 0; nleaf = 0; next_node = -1; }
            }
        } else if (n_prim >= HM_TRACE_PRIM_LANES || (n_prim > 0 && few_nodes)) {
#if HM_TRACE_POOL == 2
            // ---- prim step, pooled: up to two references per lane, 32 (ray, primitive) pairs per step ----
            s_bt[tx] = f_as_u(best.t);
            s_cur[tx] = best.prim;
            const unsigned m1 = __ballot_sync(FULL, nleaf >= 1), m2 = __ballot_sync(FULL, nleaf >= 2);
            const unsigned lt = (1u << lane) - 1u;
            const int pos = __popc(m1 & lt) + __popc(m2 & lt);
            const int npairs = min(__popc(m1) + __popc(m2), 32);
            if (nleaf >= 1 && pos < 32) {
                s_pair[wbase + pos] = (s_leaf[nleaf - 1][tx] << 5) | lane;
                if (nleaf >= 2 && pos + 1 < 32) { s_pair[wbase + pos + 1] = (s_leaf[nleaf - 2][tx] << 5) | lane; nleaf--; }
                nleaf--;
            }
            __syncwarp();
            bool park = false, won = false;
            FibreCandidate fc;
            float zlo = 0.f, wt = 0.f, wu = 0.f, wv = 0.f;
            int prim_id = 0, ow = wbase, olane = 0;
            if (lane < npairs) {
                const int pr = s_pair[wbase + lane];
                olane = pr & 31;
                ow = wbase + olane;
                const F4* p = g.wleaf_data + 4 * (size_t)(pr >> 5);
                F4 a, b, c, e;
                load_leaf64(p, a, b, c, e);
                if (stats) stats[any ? 1 : 0].prims++;
                const V3 oo(s_ray[0][ow], s_ray[1][ow], s_ray[2][ow]), dd(s_ray[3][ow], s_ray[4][ow], s_ray[5][ow]);
                const float tmax_o = u_as_f(s_bt[ow]);
                if (e.w < 0.f) {
                    if (intersect_triangle(oo, dd, tmin, tmax_o, V3(a.x, a.y, a.z), V3(b.x, b.y, b.z), V3(c.x, c.y, c.z), wt, wu, wv)) {
                        atomicMin(&s_bt[ow], f_as_u(wt));
                        won = true;
                        prim_id = f_as_i(e.x);
                    }
                } else if (f_as_i(a.w) != s_cur[ow]) {
                    const RayFrame rf2 = make_ray_frame(oo, dd);
                    if (fibre_candidate(rf2, tmin, tmax_o, f4_to_v4(a), f4_to_v4(b), f4_to_v4(c), f4_to_v4(e), fc)) {
                        park = true;
                        prim_id = f_as_i(a.w);
                        zlo = fminf(fminf(fc.k1.z, fmaf(fc.k2.z - fc.k0.z, 1.f / 6.f, fc.k1.z)),
                                    fminf(fmaf(fc.k1.z - fc.k3.z, 1.f / 6.f, fc.k2.z), fc.k2.z)) - fc.r;
                    }
                }
            }
            __syncwarp();
            if (won && s_bt[ow] == f_as_u(wt)) { s_bu[ow] = wu; s_bv[ow] = wv; s_bp[ow] = prim_id; }
            const unsigned pm = __ballot_sync(FULL, park);
            if (park) {
                const int e = wpool + __popc(pm & lt);
                pool[0][e] = fc.k0.x; pool[1][e] = fc.k0.y; pool[2][e] = fc.k0.z;
                pool[3][e] = fc.k1.x; pool[4][e] = fc.k1.y; pool[5][e] = fc.k1.z;
                pool[6][e] = fc.k2.x; pool[7][e] = fc.k2.y; pool[8][e] = fc.k2.z;
                pool[9][e] = fc.k3.x; pool[10][e] = fc.k3.y; pool[11][e] = fc.k3.z;
                pool[12][e] = fc.r; pool[13][e] = fc.u_start; pool[14][e] = zlo;
                pool[15][e] = __int_as_float((prim_id << 5) | olane);
                atomicAdd(&s_pend[ow], 1);
            }
            wpool += __popc(pm);
            __syncwarp();
            {
                const unsigned nt = s_bt[tx];
                if (nt < f_as_u(best.t)) {      // a triangle of mine was hit
                    best.t = u_as_f(nt); best.u = s_bu[tx]; best.v = s_bv[tx]; best.prim = s_bp[tx];
                    if (any) { g_bits = 0; sp = 0; nleaf = 0; next_node = -1; }
                }
            }
#else
            // ---- prim step ----
            bool park = false;
            FibreCandidate fc;
            float zlo = 0.f;
            int prim_id = 0;
            if (can_prim) {
                const int ref = s_leaf[--nleaf][tx];
                const F4* p = g.wleaf_data + 4 * (size_t)ref;
                F4 a, b, c, e;
                load_leaf64(p, a, b, c, e);
                if (stats) stats[any ? 1 : 0].prims++;
                if (e.w < 0.f) {
                    float t, b1, b2;
                    if (intersect_triangle(o, d, tmin, best.t, V3(a.x, a.y, a.z), V3(b.x, b.y, b.z), V3(c.x, c.y, c.z), t, b1, b2)) {
                        best.t = t; best.u = b1; best.v = b2; best.prim = f_as_i(e.x);
                        if (any) { g_bits = 0; sp = 0; nleaf = 0; next_node = -1; }
                    }
                } else if (f_as_i(a.w) != best.prim) {
                    if (fibre_candidate(rf, tmin, best.t, f4_to_v4(a), f4_to_v4(b), f4_to_v4(c), f4_to_v4(e), fc)) {
                        park = true;
                        prim_id = f_as_i(a.w);
                        zlo = fminf(fminf(fc.k1.z, fmaf(fc.k2.z - fc.k0.z, 1.f / 6.f, fc.k1.z)),
                                    fminf(fmaf(fc.k1.z - fc.k3.z, 1.f / 6.f, fc.k2.z), fc.k2.z)) - fc.r;
                    }
                }
            }
            const unsigned pm = __ballot_sync(FULL, park);
            if (park) {
                const int e = wpool + __popc(pm & ((1u << lane) - 1u));
                pool[0][e] = fc.k0.x; pool[1][e] = fc.k0.y; pool[2][e] = fc.k0.z;
                pool[3][e] = fc.k1.x; pool[4][e] = fc.k1.y; pool[5][e] = fc.k1.z;
                pool[6][e] = fc.k2.x; pool[7][e] = fc.k2.y; pool[8][e] = fc.k2.z;
                pool[9][e] = fc.k3.x; pool[10][e] = fc.k3.y; pool[11][e] = fc.k3.z;
                pool[12][e] = fc.r; pool[13][e] = fc.u_start; pool[14][e] = zlo;
                pool[15][e] = __int_as_float((prim_id << 5) | lane);
                nsolve++;
            }
            wpool += __popc(pm);
#endif
#else
        const bool can_prim = nleaf > 0 && nsolve < kSolveCap;
        const bool can_solve = nsolve > 0;
        const int n_node = __popc(__ballot_sync(FULL, can_node));
        const int n_prim = __popc(__ballot_sync(FULL, can_prim));
        const int n_solve = __popc(__ballot_sync(FULL, can_solve));
        const bool few_nodes = n_node < HM_TRACE_NODE_LANES;

        if (n_solve >= HM_TRACE_SOLVE_LANES || (n_solve > 0 && few_nodes && n_solve >= n_prim)) {
            // ---- solve step ----
            if (can_solve) {
                const int ref = s_solve[--nsolve][tx];
                const F4* p = g.wleaf_data + 4 * (size_t)ref;
                F4 a, b, c, e;
                load_leaf64(p, a, b, c, e);
                FibreCandidate fc;
                // re-run the rejects: best.t may have shrunk since the span was parked
                if (f_as_i(a.w) != best.prim &&
                    fibre_candidate(rf, tmin, best.t, f4_to_v4(a), f4_to_v4(b), f4_to_v4(c), f4_to_v4(e), fc)) {
                    SegHit sh;
                    if (fibre_solve(fc, tmin, best.t, sh)) {
                        best.t = sh.t; best.u = sh.u; best.v = 0.f; best.prim = f_as_i(a.w);
                        if (any) { g_bits = 0; sp = 0; nleaf = 0; nsolve = 0; next_node = -1; }
                    }
                }
            }
        } else if (n_prim >= HM_TRACE_PRIM_LANES || (n_prim > 0 && few_nodes)) {
            // ---- prim step ----
            if (can_prim) {
                const int ref = s_leaf[--nleaf][tx];
                const F4* p = g.wleaf_data + 4 * (size_t)ref;
                F4 a, b, c, e;
                load_leaf64(p, a, b, c, e);
                if (stats) stats[any ? 1 : 0].prims++;
                if (e.w < 0.f) {
                    float t, b1, b2;
                    if (intersect_triangle(o, d, tmin, best.t, V3(a.x, a.y, a.z), V3(b.x, b.y, b.z), V3(c.x, c.y, c.z), t, b1, b2)) {
                        best.t = t; best.u = b1; best.v = b2; best.prim = f_as_i(e.x);
                        if (any) { g_bits = 0; sp = 0; nleaf = 0; nsolve = 0; next_node = -1; }
                    }
                } else if (f_as_i(a.w) != best.prim) {
                    FibreCandidate fc;
                    if (fibre_candidate(rf, tmin, best.t, f4_to_v4(a), f4_to_v4(b), f4_to_v4(c), f4_to_v4(e), fc)) s_solve[nsolve++][tx] = ref;
                }
            }
#endif
        } else {
            // ---- node step(s) ----
            // HM_TRACE_NODE_REPEAT unit steps per vote: most iterations are node steps, and the
            // commit/refill/vote preamble costs about a quarter of a node test.
#pragma unroll 1
            for (int rep = 0; rep < HM_TRACE_NODE_REPEAT; ++rep) {
                const int wait_at2 = any ? HM_TRACE_ANY_WAIT : HM_TRACE_CLOSEST_WAIT;
                const bool go = id >= 0 && !(wait_at2 > 0 && nleaf >= wait_at2) && (next_node >= 0 || (g_bits >> 8) != 0 || sp > 0) && nleaf <= kLeafCap - 8;
                if (rep > 0 && __popc(__ballot_sync(FULL, go)) < HM_TRACE_NODE_LANES) break;
                if (go) {
                    int ni = next_node;
                    next_node = -1;
                    if (ni < 0) {
                        if ((g_bits >> 8) == 0) { const int2 e = stack[--sp]; g_base = e.x; g_bits = (unsigned)e.y; }
                        const int bit = top_bit(g_bits >> 8);
                        g_bits &= ~(1u << (8 + bit));
                        const int slot = bit ^ wr.octinv;
                        ni = g_base + __popc(g_bits & 0xffu & ((1u << slot) - 1u));
                    }
                    const F4* nd = g.wnodes + 5 * (size_t)ni;
                    F4 w0 = load_f4(nd + 0), w1 = load_f4(nd + 1), w2 = load_f4(nd + 2), w3 = load_f4(nd + 3), w4 = load_f4(nd + 4);
                    if (stats) stats[any ? 1 : 0].nodes++;
                    const unsigned imask = f_as_u(w0.w) >> 24, lmask = f_as_u(w1.z) & 0xffu;
                    unsigned near_key;
                    const unsigned h = wide_node_hits(w0, w2, w3, w4, wr, tmin, best.t, &near_key, g.k47);
                    // The child with the smallest entry distance goes first (an inner one is the next node, a leaf
                    // is parked last = popped first); the others keep the octant order.  On the host build of the
                    // same tree this order alone cuts closest-hit node visits by 15 % and primitive tests by 16 %
                    // (primary rays: 24 % / 36 %; scripts/bvh_stats.py with HM_STATS_SORTED=1).
                    const int near_slot = h ? (int)(near_key & 7u) : 8;
                    const unsigned near_bit = (1u << near_slot) & 0xffu;
                    const int leaf_base = f_as_i(w1.y);
                    // leaves: park far-to-near, so the nearest is popped first
                    unsigned pl = xor_permute8(h & lmask & ~near_bit, wr.octinv);
                    while (pl) {
                        const int b = __ffs((int)pl) - 1;
                        pl &= pl - 1;
                        const int sl = b ^ wr.octinv;
                        const int ref = leaf_base + __popc(lmask & ((1u << sl) - 1u));
                        s_leaf[nleaf++][tx] = ref;
                        prefetch_line(g.wleaf_data + 4 * (size_t)ref);
                    }
                    if (near_bit & lmask) s_leaf[nleaf++][tx] = leaf_base + __popc(lmask & (near_bit - 1u));
                    if (near_bit & imask) next_node = f_as_i(w1.x) + __popc(imask & (near_bit - 1u));
                    const unsigned hi = h & imask & ~near_bit;
                    if (hi) {
                        if (g_bits >> 8) stack[sp++] = make_int2(g_base, (int)g_bits);
                        g_base = f_as_i(w1.x);
                        g_bits = imask | (xor_permute8(hi, wr.octinv) << 8);
                        prefetch_group(g.wnodes + 5 * (size_t)g_base);
                    }
                }
            }
        }
    }
}

}  // namespace hm
