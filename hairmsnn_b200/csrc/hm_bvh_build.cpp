// hm_bvh_build.cpp — host-side binned-SAH builder for the layout in hm_bvh.h.
//
// The reference builds its acceleration structures inside OptiX
// (owlGroupBuildAccel, render_hair_msnn.cu:401,412); this is the replacement.
// Build time is load-time CPU work (SURVEY §8 row a27 neighbour), parallelised over
// subtrees with std::thread.
#include "hm_host.h"

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cstdlib>
#include <cmath>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>
#include <unistd.h>

namespace hm {

namespace {

struct Box {
    float lo[3], hi[3];
    void reset() { for (int k = 0; k < 3; ++k) { lo[k] = FLT_MAX; hi[k] = -FLT_MAX; } }
    void grow(const Box& b) {
        for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], b.lo[k]); hi[k] = std::max(hi[k], b.hi[k]); }
    }
    void grow(const float* p) {
        for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], p[k]); hi[k] = std::max(hi[k], p[k]); }
    }
    float half_area() const {
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (dx < 0.f) return 0.f;
        return dx * dy + dy * dz + dz * dx;
    }
};

struct PrimRef {
    Box box;
    float cen[3];
};

struct NodeRaw {
    float q[12];
    int c0, c1, pad0, pad1;
};
static_assert(sizeof(NodeRaw) == 64, "node must be 64 bytes");

struct Builder {
    const std::vector<PrimRef>& prims;
    std::vector<int>& order;         // permutation being partitioned in place
    std::vector<NodeRaw>& nodes;
    std::atomic<int> next_node{1};   // node 0 = root
    int spawn_depth;

    static constexpr int kBins = 16;

    Box bounds_of(int b, int e) const {
        Box bx; bx.reset();
        for (int i = b; i < e; ++i) bx.grow(prims[order[i]].box);
        return bx;
    }

    // Returns the child code for range [b,e) with bounds `bx`: an inner node index, or
    // ~(position in `order`) for a leaf (always a single reference).
    int build_range(int b, int e, const Box& bx, int depth) {
        int n = e - b;
        if (n <= 1) return ~b;

        Box cb; cb.reset();
        for (int i = b; i < e; ++i) cb.grow(prims[order[i]].cen);

        float best_cost = FLT_MAX;
        int best_axis = -1, best_bin = -1;

        // beyond depth 36 fall back to median splits: bounds the depth (traversal stack: kStackDepth)
        // One pass bins the references on all three axes (the gather through `order` is what costs), and the
        // winning split's child bounds are the unions of its bins — no separate bounds passes.
        Box bin_box[3][kBins];
        int bin_cnt[3][kBins];
        float scale3[3] = {0.f, 0.f, 0.f};
        const bool try_sah = n > 2 && depth < 36;
        if (try_sah) {
            for (int axis = 0; axis < 3; ++axis) {
                float ext = cb.hi[axis] - cb.lo[axis];
                scale3[axis] = ext > 0.f ? kBins / ext : 0.f;
                for (int k = 0; k < kBins; ++k) { bin_box[axis][k].reset(); bin_cnt[axis][k] = 0; }
            }
            for (int i = b; i < e; ++i) {
                const PrimRef& p = prims[order[i]];
                for (int axis = 0; axis < 3; ++axis) {
                    if (!(scale3[axis] > 0.f)) continue;
                    int k = std::min(kBins - 1, std::max(0, (int)((p.cen[axis] - cb.lo[axis]) * scale3[axis])));
                    bin_box[axis][k].grow(p.box);
                    bin_cnt[axis][k]++;
                }
            }
        }
        for (int axis = 0; axis < 3 && try_sah; ++axis) {
            if (!(scale3[axis] > 0.f)) continue;
            float right_area[kBins];
            int right_cnt[kBins];
            Box acc; acc.reset();
            int cnt = 0;
            for (int k = kBins - 1; k > 0; --k) {
                acc.grow(bin_box[axis][k]); cnt += bin_cnt[axis][k];
                right_area[k] = acc.half_area(); right_cnt[k] = cnt;
            }
            acc.reset(); cnt = 0;
            for (int k = 0; k < kBins - 1; ++k) {
                acc.grow(bin_box[axis][k]); cnt += bin_cnt[axis][k];
                if (cnt == 0 || right_cnt[k + 1] == 0) continue;
                float c = acc.half_area() * cnt + right_area[k + 1] * right_cnt[k + 1];
                if (c < best_cost) { best_cost = c; best_axis = axis; best_bin = k; }
            }
        }

        int mid;
        bool bounds_from_bins = false;
        if (best_axis < 0) {
            // two references, or all centroids coincide: split by count
            mid = b + n / 2;
        } else {
            float scale = scale3[best_axis];
            float lo = cb.lo[best_axis];
            int axis = best_axis, bin = best_bin;
            auto it = std::partition(order.begin() + b, order.begin() + e, [&](int pi) {
                int k = std::min(kBins - 1, std::max(0, (int)((prims[pi].cen[axis] - lo) * scale)));
                return k <= bin;
            });
            mid = (int)(it - order.begin());
            if (mid == b || mid == e) mid = b + n / 2;
            else bounds_from_bins = true;
        }

        Box lb, rb;
        if (bounds_from_bins) {
            lb.reset(); rb.reset();
            for (int k = 0; k <= best_bin; ++k) lb.grow(bin_box[best_axis][k]);
            for (int k = best_bin + 1; k < kBins; ++k) rb.grow(bin_box[best_axis][k]);
        } else {
            lb = bounds_of(b, mid); rb = bounds_of(mid, e);
        }
        int me = next_node.fetch_add(1);
        int cl, cr;
        if (depth < spawn_depth && n > 4096) {
            int cl_local = 0;
            std::thread th([&]() { cl_local = build_range(b, mid, lb, depth + 1); });
            cr = build_range(mid, e, rb, depth + 1);
            th.join();
            cl = cl_local;
        } else {
            cl = build_range(b, mid, lb, depth + 1);
            cr = build_range(mid, e, rb, depth + 1);
        }
        NodeRaw& nd = nodes[me];
        nd.q[0] = lb.lo[0]; nd.q[1] = lb.lo[1]; nd.q[2] = lb.lo[2];
        nd.q[3] = lb.hi[0]; nd.q[4] = lb.hi[1]; nd.q[5] = lb.hi[2];
        nd.q[6] = rb.lo[0]; nd.q[7] = rb.lo[1]; nd.q[8] = rb.lo[2];
        nd.q[9] = rb.hi[0]; nd.q[10] = rb.hi[1]; nd.q[11] = rb.hi[2];
        nd.c0 = cl; nd.c1 = cr; nd.pad0 = 0; nd.pad1 = 0;
        return me;
    }
};

// Blossom of a cubic Bezier (b0..b3) at (t0, t1, t2): control points of sub-curves come from
// evaluating it at mixed end parameters.
static void blossom(const float b[4][3], float t0, float t1, float t2, float out[3]) {
    for (int k = 0; k < 3; ++k) {
        float a0 = b[0][k] + (b[1][k] - b[0][k]) * t0, a1 = b[1][k] + (b[2][k] - b[1][k]) * t0, a2 = b[2][k] + (b[3][k] - b[2][k]) * t0;
        float c0 = a0 + (a1 - a0) * t1, c1 = a1 + (a2 - a1) * t1;
        out[k] = c0 + (c1 - c0) * t2;
    }
}

// Bounds of the part u in [u0, u1] of a Catmull-Rom span: Bezier hull of the sub-curve + radius.
void segment_bounds(const F4* cps, int cp0, float u0, float u1, PrimRef& pr) {
    const F4& k0 = cps[cp0], &k1 = cps[cp0 + 1], &k2 = cps[cp0 + 2], &k3 = cps[cp0 + 3];
    float r = std::max(std::max(k0.w, k1.w), std::max(k2.w, k3.w));
    float b[4][3] = {{k1.x, k1.y, k1.z},
                     {k1.x + (k2.x - k0.x) / 6.f, k1.y + (k2.y - k0.y) / 6.f, k1.z + (k2.z - k0.z) / 6.f},
                     {k2.x - (k3.x - k1.x) / 6.f, k2.y - (k3.y - k1.y) / 6.f, k2.z - (k3.z - k1.z) / 6.f},
                     {k2.x, k2.y, k2.z}};
    pr.box.reset();
    if (u0 <= 0.f && u1 >= 1.f) {
        for (int i = 0; i < 4; ++i) pr.box.grow(b[i]);
    } else {
        float q[3];
        blossom(b, u0, u0, u0, q); pr.box.grow(q);
        blossom(b, u0, u0, u1, q); pr.box.grow(q);
        blossom(b, u0, u1, u1, q); pr.box.grow(q);
        blossom(b, u1, u1, u1, q); pr.box.grow(q);
    }
    // pad by radius plus a relative epsilon so fp32 round-off in the ray-space
    // solver can never place a hit outside its own box
    for (int k = 0; k < 3; ++k) {
        float pad = r + 1e-5f * std::max(fabsf(pr.box.lo[k]), fabsf(pr.box.hi[k])) + 1e-6f;
        if (u0 > 0.f || u1 < 1.f) pad += 1e-4f * (pr.box.hi[k] - pr.box.lo[k]) + 1e-5f;   // blossom round-off
        pr.box.lo[k] -= pad; pr.box.hi[k] += pad;
        pr.cen[k] = 0.5f * (pr.box.lo[k] + pr.box.hi[k]);
    }
}

void triangle_bounds(const F4* tv, int ti, PrimRef& pr) {
    pr.box.reset();
    for (int v = 0; v < 3; ++v) {
        float p[3] = {tv[3 * ti + v].x, tv[3 * ti + v].y, tv[3 * ti + v].z};
        pr.box.grow(p);
    }
    for (int k = 0; k < 3; ++k) {
        float pad = 1e-5f * std::max(fabsf(pr.box.lo[k]), fabsf(pr.box.hi[k])) + 1e-6f;
        pr.box.lo[k] -= pad; pr.box.hi[k] += pad;
        pr.cen[k] = 0.5f * (pr.box.lo[k] + pr.box.hi[k]);
    }
}


// ---------------------------------------------------------------------------------
// Binary tree -> 8-wide quantised tree (layout: hm_bvh.h).
// ---------------------------------------------------------------------------------
struct WideChild {
    int code;      // binary child code: inner node index, or ~leaf slot
    Box box;
};

struct WideBuilder {
    const std::vector<NodeRaw>& bin;
    const std::vector<F4>& leaf_data;   // binary leaf slots (64 B each)
    std::vector<F4>& wnodes;
    std::vector<F4>& wleaf;
    int max_depth = 0;
    // optional (HM_BVH_COLLAPSE=dp): the surface-area-optimal cut of Ylitie et al. 2017, section 4.1, for single-
    // reference leaves: cut[n][i-1] = how the subtree of binary node n is best represented by at most i roots
    // (1..7: that many roots for the left child; 0: use the (i-1)-root solution); cut[n][7] = left share of 8
    const std::vector<unsigned char>* cut = nullptr;

    static Box child_box(const NodeRaw& n, int k) {
        Box b;
        for (int a = 0; a < 3; ++a) { b.lo[a] = n.q[6 * k + a]; b.hi[a] = n.q[6 * k + 3 + a]; }
        return b;
    }

    // children of the wide node that replaces binary node `bi`: open the inner child with the
    // largest surface area until there are 8 (or only leaves are left)
    // roots of the optimal forest of at most `budget` trees below child `code` with bounds `box`
    void expand(int code, const Box& box, int budget, WideChild* out, int& cnt) const {
        if (code < 0 || budget <= 1) { out[cnt++] = WideChild{code, box}; return; }
        const unsigned char* c = cut->data() + 8 * (size_t)code;
        int i = budget;
        while (i > 1 && c[i - 1] == 0) --i;       // fall back to fewer roots
        if (i <= 1) { out[cnt++] = WideChild{code, box}; return; }
        const NodeRaw& n = bin[code];
        const int k = c[i - 1];
        expand(n.c0, child_box(n, 0), k, out, cnt);
        expand(n.c1, child_box(n, 1), i - k, out, cnt);
    }

    int gather(int bi, WideChild* out) const {
        const NodeRaw& n = bin[bi];
        int cnt = 0;
        if (cut && !(n.q[6] > n.q[9])) {
            const int k = (*cut)[8 * (size_t)bi + 7];
            expand(n.c0, child_box(n, 0), k, out, cnt);
            expand(n.c1, child_box(n, 1), 8 - k, out, cnt);
            return cnt;
        }
        out[cnt++] = WideChild{n.c0, child_box(n, 0)};
        if (!(n.q[6] > n.q[9])) out[cnt++] = WideChild{n.c1, child_box(n, 1)};   // inverted box: the unreachable twin of a single-reference tree
        while (cnt < 8) {
            int pick = -1; float best = -1.f;
            for (int i = 0; i < cnt; ++i)
                if (out[i].code >= 0) { float a = out[i].box.half_area(); if (a > best) { best = a; pick = i; } }
            if (pick < 0) break;
            const NodeRaw& c = bin[out[pick].code];
            out[pick] = WideChild{c.c0, child_box(c, 0)};
            out[cnt++] = WideChild{c.c1, child_box(c, 1)};
        }
        return cnt;
    }

    // Fills wide node `wi` from binary node `bi`; children blocks are reserved here, so the inner
    // children of a node are contiguous (child_base + rank) and so are its leaf references.
    struct Item { int wi, bi, depth; Box bounds; };
    // `deferred` (optional): subtrees rooted at depth > defer_below are not expanded here; their
    // root slots stay reserved and the items are handed back for parallel expansion.
    void emit(int wi, int bi, const Box& bounds, int depth, std::vector<Item>* deferred = nullptr, int defer_below = 0) {
        std::vector<Item> todo;
        todo.push_back(Item{wi, bi, depth, bounds});
        while (!todo.empty()) {
            Item it = todo.back(); todo.pop_back();
            if (deferred && it.depth > defer_below) { deferred->push_back(it); continue; }
            max_depth = std::max(max_depth, it.depth);
            WideChild ch[8];
            const int cnt = gather(it.bi, ch);

            // octant slots: greedy assignment maximising sum of (slot sign) . (child centre - node centre)
            float nc[3];
            for (int a = 0; a < 3; ++a) nc[a] = 0.5f * (it.bounds.lo[a] + it.bounds.hi[a]);
            int slot_of_child[8]; bool slot_used[8] = {false, false, false, false, false, false, false, false};
            bool child_done[8] = {false, false, false, false, false, false, false, false};
            float cost[8][8];
            for (int c = 0; c < cnt; ++c)
                for (int sl = 0; sl < 8; ++sl) {
                    float v = 0.f;
                    for (int a = 0; a < 3; ++a) {
                        float rel = 0.5f * (ch[c].box.lo[a] + ch[c].box.hi[a]) - nc[a];
                        v += ((sl >> a) & 1) ? rel : -rel;
                    }
                    cost[c][sl] = v;
                }
            for (int round = 0; round < cnt; ++round) {
                int bc = -1, bs = -1; float bv = -FLT_MAX;
                for (int c = 0; c < cnt; ++c) {
                    if (child_done[c]) continue;
                    for (int sl = 0; sl < 8; ++sl)
                        if (!slot_used[sl] && cost[c][sl] > bv) { bv = cost[c][sl]; bc = c; bs = sl; }
                }
                child_done[bc] = true; slot_used[bs] = true; slot_of_child[bc] = bs;
            }
            int child_in_slot[8];
            for (int sl = 0; sl < 8; ++sl) child_in_slot[sl] = -1;
            for (int c = 0; c < cnt; ++c) child_in_slot[slot_of_child[c]] = c;

            // quantisation frame: cell = 2^e per axis with 255 cells covering the node
            unsigned ebits[3]; float cell[3];
            for (int a = 0; a < 3; ++a) {
                float ext = it.bounds.hi[a] - it.bounds.lo[a];
                int e = -100;
                if (ext > 0.f) e = std::max(-100, (int)ceilf(log2f(ext / 255.f)));
                while (it.bounds.lo[a] + 255.f * ldexpf(1.f, e) < it.bounds.hi[a]) e++;
                cell[a] = ldexpf(1.f, e);
                ebits[a] = (unsigned)(e + 127);
            }
            unsigned imask = 0, lmask = 0;
            unsigned char qlo[3][8], qhi[3][8];
            for (int sl = 0; sl < 8; ++sl) {
                const int c = child_in_slot[sl];
                for (int a = 0; a < 3; ++a) { qlo[a][sl] = 255; qhi[a][sl] = 0; }   // empty slot: inverted box
                if (c < 0) continue;
                if (ch[c].code >= 0) imask |= 1u << sl; else lmask |= 1u << sl;
                for (int a = 0; a < 3; ++a) {
                    const float org = it.bounds.lo[a];
                    int lo = (int)floorf((ch[c].box.lo[a] - org) / cell[a] - 1.f / 64.f);   // margin: hm_bvh.h byte_biased
                    lo = std::max(0, std::min(255, lo));
                    while (lo > 0 && org + (float)lo * cell[a] > ch[c].box.lo[a]) lo--;
                    int hi = (int)ceilf((ch[c].box.hi[a] - org) / cell[a] + 1.f / 64.f);
                    hi = std::max(0, std::min(255, hi));
                    while (hi < 255 && org + (float)hi * cell[a] < ch[c].box.hi[a]) hi++;
                    qlo[a][sl] = (unsigned char)lo; qhi[a][sl] = (unsigned char)hi;
                }
            }
            const int n_inner = __builtin_popcount(imask), n_leaf = __builtin_popcount(lmask);
            const int child_base = (int)(wnodes.size() / 5);
            const int leaf_base = (int)(wleaf.size() / 4);
            wnodes.resize(wnodes.size() + 5 * (size_t)n_inner);
            wleaf.resize(wleaf.size() + 4 * (size_t)n_leaf);

            auto as_f = [](unsigned u) { float f; memcpy(&f, &u, 4); return f; };
            auto pack4 = [](const unsigned char* q) { return (unsigned)q[0] | ((unsigned)q[1] << 8) | ((unsigned)q[2] << 16) | ((unsigned)q[3] << 24); };
            F4* w = wnodes.data() + 5 * (size_t)it.wi;
            w[0] = F4{it.bounds.lo[0], it.bounds.lo[1], it.bounds.lo[2], as_f(ebits[0] | (ebits[1] << 8) | (ebits[2] << 16) | (imask << 24))};
            w[1] = F4{as_f((unsigned)child_base), as_f((unsigned)leaf_base), as_f(lmask), 0.f};
            w[2] = F4{as_f(pack4(qlo[0])), as_f(pack4(qlo[0] + 4)), as_f(pack4(qlo[1])), as_f(pack4(qlo[1] + 4))};
            w[3] = F4{as_f(pack4(qlo[2])), as_f(pack4(qlo[2] + 4)), as_f(pack4(qhi[0])), as_f(pack4(qhi[0] + 4))};
            w[4] = F4{as_f(pack4(qhi[1])), as_f(pack4(qhi[1] + 4)), as_f(pack4(qhi[2])), as_f(pack4(qhi[2] + 4))};

            int ri = 0, rl = 0;
            for (int sl = 0; sl < 8; ++sl) {
                const int c = child_in_slot[sl];
                if (c < 0) continue;
                if (ch[c].code >= 0) {
                    // the child's own frame is its DEQUANTISED box clipped to nothing tighter than its true
                    // bounds: use the true bounds (they lie inside the dequantised box)
                    todo.push_back(Item{child_base + ri, ch[c].code, it.depth + 1, ch[c].box});
                    ri++;
                } else {
                    const F4* src = leaf_data.data() + 4 * (size_t)(~ch[c].code);
                    F4* dst = wleaf.data() + 4 * (size_t)(leaf_base + rl);
                    dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
                    rl++;
                }
            }
        }
    }
};

}  // namespace

void build_bvh(const HostGeometry& geo, HostBvh& out, int threads_hint) {
    const int ns = (int)geo.seg_cp.size();
    const int nt = (int)(geo.tri_verts.size() / 3);
    const int nprim = ns + nt;
    const bool verbose = getenv("HM_BVH_VERBOSE") != nullptr;
    auto t_start = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!verbose) return;
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[bvh] %-28s %.2f s\n", what, std::chrono::duration<double>(now - t_start).count());
        t_start = now;
    };
    out.nodes.clear(); out.leaf_data.clear(); out.leaf_code.clear(); out.leaf_prim.clear();
    out.wnodes.clear(); out.wleaf_data.clear();
    if (nprim == 0) return;

    // References.  Each fibre segment enters the tree as k references, one per sub-span of its
    // parameter range, each with the tight bounds of its piece of the curve: a thin diagonal
    // tube fills a tiny fraction of its own box, and k pieces cut the total box area ~k-fold
    // (measured on the curly scene, k = 4: 128 -> 82 nodes and 30 -> 7 primitive tests per
    // incoherent ray).  k follows the chord length in units of `span_len` (HM_BVH_SPAN,
    // default 5 radii), capped by HM_BVH_SPLIT (default 16).  Triangles get one reference.
    // Swept on the bench scene with the 8-wide tree (gpurun_out/sweep*.txt, B200): split/span
    // 4/20 -> 176, 8/10 -> 190, 16/5 -> 194 Mpaths/s (9.1 -> 4.6 primitive tests per ray at the
    // same ~33 node visits).
    int max_split = 16;
    if (const char* e = getenv("HM_BVH_SPLIT")) max_split = std::max(1, std::min(16, atoi(e)));
    float span_radii = 5.f;
    if (const char* e = getenv("HM_BVH_SPAN")) span_radii = std::max(1.f, (float)atof(e));
    std::vector<int> ref_first(ns + 1);
    ref_first[0] = 0;
    for (int i = 0; i < ns; ++i) {
        const F4* c = geo.cps.data() + geo.seg_cp[i];
        float dx = c[2].x - c[1].x, dy = c[2].y - c[1].y, dz = c[2].z - c[1].z;
        float len = sqrtf(dx * dx + dy * dy + dz * dz);
        float r = std::max(c[1].w, 1e-6f);
        int k = (int)ceilf(len / (span_radii * r));
        ref_first[i + 1] = ref_first[i] + std::max(1, std::min(max_split, k));
    }
    const int nsr = ref_first[ns];
    const int n = nsr + nt;
    std::vector<int> ref_prim(n);
    std::vector<PrimRef> prims(n);
    unsigned hw = threads_hint > 0 ? (unsigned)threads_hint : std::max(1u, std::thread::hardware_concurrency());
    {
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < hw; ++t) {
            pool.emplace_back([&, t]() {
                for (int i = (int)t; i < nprim; i += (int)hw) {
                    if (i < ns) {
                        const int k = ref_first[i + 1] - ref_first[i];
                        for (int j = 0; j < k; ++j) {
                            const int ri = ref_first[i] + j;
                            ref_prim[ri] = i;
                            segment_bounds(geo.cps.data(), geo.seg_cp[i], (float)j / k, (float)(j + 1) / k, prims[ri]);
                        }
                    } else {
                        const int ri = nsr + (i - ns);
                        ref_prim[ri] = i;
                        triangle_bounds(geo.tri_verts.data(), i - ns, prims[ri]);
                    }
                }
            });
        }
        for (auto& th : pool) th.join();
    }

    lap("references");
    std::vector<int> order(n);
    for (int i = 0; i < n; ++i) order[i] = i;
    std::vector<NodeRaw> nodes((size_t)std::max(n + 1, 2));   // n single-reference leaves -> n - 1 inner nodes (+ root slot)

    int spawn_depth = 0;
    while ((1u << spawn_depth) < 2 * hw && spawn_depth < 8) spawn_depth++;

    Builder bld{prims, order, nodes, {}, spawn_depth};
    bld.next_node.store(1);
    Box root; root.reset();
    for (int i = 0; i < n; ++i) root.grow(prims[i].box);

    int code = bld.build_range(0, n, root, 0);
    lap("binary SAH build");
    int used = bld.next_node.load();
    if (code < 0) {
        // a single reference: root with child0 = the leaf and an unreachable child1
        NodeRaw& nd = nodes[0];
        for (int k = 0; k < 3; ++k) { nd.q[k] = root.lo[k]; nd.q[3 + k] = root.hi[k]; nd.q[6 + k] = FLT_MAX; nd.q[9 + k] = -FLT_MAX; }
        nd.c0 = code; nd.c1 = code; nd.pad0 = nd.pad1 = 0;
        used = 1;
    } else {
        // `code` is the index of the top node; move it into slot 0 (unused so far)
        nodes[0] = nodes[code];
    }

    // Re-layout in depth-first order for locality and to drop dead slots; leaf slots are
    // handed out in the same order (first reference of a primitive wins), so primitives that
    // are neighbours in the tree are neighbours in memory.
    std::vector<NodeRaw> packed;
    packed.reserve(used);
    std::vector<int> slot_of(nprim, -1);
    std::vector<int> prim_of_slot;
    prim_of_slot.reserve(nprim);
    auto leaf_slot = [&](int leaf_code_in) {
        int prim = ref_prim[order[~leaf_code_in]];
        if (slot_of[prim] < 0) { slot_of[prim] = (int)prim_of_slot.size(); prim_of_slot.push_back(prim); }
        return ~slot_of[prim];
    };
    {
        struct Item { int src; int dst; };
        std::vector<Item> st;
        packed.push_back(nodes[0]);
        st.push_back({0, 0});
        while (!st.empty()) {
            Item it = st.back(); st.pop_back();
            NodeRaw src = nodes[it.src];
            int c[2] = {src.c0, src.c1};
            int newc[2];
            for (int k = 0; k < 2; ++k) {
                if (c[k] >= 0) {
                    newc[k] = (int)packed.size();
                    packed.push_back(nodes[c[k]]);
                } else newc[k] = leaf_slot(c[k]);
            }
            packed[it.dst].c0 = newc[0];
            packed[it.dst].c1 = newc[1];
            // push right first so the left subtree is laid out right after its parent pair
            if (c[1] >= 0) st.push_back({c[1], newc[1]});
            if (c[0] >= 0) st.push_back({c[0], newc[0]});
        }
    }

    lap("depth-first relayout");
    // the build-order arrays are done with (43 M references: 1.5 + 2.8 GB on the bench scene)
    std::vector<NodeRaw>().swap(nodes);
    std::vector<PrimRef>().swap(prims);
    std::vector<int>().swap(order);
    std::vector<int>().swap(ref_prim);
    out.nodes.resize(packed.size() * 4);
    memcpy(out.nodes.data(), packed.data(), packed.size() * sizeof(NodeRaw));
    out.leaf_code.resize(nprim);
    out.leaf_prim.resize(nprim);
    out.leaf_data.resize(4 * (size_t)nprim);
    auto id_bits = [](int id) { float f; memcpy(&f, &id, 4); return f; };
    for (int i = 0; i < nprim; ++i) {
        const int p = prim_of_slot[i];
        out.leaf_prim[i] = p;
        out.leaf_code[i] = p < ns ? geo.seg_cp[p] : ((p - ns) | kTriTag);
        F4* dst = out.leaf_data.data() + 4 * (size_t)i;
        if (p < ns) {
            const F4* cp = geo.cps.data() + geo.seg_cp[p];
            dst[0] = cp[0]; dst[1] = cp[1]; dst[2] = cp[2]; dst[3] = cp[3];
            dst[0].w = id_bits(p);
            if (dst[3].w < 0.f) dst[3].w = 0.f;
        } else {
            const F4* tv = geo.tri_verts.data() + 3 * (size_t)(p - ns);
            dst[0] = tv[0]; dst[1] = tv[1]; dst[2] = tv[2];
            dst[3] = F4{id_bits(p), 0.f, 0.f, -1.f};
        }
    }

    lap("leaf slots");
    // the structure the GPU traverses
    out.wnodes.clear(); out.wleaf_data.clear();
    out.wnodes.resize(5);
    WideBuilder wb{packed, out.leaf_data, out.wnodes, out.wleaf_data};
    std::vector<unsigned char> cut;
    // Which binary nodes become wide nodes: the surface-area-optimal cut (default) or the greedy "open the largest
    // child" rule (HM_BVH_COLLAPSE=greedy).  Bench scene: 9.4 M instead of 14.1 M wide nodes (755 MB instead of
    // 1.13 GB), 2-4 % fewer node visits per ray, 226 -> 231 Mpaths/s (profiles/r1n_sweep_dp_collapse.txt).
    {
        const char* e = getenv("HM_BVH_COLLAPSE");
        if (!(e && !strcmp(e, "greedy")) && packed.size() > 1) {
            // post-order over the binary tree (children before their parent); subtrees near the top run as
            // parallel tasks — every node's table depends on its own subtree only, so the result does not
            // depend on the schedule
            const size_t nb = packed.size();
            std::vector<float> cost(7 * nb);
            cut.assign(8 * nb, 0);
            auto C = [&](int code, int i) -> float { return code < 0 ? 0.f : cost[7 * (size_t)code + (i - 1)]; };   // leaves: constant, dropped
            auto solve_node = [&](size_t idx) {
                const NodeRaw& n = packed[idx];
                Box nbx = WideBuilder::child_box(n, 0);
                if (!(n.q[6] > n.q[9])) nbx.grow(WideBuilder::child_box(n, 1));
                auto distribute = [&](int j, int& best_k) {
                    float best = FLT_MAX; best_k = 1;
                    for (int k = 1; k < j; ++k) {
                        float v = C(n.c0, std::min(k, 7)) + C(n.c1, std::min(j - k, 7));
                        if (v < best) { best = v; best_k = k; }
                    }
                    return best;
                };
                int k8;
                const float d8 = distribute(8, k8);
                cut[8 * idx + 7] = (unsigned char)k8;
                cost[7 * idx + 0] = d8 + nbx.half_area();          // as one root: a wide node of its own
                for (int i = 2; i <= 7; ++i) {
                    int k;
                    const float d = distribute(i, k);
                    if (d < cost[7 * idx + (i - 2)]) { cost[7 * idx + (i - 1)] = d; cut[8 * idx + (i - 1)] = (unsigned char)k; }
                    else { cost[7 * idx + (i - 1)] = cost[7 * idx + (i - 2)]; cut[8 * idx + (i - 1)] = 0; }
                }
            };
            // iterative post-order of one subtree
            auto solve_subtree = [&](int root_idx) {
                std::vector<std::pair<int, bool>> st;
                st.push_back({root_idx, false});
                while (!st.empty()) {
                    auto [idx, expanded] = st.back();
                    if (expanded) { st.pop_back(); solve_node((size_t)idx); continue; }
                    st.back().second = true;
                    const NodeRaw& n = packed[idx];
                    if (n.c1 >= 0 && n.c1 != idx) st.push_back({n.c1, false});
                    if (n.c0 >= 0) st.push_back({n.c0, false});
                }
            };
            // frontier: inner nodes at depth `par_depth`; everything above is solved afterwards, bottom-up
            const int par_depth = hw > 1 ? 10 : 0;
            std::vector<int> frontier, upper;
            {
                std::vector<std::pair<int, int>> st;
                st.push_back({0, 0});
                while (!st.empty()) {
                    auto [idx, dep] = st.back(); st.pop_back();
                    if (dep >= par_depth) { frontier.push_back(idx); continue; }
                    upper.push_back(idx);     // pre-order: parents before children
                    const NodeRaw& n = packed[idx];
                    if (n.c0 >= 0) st.push_back({n.c0, dep + 1});
                    if (n.c1 >= 0 && !(n.q[6] > n.q[9])) st.push_back({n.c1, dep + 1});
                }
            }
            {
                std::atomic<size_t> next{0};
                std::vector<std::thread> pool;
                for (unsigned t = 0; t < hw; ++t)
                    pool.emplace_back([&]() { for (size_t i; (i = next.fetch_add(1)) < frontier.size();) solve_subtree(frontier[i]); });
                for (auto& th : pool) th.join();
            }
            for (size_t i = upper.size(); i-- > 0;) solve_node((size_t)upper[i]);   // reverse pre-order: children first
            wb.cut = &cut;
        }
    }
    lap("optimal cut (dynamic programme)");
    // top levels sequentially, then one task per deferred subtree: each expands into private arrays
    // (its root at local index 0) that are appended to the global ones with their base indices shifted
    std::vector<WideBuilder::Item> subtrees;
    wb.emit(0, 0, root, 1, &subtrees, hw > 1 ? 3 : kWideStack + 1);
    int max_depth = wb.max_depth;
    if (!subtrees.empty()) {
        struct Local { std::vector<F4> nodes, leaves; int depth = 0; };
        std::vector<Local> loc(subtrees.size());
        std::atomic<size_t> next{0};
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < hw; ++t)
            pool.emplace_back([&]() {
                for (size_t i; (i = next.fetch_add(1)) < subtrees.size();) {
                    loc[i].nodes.resize(5);
                    WideBuilder lb{packed, out.leaf_data, loc[i].nodes, loc[i].leaves};
                    lb.cut = wb.cut;
                    lb.emit(0, subtrees[i].bi, subtrees[i].bounds, subtrees[i].depth);
                    loc[i].depth = lb.max_depth;
                }
            });
        for (auto& th : pool) th.join();
        auto as_u = [](float f) { unsigned u; memcpy(&u, &f, 4); return u; };
        auto as_f = [](unsigned u) { float f; memcpy(&f, &u, 4); return f; };
        // placement: subtree i's nodes 1.. go to node_off[i].., its leaves to leaf_off[i].. (prefix sums in task
        // order, so the arrays come out the same for any schedule); then every task shifts its base indices
        // and copies itself into place
        const size_t ns_ = subtrees.size();
        std::vector<size_t> node_off(ns_ + 1), leaf_off(ns_ + 1);
        node_off[0] = out.wnodes.size() / 5; leaf_off[0] = out.wleaf_data.size() / 4;
        for (size_t i = 0; i < ns_; ++i) {
            max_depth = std::max(max_depth, loc[i].depth);
            node_off[i + 1] = node_off[i] + (loc[i].nodes.size() / 5 - 1);
            leaf_off[i + 1] = leaf_off[i] + loc[i].leaves.size() / 4;
        }
        out.wnodes.resize(5 * node_off[ns_]);
        out.wleaf_data.resize(4 * leaf_off[ns_]);
        next.store(0);
        pool.clear();
        for (unsigned t = 0; t < hw; ++t)
            pool.emplace_back([&]() {
                for (size_t i; (i = next.fetch_add(1)) < ns_;) {
                    const size_t n_local = loc[i].nodes.size() / 5;
                    const unsigned noff = (unsigned)node_off[i], loff = (unsigned)leaf_off[i];   // local index k >= 1 -> noff + k - 1
                    for (size_t k = 0; k < n_local; ++k) {
                        F4* w = loc[i].nodes.data() + 5 * k;
                        w[1].x = as_f(as_u(w[1].x) + noff - 1u);
                        w[1].y = as_f(as_u(w[1].y) + loff);
                    }
                    memcpy(out.wnodes.data() + 5 * (size_t)subtrees[i].wi, loc[i].nodes.data(), 5 * sizeof(F4));
                    if (n_local > 1) memcpy(out.wnodes.data() + 5 * node_off[i], loc[i].nodes.data() + 5, (n_local - 1) * 5 * sizeof(F4));
                    if (!loc[i].leaves.empty()) memcpy(out.wleaf_data.data() + 4 * leaf_off[i], loc[i].leaves.data(), loc[i].leaves.size() * sizeof(F4));
                    std::vector<F4>().swap(loc[i].nodes);
                    std::vector<F4>().swap(loc[i].leaves);
                }
            });
        for (auto& th : pool) th.join();
    }
    out.wide_depth = max_depth;
    wb.max_depth = max_depth;
    lap("wide collapse");
    if (wb.max_depth > kWideStack)
        throw std::runtime_error("BVH: wide tree deeper than the traversal stack");
}

// ---------------------------------------------------------------------------------
// On-disk cache of the tree the GPU traverses (build once, reuse).
// ---------------------------------------------------------------------------------
namespace {
constexpr uint64_t kCacheMagic = 0x31485642574d4848ull;   // "HHMWBVH1"

uint64_t fnv1a(uint64_t h, const void* data, size_t n) {
    // 8 bytes per step (word-wise FNV variant): this is a cache key, not a cryptographic digest
    const unsigned char* p = (const unsigned char*)data;
    size_t i = 0;
    for (; i + 8 <= n; i += 8) { uint64_t w; memcpy(&w, p + i, 8); h = (h ^ w) * 0x100000001b3ull; }
    for (; i < n; ++i) h = (h ^ p[i]) * 0x100000001b3ull;
    return h;
}
}  // namespace

uint64_t bvh_cache_key(const HostGeometry& geo) {
    uint64_t h = 0xcbf29ce484222325ull;
    h = fnv1a(h, geo.cps.data(), geo.cps.size() * sizeof(F4));
    h = fnv1a(h, geo.seg_cp.data(), geo.seg_cp.size() * sizeof(int));
    h = fnv1a(h, geo.tri_verts.data(), geo.tri_verts.size() * sizeof(F4));
    const char* split = getenv("HM_BVH_SPLIT");
    const char* span = getenv("HM_BVH_SPAN");
    const char* collapse = getenv("HM_BVH_COLLAPSE");
    std::string params = std::string("v4|") + (split ? split : "-") + "|" + (span ? span : "-") + "|" + (collapse ? collapse : "-");
    h = fnv1a(h, params.data(), params.size());
    return h;
}

// Wide tree only: a scene restored from the cache has no binary tree (hm_scene_get_arrays reports 0 nodes).
bool load_bvh_cache(const std::string& path, uint64_t key, HostBvh& out) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    uint64_t hdr[6];
    bool ok = fread(hdr, 8, 6, f) == 6 && hdr[0] == kCacheMagic && hdr[1] == key;
    if (ok) {
        out.nodes.clear(); out.leaf_data.clear(); out.leaf_code.clear();
        out.wnodes.resize((size_t)hdr[2]); out.wleaf_data.resize((size_t)hdr[3]);
        out.wide_depth = (int)hdr[4];
        out.leaf_prim.resize((size_t)hdr[5]);
        ok = fread(out.wnodes.data(), sizeof(F4), out.wnodes.size(), f) == out.wnodes.size() &&
             fread(out.wleaf_data.data(), sizeof(F4), out.wleaf_data.size(), f) == out.wleaf_data.size() &&
             fread(out.leaf_prim.data(), sizeof(int), out.leaf_prim.size(), f) == out.leaf_prim.size();
    }
    fclose(f);
    if (!ok) { out.wnodes.clear(); out.wleaf_data.clear(); out.leaf_prim.clear(); }
    return ok;
}

void save_bvh_cache(const std::string& path, uint64_t key, const HostBvh& b) {
    const std::string tmp = path + ".tmp" + std::to_string((unsigned long long)(uintptr_t)&b) + std::to_string((long long)getpid());
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) return;                                        // an unwritable cache directory is not an error
    uint64_t hdr[6] = {kCacheMagic, key, b.wnodes.size(), b.wleaf_data.size(), (uint64_t)b.wide_depth, b.leaf_prim.size()};
    bool ok = fwrite(hdr, 8, 6, f) == 6 &&
              fwrite(b.wnodes.data(), sizeof(F4), b.wnodes.size(), f) == b.wnodes.size() &&
              fwrite(b.wleaf_data.data(), sizeof(F4), b.wleaf_data.size(), f) == b.wleaf_data.size() &&
              fwrite(b.leaf_prim.data(), sizeof(int), b.leaf_prim.size(), f) == b.leaf_prim.size();
    ok = fclose(f) == 0 && ok;
    if (ok) ok = rename(tmp.c_str(), path.c_str()) == 0;   // atomic: readers see a whole file or none
    if (!ok) remove(tmp.c_str());
}

// build_bvh through the cache directory named by HM_BVH_CACHE (unset: always build).
void build_bvh_cached(const HostGeometry& geo, HostBvh& out, int threads_hint) {
    const char* dir = getenv("HM_BVH_CACHE");
    if (!dir || !*dir) { build_bvh(geo, out, threads_hint); return; }
    const uint64_t key = bvh_cache_key(geo);
    char name[64];
    snprintf(name, sizeof(name), "/hm_bvh_%016llx.bin", (unsigned long long)key);
    const std::string path = std::string(dir) + name;
    if (load_bvh_cache(path, key, out)) return;
    build_bvh(geo, out, threads_hint);
    save_bvh_cache(path, key, out);
}

}  // namespace hm
