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
#include <cstring>
#include <thread>

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
    std::vector<int>& leaf_first;    // per-leaf bookkeeping is implicit: leaves index `order`
    float cost_isect;
    int max_leaf;
    int spawn_depth;

    static constexpr int kBins = 16;

    Box bounds_of(int b, int e) const {
        Box bx; bx.reset();
        for (int i = b; i < e; ++i) bx.grow(prims[order[i]].box);
        return bx;
    }

    // Returns the child code for range [b,e) with bounds `bx`.
    int build_range(int b, int e, const Box& bx, int depth) {
        int n = e - b;
        if (n <= 1) return make_leaf(b, e);

        Box cb; cb.reset();
        for (int i = b; i < e; ++i) cb.grow(prims[order[i]].cen);

        float best_cost = FLT_MAX;
        int best_axis = -1, best_bin = -1;
        float parent_area = std::max(bx.half_area(), 1e-30f);

        for (int axis = 0; axis < 3; ++axis) {
            float ext = cb.hi[axis] - cb.lo[axis];
            if (!(ext > 0.f)) continue;
            Box bin_box[kBins];
            int bin_cnt[kBins];
            for (int k = 0; k < kBins; ++k) { bin_box[k].reset(); bin_cnt[k] = 0; }
            float scale = kBins / ext;
            for (int i = b; i < e; ++i) {
                const PrimRef& p = prims[order[i]];
                int k = std::min(kBins - 1, std::max(0, (int)((p.cen[axis] - cb.lo[axis]) * scale)));
                bin_box[k].grow(p.box);
                bin_cnt[k]++;
            }
            float right_area[kBins];
            int right_cnt[kBins];
            Box acc; acc.reset();
            int cnt = 0;
            for (int k = kBins - 1; k > 0; --k) {
                acc.grow(bin_box[k]); cnt += bin_cnt[k];
                right_area[k] = acc.half_area(); right_cnt[k] = cnt;
            }
            acc.reset(); cnt = 0;
            for (int k = 0; k < kBins - 1; ++k) {
                acc.grow(bin_box[k]); cnt += bin_cnt[k];
                if (cnt == 0 || right_cnt[k + 1] == 0) continue;
                float c = 1.f + cost_isect * (acc.half_area() * cnt + right_area[k + 1] * right_cnt[k + 1]) / parent_area;
                if (c < best_cost) { best_cost = c; best_axis = axis; best_bin = k; }
            }
        }

        float leaf_cost = cost_isect * n;
        if (n <= max_leaf && (best_axis < 0 || leaf_cost <= best_cost)) return make_leaf(b, e);

        int mid;
        if (best_axis < 0) {
            // all centroids coincide: split by count
            mid = b + n / 2;
        } else {
            float ext = cb.hi[best_axis] - cb.lo[best_axis];
            float scale = kBins / ext;
            float lo = cb.lo[best_axis];
            int axis = best_axis, bin = best_bin;
            auto it = std::partition(order.begin() + b, order.begin() + e, [&](int pi) {
                int k = std::min(kBins - 1, std::max(0, (int)((prims[pi].cen[axis] - lo) * scale)));
                return k <= bin;
            });
            mid = (int)(it - order.begin());
            if (mid == b || mid == e) mid = b + n / 2;
        }

        Box lb = bounds_of(b, mid), rb = bounds_of(mid, e);
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

    int make_leaf(int b, int e) {
        // leaves index `order` directly: slot range [b, e)
        int count = e - b;
        if (count > kMaxLeaf) {
            // cannot happen with max_leaf <= kMaxLeaf except for coincident centroids;
            // split evenly until it fits
            int mid = b + count / 2;
            Box lb = bounds_of(b, mid), rb = bounds_of(mid, e);
            int me = next_node.fetch_add(1);
            int cl = make_leaf(b, mid), cr = make_leaf(mid, e);
            NodeRaw& nd = nodes[me];
            for (int k = 0; k < 3; ++k) { nd.q[k] = lb.lo[k]; nd.q[3 + k] = lb.hi[k]; nd.q[6 + k] = rb.lo[k]; nd.q[9 + k] = rb.hi[k]; }
            nd.c0 = cl; nd.c1 = cr; nd.pad0 = nd.pad1 = 0;
            return me;
        }
        return ~((b << 3) | (count - 1));
    }
};

void segment_bounds(const F4* cps, int cp0, PrimRef& pr) {
    // Bezier hull of the Catmull-Rom span + radius (max over the 4 control radii)
    const F4& k0 = cps[cp0], &k1 = cps[cp0 + 1], &k2 = cps[cp0 + 2], &k3 = cps[cp0 + 3];
    float r = std::max(std::max(k0.w, k1.w), std::max(k2.w, k3.w));
    float b0[3] = {k1.x, k1.y, k1.z};
    float b3[3] = {k2.x, k2.y, k2.z};
    float b1[3] = {k1.x + (k2.x - k0.x) / 6.f, k1.y + (k2.y - k0.y) / 6.f, k1.z + (k2.z - k0.z) / 6.f};
    float b2[3] = {k2.x - (k3.x - k1.x) / 6.f, k2.y - (k3.y - k1.y) / 6.f, k2.z - (k3.z - k1.z) / 6.f};
    pr.box.reset();
    pr.box.grow(b0); pr.box.grow(b1); pr.box.grow(b2); pr.box.grow(b3);
    // pad by radius plus a relative epsilon so fp32 round-off in the ray-space
    // solver can never place a hit outside its own box
    for (int k = 0; k < 3; ++k) {
        float pad = r + 1e-5f * std::max(fabsf(pr.box.lo[k]), fabsf(pr.box.hi[k])) + 1e-6f;
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

}  // namespace

void build_bvh(const HostGeometry& geo, HostBvh& out, int threads_hint) {
    const int ns = (int)geo.seg_cp.size();
    const int nt = (int)(geo.tri_verts.size() / 3);
    const int n = ns + nt;
    out.nodes.clear(); out.leaf_data.clear(); out.leaf_code.clear(); out.leaf_prim.clear();
    if (n == 0) return;

    std::vector<PrimRef> prims(n);
    {
        unsigned hw = threads_hint > 0 ? (unsigned)threads_hint : std::max(1u, std::thread::hardware_concurrency());
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < hw; ++t) {
            pool.emplace_back([&, t]() {
                for (int i = (int)t; i < n; i += (int)hw) {
                    if (i < ns) segment_bounds(geo.cps.data(), geo.seg_cp[i], prims[i]);
                    else triangle_bounds(geo.tri_verts.data(), i - ns, prims[i]);
                }
            });
        }
        for (auto& th : pool) th.join();
    }

    std::vector<int> order(n);
    for (int i = 0; i < n; ++i) order[i] = i;
    std::vector<NodeRaw> nodes((size_t)std::max(2 * n, 2));
    std::vector<int> dummy;

    unsigned hw = threads_hint > 0 ? (unsigned)threads_hint : std::max(1u, std::thread::hardware_concurrency());
    int spawn_depth = 0;
    while ((1u << spawn_depth) < 2 * hw && spawn_depth < 8) spawn_depth++;

    Builder bld{prims, order, nodes, {}, dummy, 2.0f, 4, spawn_depth};
    bld.next_node.store(1);
    Box root; root.reset();
    for (int i = 0; i < n; ++i) root.grow(prims[i].box);

    // root node is index 0: build it by hand so it is always an inner node
    int code;
    {
        // temporarily let the builder allocate; then move the produced top node to slot 0
        code = bld.build_range(0, n, root, 0);
    }
    int used = bld.next_node.load();
    if (code < 0) {
        // whole scene fits one leaf: child0 = leaf, child1 = empty
        NodeRaw& nd = nodes[0];
        for (int k = 0; k < 3; ++k) { nd.q[k] = root.lo[k]; nd.q[3 + k] = root.hi[k]; nd.q[6 + k] = FLT_MAX; nd.q[9 + k] = -FLT_MAX; }
        nd.c0 = code; nd.c1 = code; nd.pad0 = nd.pad1 = 0;
        // make child1 unreachable
        nd.q[6] = nd.q[7] = nd.q[8] = FLT_MAX; nd.q[9] = nd.q[10] = nd.q[11] = -FLT_MAX;
        used = 1;
    } else {
        // `code` is the index of the top node; swap it into slot 0 (slot 0 is unused so far)
        nodes[0] = nodes[code];
        // slot `code` is now dead; harmless (never referenced)
    }

    // Re-layout in depth-first order for locality and to drop dead slots.
    std::vector<NodeRaw> packed;
    packed.reserve(used);
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
                } else newc[k] = c[k];
            }
            packed[it.dst].c0 = newc[0];
            packed[it.dst].c1 = newc[1];
            // push right first so the left subtree is laid out right after its parent pair
            if (c[1] >= 0) st.push_back({c[1], newc[1]});
            if (c[0] >= 0) st.push_back({c[0], newc[0]});
        }
    }

    out.nodes.resize(packed.size() * 4);
    memcpy(out.nodes.data(), packed.data(), packed.size() * sizeof(NodeRaw));
    out.leaf_code.resize(n);
    out.leaf_prim.resize(n);
    out.leaf_data.resize(4 * (size_t)n);
    for (int i = 0; i < n; ++i) {
        int p = order[i];
        out.leaf_prim[i] = p;
        out.leaf_code[i] = p < ns ? geo.seg_cp[p] : ((p - ns) | kTriTag);
        F4* dst = out.leaf_data.data() + 4 * (size_t)i;
        if (p < ns) {
            const F4* cp = geo.cps.data() + geo.seg_cp[p];
            dst[0] = cp[0]; dst[1] = cp[1]; dst[2] = cp[2]; dst[3] = cp[3];
            if (dst[3].w < 0.f) dst[3].w = 0.f;
        } else {
            const F4* tv = geo.tri_verts.data() + 3 * (size_t)(p - ns);
            dst[0] = tv[0]; dst[1] = tv[1]; dst[2] = tv[2];
            dst[3] = F4{0.f, 0.f, 0.f, -1.f};
        }
    }
}

}  // namespace hm
