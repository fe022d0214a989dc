// hm_scene_util.cpp — start-up derivations shared by every renderer flavour: environment
// importance tables, scene bounds/scale, camera basis.  Load-time CPU work.
#include <algorithm>
#include <cmath>
#include <limits>
#include <random>
#include <mutex>
#include <thread>

#include "hm_host.h"

namespace hm {

// Same tables as generateEnvSamplingTables (scene.cpp:349-425): per-row conditional
// pdf/cdf over W+1 entries (last pdf entry holds the row total), marginal over H+1.
// Rows are independent, so they are built in parallel; each row's running sum is
// sequential to keep the float accumulation order.
void build_env_tables(HostScene& s) {
    const int W = s.env_w, H = s.env_h, cw = W + 1;
    s.cpdf.assign((size_t)cw * H, 0.f);
    s.ccdf.assign((size_t)cw * H, 0.f);
    s.mpdf.assign(H + 1, 0.f);
    s.mcdf.assign(H + 1, 0.f);
    if (W <= 0 || H <= 0) return;
    const float* env = s.env.data();
    auto row = [&](int y) {
        const float sin_theta = sinf(3.14159f * (y + 0.5f) / H);     // == env_row_sines()[y]
        float* pdf = s.cpdf.data() + (size_t)y * cw;
        float* cdf = s.ccdf.data() + (size_t)y * cw;
        auto avg = [&](int x) {
            const float* p = env + 4 * ((size_t)y * W + x);
            return (p[0] + p[1] + p[2]) * (1.0f / 3.0f);
        };
        pdf[0] = avg(0) * sin_theta;
        cdf[0] = 0.f;
        for (int x = 1; x < W; ++x) {
            pdf[x] = avg(x) * sin_theta;
            cdf[x] = cdf[x - 1] + pdf[x - 1] / W;
        }
        const float total = cdf[W - 1] + pdf[W - 1] / W;
        pdf[W] = total;
        if (total > 0.f) {
            const float inv = 1.0f / total;
            for (int x = 1; x < W; ++x) cdf[x] *= inv;
        }
        cdf[W] = 1.0f;
    };
    unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < hw; ++t)
        pool.emplace_back([&, t]() { for (int y = (int)t; y < H; y += (int)hw) row(y); });
    for (auto& th : pool) th.join();

    s.mpdf[0] = s.cpdf[W];
    s.mcdf[0] = 0.f;
    for (int i = 1; i < H; ++i) {
        s.mpdf[i] = s.cpdf[(size_t)i * cw + W];
        s.mcdf[i] = s.mcdf[i - 1] + s.mpdf[i - 1] / H;
    }
    float total = s.mcdf[H - 1] + s.mpdf[H - 1] / H;
    s.mpdf[H] = total;
    if (total > 0.f)
        for (int i = 1; i < H; ++i) s.mcdf[i] /= total;
    s.mcdf[H] = 1.0f;
}

void env_row_sines(int H, std::vector<float>& out) {
    out.resize(H > 0 ? H : 0);
    for (int y = 0; y < H; ++y) out[y] = sinf(3.14159f * (y + 0.5f) / H);
}

void ensure_env_tables(const HostScene& cs) {
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    HostScene& s = const_cast<HostScene&>(cs);      // a cache: the tables are a pure function of the map
    if (s.has_env && s.mcdf.empty()) build_env_tables(s);
}

// RenderWindow_HairMSNN::fetchSceneSamples (render_hair_msnn.cu:34-97), the part the TRAIN_DATA_GEN
// pass reads: ~num_samples points on the strands, each segment getting
// int(len / total_len * num_samples) of them (at least one), len = control-polygon length
// (render_hair_msnn.cu:359-368); positions at uniform random curve parameters.  The reference seeds
// std::default_random_engine from the wall clock; `seed` makes the set reproducible.
void build_scene_samples(const HostGeometry& g, int num_samples, unsigned seed, std::vector<float>& points3) {
    points3.clear();
    const size_t ns = g.seg_cp.size();
    std::vector<float> len(ns);
    float total = 0.f;
    auto dist = [](const F4& a, const F4& b) {
        V3 d = V3(a.x, a.y, a.z) - V3(b.x, b.y, b.z);
        return length(d);
    };
    for (size_t i = 0; i < ns; ++i) {
        const F4* c = g.cps.data() + g.seg_cp[i];
        float l = dist(c[0], c[1]) + dist(c[2], c[1]) + dist(c[3], c[2]);
        total += l;
        len[i] = l;
    }
    std::default_random_engine gen(seed);
    points3.reserve(3 * (size_t)(num_samples + ns));
    for (size_t i = 0; i < ns; ++i) {
        int k = (int)(len[i] / total * (float)num_samples);
        if (k == 0) k = 1;
        const F4* c = g.cps.data() + g.seg_cp[i];
        CubicSeg seg;
        seg.from_catmull_rom(f4_to_v4(c[0]), f4_to_v4(c[1]), f4_to_v4(c[2]), f4_to_v4(c[3]));
        for (int j = 0; j < k; ++j) {
            float u = std::generate_canonical<float, std::numeric_limits<float>::digits>(gen);
            V4 p = seg.pos4(u);
            points3.push_back(p.x); points3.push_back(p.y); points3.push_back(p.z);
        }
    }
}

void finalize_geometry(HostScene& s) {
    HostGeometry& g = s.geo;
    // hair bounds start at the origin (headers/model.h:93-94) and cover the REAL points
    // only; callers that build HostGeometry by hand have filled hair_min/hair_max, the
    // loaders do it while extracting strands.
    float d[3] = {g.hair_max[0] - g.hair_min[0], g.hair_max[1] - g.hair_min[1], g.hair_max[2] - g.hair_min[2]};
    g.hair_scale = g.seg_cp.empty() ? 0.f : sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    float surf_scale = 0.f;
    const size_t nt = g.tri_verts.size() / 3;
    if (nt) {
        // window bounds start at max = 1e-30, min = 1e30 (headers/render_hair_msnn.h:94) and
        // grow by the FIRST corner of each triangle (model.cpp:309-316)
        float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {1e-30f, 1e-30f, 1e-30f};
        for (size_t t = 0; t < nt; ++t) {
            const F4& v = g.tri_verts[3 * t];
            const float p[3] = {v.x, v.y, v.z};
            for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], p[k]); hi[k] = std::max(hi[k], p[k]); }
        }
        for (int k = 0; k < 3; ++k) { g.mesh_min[k] = lo[k]; g.mesh_max[k] = hi[k]; }
        float e[3] = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
        surf_scale = sqrtf(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
    }
    g.scene_scale = std::max(g.hair_scale, surf_scale);
}

namespace {
struct H3 { float x, y, z; };
inline H3 sub(H3 a, H3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline H3 mul(float s, H3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline float dot3(H3 a, H3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline H3 cross3(H3 a, H3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline H3 norm3(H3 a) { float r = 1.f / sqrtf(dot3(a, a)); return {a.x * r, a.y * r, a.z * r}; }
}  // namespace

void camera_basis(const HostScene& s, int W, int H, float pos[3], float d00[3], float du[3], float dv[3]) {
    H3 from{s.cam_from[0], s.cam_from[1], s.cam_from[2]}, at{s.cam_to[0], s.cam_to[1], s.cam_to[2]};
    H3 up{s.cam_up[0], s.cam_up[1], s.cam_up[2]};
    // Camera::setOrientation + forceUpFrame
    const float deg = (float)(M_PI / 180.f);
    float fovy_deg = acosf(s.cos_fovy) / deg;
    bool same = from.x == at.x && from.y == at.y && from.z == at.z;
    H3 vz = same ? H3{0, 0, 1} : mul(-1.f, norm3(sub(at, from)));
    H3 vx = cross3(up, vz);
    if (dot3(vx, vx) < 1e-8f) vx = H3{0, 1, 0};
    else vx = norm3(vx);
    H3 vy = norm3(cross3(vz, vx));
    if (!(fabsf(dot3(vz, up)) < 1e-6f)) {
        vx = norm3(cross3(up, vz));
        vy = norm3(cross3(vz, vx));
    }
    // cameraChanged(): lookAt = position - vz, lookUp = vy, cosFovy through degrees
    H3 look_at = sub(from, vz);
    float cos_fovy = cosf(fovy_deg * deg);
    H3 c00 = norm3(sub(look_at, from));
    float aspect = W / float(H);
    H3 cdu = mul(cos_fovy * aspect, norm3(cross3(c00, vy)));
    H3 cdv = mul(cos_fovy, norm3(cross3(cdu, c00)));
    c00 = sub(c00, mul(0.5f, cdu));
    c00 = sub(c00, mul(0.5f, cdv));
    pos[0] = from.x; pos[1] = from.y; pos[2] = from.z;
    d00[0] = c00.x; d00[1] = c00.y; d00[2] = c00.z;
    du[0] = cdu.x; du[1] = cdu.y; du[2] = cdu.z;
    dv[0] = cdv.x; dv[1] = cdv.y; dv[2] = cdv.z;
}

}  // namespace hm
