// hm_light.h — lat-long environment light (radiance lookup, pdf, importance sampling)
// and directional lights.
//
// Behavioural contract: getEnvironmentRadiance / getEnvironmentPdf /
// sampleEnvironmentLight (cuda_headers/optix_common.cuh:10-149) over the tables that
// generateEnvSamplingTables builds (scene.cpp:349-425).
//
// The reference reads its five tables through CUDA texture objects (NEAREST/CLAMP on
// normalised coordinates for the pdf/cdf tables, LINEAR/CLAMP for the RGBA32F map,
// render_hair_msnn.cu:161-198).  Here they are plain arrays in HBM read with __ldg and
// the addressing/filtering arithmetic is spelled out (CUDA C Programming Guide,
// "Texture Fetching": i = floor(x*N) clamped; bilinear taps at x*N-0.5 with the
// fractional weight held in 8 fractional bits).  That keeps every lookup bit-exact
// between the sm_100a build and a host build of this same header, and the ~200 MB of
// tables sit in B200's L2 + HBM3e without the texture path's 8-bit coordinate cache.
#pragma once
#include "hm_math.h"

namespace hm {

static constexpr int kMaxDirLights = 8;

struct EnvView {
    const float* env;   // RGBA32F, W*H*4
    const float* cpdf;  // (W+1)*H
    const float* ccdf;  // (W+1)*H
    const float* mpdf;  // H+1
    const float* mcdf;  // H+1
    // every 64th entry of each conditional-cdf row, [H][W / 64] (device only, power-of-two W <= 4096; else null): the
    // first level of the two-level search below
    const float* ccoarse;
    int W, H;
    float scale, rot_phi;
    int has_env;
    int pdf_sampling;
};

struct LightSet {
    EnvView env;
    int num_dlights;
    int num_total;
    float dl_from[kMaxDirLights][3];   // already normalised by the scene loader (scene.cpp:263)
    float dl_emit[kMaxDirLights][3];
};

#if defined(__CUDA_ARCH__)
HM_D float ldf(const float* p) { return __ldg(p); }
#else
inline float ldf(const float* p) { return *p; }
#endif

HM_HD int clamp_idx(int v, int n) { return v < 0 ? 0 : (v > n - 1 ? n - 1 : v); }

// point-sampled single-channel table, normalised coordinates, clamp addressing
HM_HD float table_fetch(const float* t, int w, int h, float xn, float yn) {
    int i = clamp_idx((int)floorf(xn * (float)w), w);
    int j = clamp_idx((int)floorf(yn * (float)h), h);
    return ldf(t + (size_t)j * w + i);
}

// bilinear RGBA fetch, normalised coordinates, clamp addressing
HM_HD V3 env_fetch(const EnvView& e, float xn, float yn) {
    float xb = xn * (float)e.W - 0.5f, yb = yn * (float)e.H - 0.5f;
    float fi = floorf(xb), fj = floorf(yb);
    int i = (int)fi, j = (int)fj;
    float a = floorf((xb - fi) * 256.f + 0.5f) / 256.f;
    float b = floorf((yb - fj) * 256.f + 0.5f) / 256.f;
    int i0 = clamp_idx(i, e.W), i1 = clamp_idx(i + 1, e.W);
    int j0 = clamp_idx(j, e.H), j1 = clamp_idx(j + 1, e.H);
    const float* p00 = e.env + 4 * ((size_t)j0 * e.W + i0);
    const float* p10 = e.env + 4 * ((size_t)j0 * e.W + i1);
    const float* p01 = e.env + 4 * ((size_t)j1 * e.W + i0);
    const float* p11 = e.env + 4 * ((size_t)j1 * e.W + i1);
    float w00 = (1 - a) * (1 - b), w10 = a * (1 - b), w01 = (1 - a) * b, w11 = a * b;
    V3 r;
    r.x = w00 * ldf(p00 + 0) + w10 * ldf(p10 + 0) + w01 * ldf(p01 + 0) + w11 * ldf(p11 + 0);
    r.y = w00 * ldf(p00 + 1) + w10 * ldf(p10 + 1) + w01 * ldf(p01 + 1) + w11 * ldf(p11 + 1);
    r.z = w00 * ldf(p00 + 2) + w10 * ldf(p10 + 2) + w01 * ldf(p01 + 2) + w11 * ldf(p11 + 2);
    return r;
}

HM_HD float spherical_phi(V3 v) {
    float p = atan2f(v.y, v.x);
    return (p < 0) ? (p + kTwoPi) : p;
}
HM_HD float spherical_theta(V3 v) { return acosf(v.z); }

static HM_HD_OUTLINE V3 env_radiance(const EnvView& e, V3 dir) {
    float theta = spherical_theta(dir);
    float phi = spherical_phi(dir) + e.rot_phi;
    if (phi > kTwoPiLoose) phi = phi - kTwoPiLoose;
    float x = phi / kTwoPi;
    float y = theta / kPi;
    return e.scale * env_fetch(e, x, y);
}

HM_HD float env_pdf_from_cell(const EnvView& e, int index_u, int index_v, float sin_theta) {
    const float width = (float)e.W, height = (float)e.H;
    const int cw = e.W + 1, mh = e.H + 1;
    float last_u = table_fetch(e.cpdf, cw, e.H, 1.f, index_v / height);
    float last_v = table_fetch(e.mpdf, mh, 1, 1.f, 0.f);
    float denom = (2.f * kPi * kPi * sin_theta) * last_u * last_v;
    float pu = table_fetch(e.cpdf, cw, e.H, index_u / width, index_v / height);
    float pv = table_fetch(e.mpdf, mh, 1, index_v / height, 0.f);
    return denom ? (pu * pv) / denom : 0.f;
}

HM_HD float env_pdf(const EnvView& e, V3 wi) {
    if (!e.pdf_sampling) return 1.f / (4.f * kPi);
    float theta = spherical_theta(wi);
    float phi = spherical_phi(wi) + e.rot_phi;
    if (phi > kTwoPiLoose) phi = phi - kTwoPiLoose;
    float u = phi / (2.f * kPi);
    float v = theta / kPi;
    int index_u = clampi((int)(u * e.W), 0, e.W - 1);
    int index_v = clampi((int)(v * e.H), 0, e.H - 1);
    return env_pdf_from_cell(e, index_u, index_v, sinf(theta));
}

// std::lower_bound over a point-sampled cdf row, as the reference spells it
HM_HD int cdf_lower_bound(float u, const float* t, int w, int h, float yn, float size) {
    int first = 0;
    int count = (int)size;
    while (count > 0) {
        int step = count >> 1;
        int middle = first + step;
        if (table_fetch(t, w, h, middle / size, yn) < u) {
            first = middle + 1;
            count -= step + 1;
        } else {
            count = step;
        }
    }
    return first - 1 > 0 ? first - 1 : 0;
}

// The same index for tables whose `size` is a power of two <= 4096 (every shipped map): there middle / size is exact and
// (middle / size) * (size + 1) = middle + middle / size needs at most 24 bits, so table_fetch's floor() returns exactly
// `middle` — the normalised-coordinate arithmetic can be dropped.  Two levels: lower_bound over every 64th entry
// (`coarse`, 256 contiguous bytes per row), then over the 63 entries of the bracket it names.  A lower_bound of a sorted
// row is unique, so the result equals cdf_lower_bound's whatever the probe order; what changes is the memory behaviour —
// the one-level search makes ~7 dependent L2 round trips per row (its first probes are 8 KB, 4 KB, 2 KB ... apart), this
// one ~3.  (k_shade: long-scoreboard 14 of 27 stall cycles per issue, this loop its hottest line; profiles/r2m_shade.*)
HM_HD int cdf_lower_bound_two_level(float u, const float* row, const float* coarse, int size) {
    // first coarse entry k in [0, K) with !(row[64 k] < u); K if none
    const int K = size >> 6;
    int first = 0, count = K;
    while (count > 0) {
        const int step = count >> 1, middle = first + step;
        if (ldf(coarse + middle) < u) { first = middle + 1; count -= step + 1; }
        else count = step;
    }
    int F = 0;
    if (first > 0) {
        // row[64 (first - 1)] < u and (first == K or row[64 first] >= u): the answer lies in (64 (first - 1), 64 first]
        int lo = 64 * (first - 1) + 1;
        int cnt = (first < K ? 64 * first : size) - lo;
        while (cnt > 0) {
            const int step = cnt >> 1, middle = lo + step;
            if (ldf(row + middle) < u) { lo = middle + 1; cnt -= step + 1; }
            else cnt = step;
        }
        F = lo;
    }
    return F - 1 > 0 ? F - 1 : 0;
}
HM_HD bool cdf_direct_ok(int size) { return size >= 64 && size <= 4096 && (size & (size - 1)) == 0; }

// (A 4-ary variant of this search — half the dependent round trips, three probes each — was measured on the B200
// against the real scenes/curly frame: k_shade 1.77 vs 1.59 ms per frame, profiles/r2d_sweep_knobs.txt `env4`.
// The extra probes cost more than the saved round trips; it was removed.)
HM_HD int cdf_search(float u, const float* t, int w, int h, float yn, float size) { return cdf_lower_bound(u, t, w, h, yn, size); }

HM_HD V3 uniform_sample_sphere(float u0, float u1) {
    float z = 1 - 2 * u0;
    float r = sqrtf(fmaxf(0.f, 1.f - z * z));
    float phi = 2 * kPi * u1;
    return normalize(V3(r * cosf(phi), r * sinf(phi), z));
}

// u0 = first draw (pairs with the conditional/u axis), u1 = second draw.
// Returns radiance; wi and pdf by reference.
static HM_HD_OUTLINE V3 env_sample(const EnvView& e, float u0, float u1, V3& wi, float& pdf) {
    if (!e.pdf_sampling) {
        wi = uniform_sample_sphere(u0, u1);
        pdf = 1.f / (4.f * kPi);
        return env_radiance(e, wi);
    }
    const float width = (float)e.W, height = (float)e.H;
    const int cw = e.W + 1, mh = e.H + 1;

    int index_v = cdf_search(u1, e.mcdf, mh, 1, 0.f, height);
    float cdf_v = table_fetch(e.mcdf, mh, 1, index_v / height, 0.f);
    float cdf_next_v = table_fetch(e.mcdf, mh, 1, (index_v + 1) / height, 0.f);
    float dv = (cdf_next_v - u1) / (cdf_next_v - cdf_v);
    float v = (index_v + dv) / height;

    int index_u;
    if (e.ccoarse && cdf_direct_ok(e.W)) {
        // row index exactly as table_fetch derives it from index_v / height
        const int j = clamp_idx((int)floorf((index_v / height) * (float)e.H), e.H);
        index_u = cdf_lower_bound_two_level(u0, e.ccdf + (size_t)j * cw, e.ccoarse + (size_t)j * (e.W >> 6), e.W);
    } else {
        index_u = cdf_search(u0, e.ccdf, cw, e.H, index_v / height, width);
    }
    float cdf_u = table_fetch(e.ccdf, cw, e.H, index_u / width, index_v / height);
    float cdf_next_u = table_fetch(e.ccdf, cw, e.H, (index_u + 1) / width, index_v / height);
    float du = (cdf_next_u - u0) / (cdf_next_u - cdf_u);
    float u = (index_u + du) / width;

    V3 rad = e.scale * env_fetch(e, u, v);

    float theta = kPi * v;
    float sin_theta = sinf(theta);
    pdf = env_pdf_from_cell(e, index_u, index_v, sin_theta);

    float phi = 2.f * kPi * u - e.rot_phi;
    if (phi < 0) phi = kTwoPiLoose + phi;
    wi = V3(sin_theta * cosf(phi), sin_theta * sinf(phi), cosf(theta));
    return rad;
}

}  // namespace hm
