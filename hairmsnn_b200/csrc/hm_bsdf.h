// hm_bsdf.h — fibre (Chiang "Disney" hair) and head-surface scattering models.
//
// Behavioural contract:
//   * hair_eval   == disney_hair          (cuda_headers/disney_hair.cuh:185-266) : returns f*cos and pdf
//   * hair_sample == sample_disney_hair   (cuda_headers/disney_hair.cuh:276-386)
//   * HairLobes::setup == setupHairShading (cuda_headers/disney_hair.cuh:25-41)
//   * surf_eval / surf_sample / surf_pdf == frostbite_GGX / sample_GGX / pdf_GGX
//                                           (cuda_headers/frostbite_anisotropic.cuh:84-158)
//
// Design differences from the reference: the lobe constants are scene-uniform, so
// they are computed once on the host (HairLobes) instead of per hit; the three
// tilted lobes run through one loop over a small table; attenuation terms are shared
// between eval and the lobe-selection pdf.  Arithmetic (operation order, float vs
// double promotion of the reference's literals) is kept so results agree to fp32
// round-off:
//   - LogI0's large-x branch promotes to double through the literal 0.5
//   - asin(h) is NOT clamped (|h| can exceed 1 by an ulp -> NaN, which callers scrub
//     exactly as the reference does)
//   - the residual lobe (p == 3) is *sampled* with longitudinal variance v[3], which
//     in the reference aliases the next struct member `s` (float v[3]; float s;
//     common.cuh:39-40).  Reproduced on purpose via v_sample[3] = s.
#pragma once
#include "hm_math.h"

namespace hm {

struct HairLobes {
    V3 sigma_a;
    float gain[4];      // R, TT, TRT, TRRT
    float v[3];         // longitudinal variances
    float s;            // azimuthal logistic scale
    float sin2k[3], cos2k[3];
    float v_sample[4];  // v[0..2], s  (see header note)
    float radius_unused;
    // Scene-uniform subexpressions of Mp / Np / the lobe sampler, evaluated once here with the reference's own
    // operations instead of at every evaluation (a vertex evaluates the model up to three times, each with four Mp and
    // three Np).  What the kernels gain: an IEEE division is ~10 instructions plus a slow-path branch, and k_shade was
    // spending 44 % of its instructions in them (profiles/r2b_main.summary.txt: 6.2 M FCHK per launch).  Products with
    // these reciprocals differ from the reference's quotients by at most one ulp per operation.
    float inv_v[3];        // 1 / v[p]
    float mp_log_term[3];  // logf(1 / (2 v[p]))                       (Mp, v <= 0.1 branch)
    float mp_inv_den[3];   // 1 / (sinhf(1 / v[p]) * 2 * v[p])         (Mp, v > 0.1 branch)
    float inv_s;           // 1 / s
    float az_inv_norm;     // 1 / (logistic_cdf(pi, s) - logistic_cdf(-pi, s))
    float az_cdf_lo;       // logistic_cdf(-pi, s)
    float az_cdf_span;     // logistic_cdf(pi, s) - logistic_cdf(-pi, s)
    float smp_exp[4];      // expf(-2 / v_sample[p])

    void setup(float beta_m, float beta_n, float alpha_rad) {
        v[0] = sqr(0.726f * beta_m + 0.812f * sqr(beta_m) + 3.7f * powf(beta_m, 20.f));
        v[1] = (float)(.25 * v[0]);
        v[2] = 4 * v[0];
        s = 0.626657069f * (0.265f * beta_n + 1.194f * sqr(beta_n) + 5.372f * powf(beta_n, 22.f));
        sin2k[0] = sinf(alpha_rad);
        cos2k[0] = safe_sqrt(1 - sqr(sin2k[0]));
        for (int i = 1; i < 3; ++i) {
            sin2k[i] = 2 * cos2k[i - 1] * sin2k[i - 1];
            cos2k[i] = sqr(cos2k[i - 1]) - sqr(sin2k[i - 1]);
        }
        v_sample[0] = v[0]; v_sample[1] = v[1]; v_sample[2] = v[2]; v_sample[3] = s;
        radius_unused = 0.f;
        const float pi = 3.1415926f;   // kPi (utils.cuh:10)
        for (int i = 0; i < 3; ++i) {
            inv_v[i] = 1 / v[i];
            mp_log_term[i] = logf(1 / (2 * v[i]));
            mp_inv_den[i] = 1 / (sinhf(1 / v[i]) * 2 * v[i]);
        }
        inv_s = 1 / s;
        const float hi = 1 / (1 + expf(-pi / s)), lo = 1 / (1 + expf(pi / s));
        az_cdf_lo = lo;
        az_cdf_span = hi - lo;
        az_inv_norm = 1 / (hi - lo);
        for (int i = 0; i < 4; ++i) smp_exp[i] = expf(-2.f / v_sample[i]);
    }
};

namespace hairdetail {

static constexpr float kEta = 1.55f;

HM_HD float schlick(float cos_theta) {
    const float f0 = sqr(1 - kEta) / sqr(1 + kEta);
    float xd = (1 - cos_theta);
    return f0 + (1 - f0) * xd * xd * xd * xd * xd;
}

// 10-term power series of the modified Bessel function (disney_hair.cuh:53-67); the denominators are
// 4^i * (i!)^2 evaluated the way the reference's mixed int/float expression does.  The reference divides by them;
// here each term is multiplied by the correctly rounded reciprocal (exact for the first three, a power of two each;
// at most one ulp per term otherwise) — ten IEEE divisions per call, twelve calls per vertex, were a third of k_shade.
HM_HD float bessel_i0(float x) {
    const float rden[10] = {
        1.f / (1.f * (1.f * 1.f)),
        1.f / (4.f * (1.f * 1.f)),
        1.f / (16.f * (2.f * 2.f)),
        1.f / (64.f * (6.f * 6.f)),
        1.f / (256.f * (24.f * 24.f)),
        1.f / (1024.f * (120.f * 120.f)),
        1.f / (4096.f * (720.f * 720.f)),
        1.f / (16384.f * (5040.f * 5040.f)),
        1.f / (65536.f * (40320.f * 40320.f)),
        1.f / (262144.f * (362880.f * 362880.f))};
    float val = 0.f, x2i = 1.f;
    const float xx = x * x;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        val += x2i * rden[i];
        x2i *= xx;
    }
    return val;
}

HM_HD float log_bessel_i0(float x) {
    if (x > 12)
        return (float)(x + 0.5 * (-logf(kTwoPi) + logf(1 / x) + 1 / (8 * x)));
    return logf(bessel_i0(x));
}

// longitudinal scattering M_p
// L.v[k] is the variance; the quotients by it use the reciprocals of HairLobes::setup (same operation order as
// disney_hair.cuh:77-88).
static HM_HD_OUTLINE float longitudinal(const HairLobes& L, int k, float cos_i, float cos_o, float sin_i, float sin_o) {
    // a and b stay true quotients: for small v they reach several hundred and sit inside an exponential, where one ulp
    // of them is 1e-4 of the result
    float a = div_exact(cos_i * cos_o, L.v[k]);
    float b = div_exact(sin_i * sin_o, L.v[k]);
    if (L.v[k] <= 0.1f)
        return expf(log_bessel_i0(a) - b - L.inv_v[k] + 0.6931f + L.mp_log_term[k]);
    return (expf(-b) * bessel_i0(a)) * L.mp_inv_den[k];
}

HM_HD float logistic(float x, float s, float inv_s) {
    x = fabsf(x);
    const float e = expf(-x * inv_s);
    return e / (s * sqr(1 + e));
}
HM_HD float logistic_cdf(float x, float s) { return 1 / (1 + expf(-x / s)); }

HM_HD float net_phi(int p, float gamma_o, float gamma_t) {
    return 2 * p * gamma_t - 2 * gamma_o + p * kPi;
}

// azimuthal scattering N_p (trimmed logistic on [-pi, pi])
HM_HD float azimuthal(const HairLobes& L, float phi, int p, float gamma_o, float gamma_t) {
    float dphi = phi - net_phi(p, gamma_o, gamma_t);
    while (dphi > kPi) dphi -= kTwoPi;
    while (dphi < -kPi) dphi += kTwoPi;
    return logistic(dphi, L.s, L.inv_s) * L.az_inv_norm;      // trimmed to [-pi, pi]: the norm is scene-uniform
}

// SampleTrimmedLogistic on [-pi, pi] (disney_hair.cuh:131-139) with the two cdf values from HairLobes::setup
HM_HD float sample_trimmed_logistic(const HairLobes& L, float u) {
    float x = -L.s * logf(1 / (u * L.az_cdf_span + L.az_cdf_lo) - 1);
    return clampf(x, -kPi, kPi);
}

// Everything that depends only on the outgoing direction and the azimuthal offset h.
struct FibreGeom {
    float sin_o, cos_o, phi_o;
    float gamma_o, cos_gamma_o, gamma_t;
    V3 T;            // single-pass transmittance
    float fresnel;
    float ap_pdf[4]; // lobe selection probabilities
    V3 ap[4];        // attenuations A_0..A_2, residual
};

static HM_HD_OUTLINE void fibre_geom(const HairLobes& L, V3 wo, float h, FibreGeom& g) {
    g.sin_o = wo.x;
    g.cos_o = safe_sqrt(1 - sqr(g.sin_o));
    g.phi_o = atan2f(wo.z, wo.y);

    float sin_t = g.sin_o / kEta;
    float cos_t = safe_sqrt(1 - sqr(sin_t));

    g.gamma_o = asinf(h);  // deliberately unclamped
    g.cos_gamma_o = cosf(g.gamma_o);

    float etap = sqrtf(kEta * kEta - sqr(g.sin_o)) / g.cos_o;
    float sin_gamma_t = h / etap;
    float cos_gamma_t = safe_sqrt(1 - sqr(sin_gamma_t));
    g.gamma_t = safe_asin(sin_gamma_t);

    float fac = (2.f * cos_gamma_t / cos_t);
    g.T = V3(expf(-L.sigma_a.x * fac), expf(-L.sigma_a.y * fac), expf(-L.sigma_a.z * fac));
    g.fresnel = schlick(g.cos_o * g.cos_gamma_o);

    const float f = g.fresnel;
    g.ap[0] = V3(f);
    g.ap[1] = sqr(1 - f) * g.T;
    g.ap[2] = sqr(1 - f) * g.T * g.T * f;
    g.ap[3] = g.ap[2] * f * g.T / (V3(1.f) - g.T * f);

    float m[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) m[i] = (g.ap[i].x + g.ap[i].y + g.ap[i].z) / 3.f;
    float sum = m[0] + m[1] + m[2] + m[3];
#pragma unroll
    for (int i = 0; i < 4; ++i) g.ap_pdf[i] = m[i] / sum;
}

// cuticle tilt of the outgoing angle for lobe p (R uses 2*alpha with a flipped sign,
// TT uses alpha, TRT uses 4*alpha).
HM_HD void tilt(const HairLobes& L, int p, float sin_o, float cos_o, float& sin_op, float& cos_op) {
    if (p == 0) {
        sin_op = sin_o * L.cos2k[1] - cos_o * L.sin2k[1];
        cos_op = cos_o * L.cos2k[1] + sin_o * L.sin2k[1];
    } else if (p == 1) {
        sin_op = sin_o * L.cos2k[0] + cos_o * L.sin2k[0];
        cos_op = cos_o * L.cos2k[0] - sin_o * L.sin2k[0];
    } else if (p == 2) {
        sin_op = sin_o * L.cos2k[2] + cos_o * L.sin2k[2];
        cos_op = cos_o * L.cos2k[2] - sin_o * L.sin2k[2];
    } else {
        sin_op = sin_o;
        cos_op = cos_o;
    }
}

// Outlined on the device: a vertex evaluates the model up to three times (light sample, BSDF
// sample, continuation) and the inlined copies made k_shade 16.6 k instructions — ncu showed it
// stalled on instruction fetch (stall_no_instruction 9 of 19 cycles per issue).
static HM_HD_OUTLINE V3 eval_with_geom(const HairLobes& L, const FibreGeom& g, V3 wi, float* pdf) {
    float sin_i = wi.x;
    float cos_i = safe_sqrt(1 - sqr(sin_i));
    float phi_i = atan2f(wi.z, wi.y);
    float phi = phi_i - g.phi_o;

    V3 f(0.f);
    float p_acc = 0.f;
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        float sin_op, cos_op;
        tilt(L, p, g.sin_o, g.cos_o, sin_op, cos_op);
        cos_op = fabsf(cos_op);
        float mp = longitudinal(L, p, cos_i, cos_op, sin_i, sin_op);
        float np = azimuthal(L, phi, p, g.gamma_o, g.gamma_t);
        f += ((L.gain[p] * mp) * g.ap[p]) * np;
        p_acc = p_acc + mp * g.ap_pdf[p] * np;
    }
    float mp_res = longitudinal(L, 2, cos_i, g.cos_o, sin_i, g.sin_o);
    float np_res = 1.f / (2.f * kPi);
    f += ((L.gain[3] * mp_res) * g.ap[3]) * np_res;
    p_acc = p_acc + mp_res * g.ap_pdf[3] * np_res;

    *pdf = p_acc;
    return f;
}

}  // namespace hairdetail

// f * cos and pdf for local directions (x = fibre tangent).  h = dot(Y, n).
HM_HD V3 hair_eval(const HairLobes& L, V3 wo_local, V3 wi_local, float h, float* pdf) {
    hairdetail::FibreGeom g;
    hairdetail::fibre_geom(L, wo_local, h, g);
    return hairdetail::eval_with_geom(L, g, wi_local, pdf);
}

// Draws wi_local from the lobe mixture with u = (lobe, theta, phi-of-theta, dphi).
// The caller maps it to world space, renormalises, and evaluates with hair_eval on
// the *local* direction returned here (as the reference does).
// Same two entry points with the per-vertex part (hairdetail::FibreGeom: everything that depends on
// wo and h only) computed once by the caller and shared by all evaluations at the vertex.
HM_HD V3 hair_eval_geom(const HairLobes& L, const hairdetail::FibreGeom& g, V3 wi_local, float* pdf) {
    return hairdetail::eval_with_geom(L, g, wi_local, pdf);
}
static HM_HD_OUTLINE V3 hair_sample_dir_geom(const HairLobes& L, const hairdetail::FibreGeom& g, float u0, float u1, float u2, float u3);

HM_HD V3 hair_sample_dir(const HairLobes& L, V3 wo_local, float h, float u0, float u1, float u2, float u3) {
    hairdetail::FibreGeom g;
    hairdetail::fibre_geom(L, wo_local, h, g);
    return hair_sample_dir_geom(L, g, u0, u1, u2, u3);
}

static HM_HD_OUTLINE V3 hair_sample_dir_geom(const HairLobes& L, const hairdetail::FibreGeom& g, float u0, float u1, float u2, float u3) {
    using namespace hairdetail;

    int p = 0;
    float eps1 = u0;
    for (p = 0; p < 3; ++p) {
        if (eps1 < g.ap_pdf[p]) break;
        eps1 -= g.ap_pdf[p];
    }
    float sin_op, cos_op;
    tilt(L, p, g.sin_o, g.cos_o, sin_op, cos_op);

    float vs = L.v_sample[p];
    float eps2 = fmaxf(u1, 1e-5f);
    float cos_theta = 1.f + vs * logf(eps2 + (1.f - eps2) * L.smp_exp[p]);
    float sin_theta = safe_sqrt(1 - sqr(cos_theta));
    float cos_phi = cosf(kTwoPi * u2);
    float sin_i = -cos_theta * sin_op + sin_theta * cos_phi * cos_op;
    float cos_i = safe_sqrt(1 - sqr(sin_i));

    float dphi;
    if (p < 3)
        dphi = net_phi(p, g.gamma_o, g.gamma_t) + sample_trimmed_logistic(L, u3);
    else
        dphi = kTwoPi * u3;

    float phi_i = g.phi_o + dphi;
    return normalize(V3(sin_i, cos_i * cosf(phi_i), cos_i * sinf(phi_i)));
}

// ---------------------------------------------------------------------------
// Head surface: 0.5 * Lambert + 0.5 * GGX (Heitz VNDF sampling), isotropic alpha.
// ---------------------------------------------------------------------------
namespace surfdetail {

HM_HD float ggx_d(float a, V3 n) {
    float t1 = n.x / a, t2 = n.y / a, t3 = n.z;
    float value = kPi * a * a * powf(t1 * t1 + t2 * t2 + t3 * t3, 2.0f);
    return 1.0f / value;
}
HM_HD float ggx_lambda(float a, V3 v) {
    float t1 = v.x * a, t2 = v.y * a, t3 = v.z;
    float t4 = sqrtf(1.0f + (t1 * t1 + t2 * t2) / (t3 * t3));
    return 0.5f * (-1.0f + t4);
}
HM_HD float ggx_g1(float a, V3 v) {
    if (v.z <= 0.0f) return 0.0f;
    return 1.0f / (1.0f + ggx_lambda(a, v));
}
HM_HD float ggx_g2(float a, V3 v, V3 l) {
    if (v.z <= 0.0f || l.z <= 0.0f) return 0.0f;
    return 1.0f / (1.0f + ggx_lambda(a, v) + ggx_lambda(a, l));
}
HM_HD float ggx_spec(float a, V3 v, V3 l) {
    V3 hv = normalize(v + l);
    return ggx_d(a, hv) * ggx_g2(a, v, l) / 4.0f / v.z / l.z;
}
HM_HD V3 sample_vndf(float a, V3 v, float u1, float u2) {
    V3 vh = normalize(v * V3(a, a, 1.0f));
    float lensq = vh.x * vh.x + vh.y * vh.y;
    V3 t1 = lensq > 0.0f ? V3(-vh.y, vh.x, 0.f) / sqrtf(lensq) : V3(1.f, 0.f, 0.f);
    V3 t2 = cross(vh, t1);
    float r = sqrtf(u1);
    float phi = 2.0f * kPi * u2;
    float a1 = r * cosf(phi);
    float a2 = r * sinf(phi);
    float s = 0.5f * (1.0f + vh.z);
    a2 = (1.0f - s) * sqrtf(1.0f - a1 * a1) + s * a2;
    V3 nh = a1 * t1 + a2 * t2 + sqrtf(fmaxf(0.0f, 1.0f - a1 * a1 - a2 * a2)) * vh;
    return normalize(V3(a * nh.x, a * nh.y, fmaxf(0.0f, nh.z)));
}

}  // namespace surfdetail

// f * cos
static HM_HD_OUTLINE V3 surf_eval(V3 wo, V3 wi, V3 kd, float alpha) {
    V3 brdf(0.f);
    if (wo.z > 0.f && wi.z > 0.f) {
        brdf += 0.5f * kd / kPi;
        brdf += V3(0.5f * surfdetail::ggx_spec(alpha, wo, wi));
    }
    return brdf * fabsf(wi.z);
}

static HM_HD_OUTLINE float surf_pdf(float alpha, V3 v, V3 ne) {
    float g1 = surfdetail::ggx_g1(alpha, v);
    float m = fmaxf(0.f, dot(v, ne));
    float d = surfdetail::ggx_d(alpha, ne);
    float dv = g1 * m * d / v.z;
    return dv / (4.f * dot(v, ne));
}

static HM_HD_OUTLINE V3 surf_sample(float u1, float u2, float alpha, V3 v, float* pdf) {
    V3 n = surfdetail::sample_vndf(alpha, v, u1, u2);
    V3 l = -v + 2.0f * n * dot(v, n);
    *pdf = surf_pdf(alpha, v, n);
    return normalize(l);
}

}  // namespace hm
