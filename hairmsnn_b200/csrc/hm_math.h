// hm_math.h — small fp32 vector/scalar toolkit shared by host and device code.
//
// Everything on the per-path hot loop is fp32; the helpers here
// pin the few constants whose exact bit pattern matters for parity with the
// reference (cuda_headers/utils.cuh:10-11 defines Pi = 3.1415926f and an
// unparenthesised TWO_Pi = 2.f * 3.14159f — both reproduced as named floats).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define HM_HD __host__ __device__ __forceinline__
#define HM_D __device__ __forceinline__
// one shared copy per kernel instead of one per call site (instruction-cache footprint)
#define HM_HD_OUTLINE __host__ __device__ __noinline__
#else
#define HM_HD inline
#define HM_D inline
#define HM_HD_OUTLINE inline
#endif

namespace hm {

// reference constants (utils.cuh:10-11)
static constexpr float kPi = 3.1415926f;
static constexpr float kTwoPiLoose = 2.f * 3.14159f;   // "TWO_Pi" as the reference expands it
static constexpr float kTwoPi = 2.f * 3.1415926f;      // "2 * Pi" (int * float)

struct V3 {
    float x, y, z;
    HM_HD V3() : x(0.f), y(0.f), z(0.f) {}
    HM_HD explicit V3(float a) : x(a), y(a), z(a) {}
    HM_HD V3(float a, float b, float c) : x(a), y(b), z(c) {}
    HM_HD float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};

HM_HD V3 operator+(V3 a, V3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
HM_HD V3 operator-(V3 a, V3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
HM_HD V3 operator-(V3 a) { return V3(-a.x, -a.y, -a.z); }
HM_HD V3 operator*(V3 a, V3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
HM_HD V3 operator/(V3 a, V3 b) { return V3(a.x / b.x, a.y / b.y, a.z / b.z); }
HM_HD V3 operator*(float s, V3 a) { return V3(s * a.x, s * a.y, s * a.z); }
HM_HD V3 operator*(V3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
HM_HD V3 operator/(V3 a, float s) { return V3(a.x / s, a.y / s, a.z / s); }
HM_HD V3& operator+=(V3& a, V3 b) { a = a + b; return a; }
HM_HD V3& operator-=(V3& a, V3 b) { a = a - b; return a; }
HM_HD V3& operator*=(V3& a, float s) { a = a * s; return a; }
HM_HD bool operator==(V3 a, V3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }

HM_HD float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
HM_HD V3 cross(V3 a, V3 b) {
    return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
HM_HD float length(V3 a) { return sqrtf(dot(a, a)); }
// owl's normalize is v * (1/sqrt(dot)) (owl/common/math/vec/functors.h); keep the
// reciprocal-multiply form so directions round the same way as the reference.
HM_HD V3 normalize(V3 a) { return a * (1.f / sqrtf(dot(a, a))); }
HM_HD V3 vmin(V3 a, V3 b) { return V3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
HM_HD V3 vmax(V3 a, V3 b) { return V3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
HM_HD bool any_nan(V3 a) { return isnan(a.x) || isnan(a.y) || isnan(a.z); }
HM_HD bool any_inf(V3 a) { return isinf(a.x) || isinf(a.y) || isinf(a.z); }

struct V4 {
    float x, y, z, w;
    HM_HD V4() : x(0.f), y(0.f), z(0.f), w(0.f) {}
    HM_HD V4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    HM_HD V4(V3 v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
    HM_HD V3 xyz() const { return V3(x, y, z); }
};
HM_HD V4 operator+(V4 a, V4 b) { return V4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
HM_HD V4 operator-(V4 a, V4 b) { return V4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
HM_HD V4 operator*(float s, V4 a) { return V4(s * a.x, s * a.y, s * a.z, s * a.w); }
HM_HD V4 operator*(V4 a, float s) { return V4(a.x * s, a.y * s, a.z * s, a.w * s); }
HM_HD V4 operator/(V4 a, float s) { return V4(a.x / s, a.y / s, a.z / s, a.w / s); }

HM_HD float sqr(float v) { return v * v; }
// IEEE quotient whatever the translation unit's -prec-div setting (hm_shade_kernels.cu is built with approximate division)
HM_HD float div_exact(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdiv_rn(a, b);
#else
    return a / b;
#endif
}
HM_HD float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }
HM_HD int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
HM_HD float safe_sqrt(float v) { return sqrtf(fmaxf(0.f, v)); }
HM_HD float safe_asin(float v) { return asinf(clampf(v, -1.f, 1.f)); }

// Row-vector 3x3: rows m0,m1,m2; apply = (dot(m0,v), dot(m1,v), dot(m2,v)).
struct M3 {
    V3 r0, r1, r2;
    HM_HD V3 apply(V3 v) const { return V3(dot(r0, v), dot(r1, v), dot(r2, v)); }
    HM_HD M3 transposed() const {
        M3 t;
        t.r0 = V3(r0.x, r1.x, r2.x);
        t.r1 = V3(r0.y, r1.y, r2.y);
        t.r2 = V3(r0.z, r1.z, r2.z);
        return t;
    }
};

// Rec.709 luminance used for Russian roulette (utils.cuh:84-88).
HM_HD float luminance709(V3 c) { return 0.212671f * c.x + 0.715160f * c.y + 0.072169f * c.z; }

// MIS power heuristic with nf = ng = 1 (utils.cuh:289-293).
HM_HD float power_heuristic(float f, float g) { return (f * f) / (f * f + g * g); }

// Display transform of the 8-bit framebuffer (owl_device.h:72-91).
HM_HD float linear_to_srgb(float x) {
    if (x <= 0.0031308f) return 12.92f * x;
    return 1.055f * powf(x, 1.f / 2.4f) - 0.055f;
}
HM_HD uint32_t to_8bit(float f) {
    int v = (int)(f * 256.f);
    return (uint32_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}
HM_HD uint32_t pack_rgba8(V3 c) {
    return (to_8bit(c.x) << 0) + (to_8bit(c.y) << 8) + (to_8bit(c.z) << 16) + (0xffU << 24);
}

}  // namespace hm
