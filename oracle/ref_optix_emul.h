// ref_optix_emul.h — TEST INFRASTRUCTURE (oracle), not product code.
//
// Host-side stand-ins for the handful of OptiX / OWL device intrinsics the
// reference's per-path headers name, so that the reference's OWN source files
// (/root/reference/cuda_headers/*.cuh, /root/reference/cuda/*.cu) compile with g++
// and run on the CPU unmodified (SURVEY §8c recipe).  Nothing from the reference is
// copied: its files are #included from where they lie.
//
// What is emulated and how:
//   * __device__/__host__/__forceinline__/__constant__   -> nothing / inline
//   * optixGet*() hit attributes, owl::getPRD/getProgramData/getLaunchIndex
//         -> a thread_local "current hit" record filled by our traceRay()
//   * owl::traceRay(handle, ray, prd)
//         -> the build's own BVH + intersector compiled for the host
//            (hairmsnn_b200/csrc/hm_bvh.h), followed by a call into the reference's
//            closest-hit / any-hit / miss programs exactly as the OptiX pipeline
//            would dispatch them (SBT: ray type 0 radiance, 1 shadow, 2 multiscatter)
//   * tex2D<float4|float>(obj, x, y)
//         -> CUDA texture addressing restated from the CUDA C Programming Guide,
//            "Texture Fetching": normalized coords, clamp addressing, nearest-point
//            (floor) for the CDF/PDF tables, bilinear with 8-bit fixed-point
//            fractional weights for the RGBA32F environment map.
//   * RNG argument order: g++ evaluates constructor arguments right-to-left, nvcc
//     device code left-to-right (verified).  The reference draws randoms
//     inside constructor argument lists; to reproduce the GPU order on the host the
//     including .cpp wraps lcg_randomf in a small reordering queue (see
//     REF_DRAW_FIX below).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include <vector_types.h>
#include <vector_functions.h>
#include <cuda_runtime.h>

#undef __device__
#undef __host__
#undef __forceinline__
#undef __inline__
#undef __constant__
#define __device__
#define __host__
#define __forceinline__ inline
#define __inline__ inline
#define __constant__
#ifndef __CUDACC__
#define __CUDACC__ 1
#endif

typedef unsigned long long OptixTraversableHandle;
typedef unsigned int OptixVisibilityMask;

#include "owl/common/math/vec.h"
using namespace owl;
using std::abs;
using std::isinf;
using std::isnan;
using std::max;
using std::min;

// product geometry code, host build (the reference has no intersector source)
#include "../hairmsnn_b200/csrc/hm_bvh.h"

namespace refemu {

struct HostTexture {
    const float* data;  // channels floats per texel
    int w, h, channels;
    bool linear;
};

struct HitContext {
    // current ray
    float3 org, dir;
    float tmax;
    // current hit
    unsigned prim;       // index within its geometry (segment id or triangle id)
    float u;             // curve parameter
    float2 bary;
    void* prd;
    const void* program_data;
    int launch_x, launch_y;
    bool terminated;
};

extern thread_local HitContext g_ctx;
extern std::vector<HostTexture> g_textures;   // cudaTextureObject_t = index + 1
extern hm::GeomView g_geom;
extern const void* g_triangle_program_data;   // TriangleMeshData*
extern const void* g_raygen_program_data;     // RayGenData*
extern const void* g_miss_program_data;       // MissProgData*

inline float tex_fetch(const HostTexture& t, int x, int y, int c) {
    x = std::min(std::max(x, 0), t.w - 1);
    y = std::min(std::max(y, 0), t.h - 1);
    return t.data[((size_t)y * t.w + x) * t.channels + c];
}

inline void tex_sample(const HostTexture& t, float xn, float yn, float* out) {
    // clamp addressing on normalized coordinates
    float x = xn * (float)t.w, y = yn * (float)t.h;
    if (!t.linear) {
        int i = (int)floorf(x), j = (int)floorf(y);
        for (int c = 0; c < t.channels; ++c) out[c] = tex_fetch(t, i, j, c);
        return;
    }
    float xb = x - 0.5f, yb = y - 0.5f;
    float fi = floorf(xb), fj = floorf(yb);
    int i = (int)fi, j = (int)fj;
    // 9-bit fixed point with 8 fractional bits
    float a = floorf((xb - fi) * 256.f + 0.5f) / 256.f;
    float b = floorf((yb - fj) * 256.f + 0.5f) / 256.f;
    for (int c = 0; c < t.channels; ++c) {
        float t00 = tex_fetch(t, i, j, c), t10 = tex_fetch(t, i + 1, j, c);
        float t01 = tex_fetch(t, i, j + 1, c), t11 = tex_fetch(t, i + 1, j + 1, c);
        out[c] = (1 - a) * (1 - b) * t00 + a * (1 - b) * t10 + (1 - a) * b * t01 + a * b * t11;
    }
}

}  // namespace refemu

extern thread_local int g_ctx_cp0;   // first control point of the segment just hit
extern const int* g_seg_cp;          // segment id -> first control point

template <typename T>
inline T tex2D(cudaTextureObject_t obj, float x, float y);

template <>
inline float4 tex2D<float4>(cudaTextureObject_t obj, float x, float y) {
    float v[4] = {0, 0, 0, 0};
    refemu::tex_sample(refemu::g_textures[(size_t)obj - 1], x, y, v);
    return make_float4(v[0], v[1], v[2], v[3]);
}
template <>
inline float tex2D<float>(cudaTextureObject_t obj, float x, float y) {
    float v[4] = {0, 0, 0, 0};
    refemu::tex_sample(refemu::g_textures[(size_t)obj - 1], x, y, v);
    return v[0];
}

// ---- OptiX device intrinsics named by the reference ------------------------------
inline float optixGetRayTmax() { return refemu::g_ctx.tmax; }
inline float3 optixGetWorldRayOrigin() { return refemu::g_ctx.org; }
inline float3 optixGetWorldRayDirection() { return refemu::g_ctx.dir; }
inline OptixTraversableHandle optixGetGASTraversableHandle() { return 0; }
inline unsigned optixGetSbtGASIndex() { return 0; }
inline unsigned optixGetPrimitiveIndex() { return refemu::g_ctx.prim; }
inline float optixGetCurveParameter() { return refemu::g_ctx.u; }
inline float2 optixGetTriangleBarycentrics() { return refemu::g_ctx.bary; }
inline void optixTerminateRay() { refemu::g_ctx.terminated = true; }
inline void optixIgnoreIntersection() {}
inline void optixGetCatmullRomVertexData(OptixTraversableHandle, unsigned prim, unsigned, float, float4 data[4]) {
    const hm::GeomView& g = refemu::g_geom;
    // prim is a segment id; the host keeps seg -> first control point in leaf order,
    // so the caller of traceRay stores the CP index in g_ctx (see below)
    for (int k = 0; k < 4; ++k) {
        hm::F4 c = g.cps[g_ctx_cp0 + k];
        data[k] = make_float4(c.x, c.y, c.z, c.w);
    }
    (void)prim;
}
inline void optixGetLinearCurveVertexData(OptixTraversableHandle, unsigned, unsigned, float, float4*) {}

#define OPTIX_RAYGEN_PROGRAM(name) extern "C" void ref_raygen_##name
#define OPTIX_CLOSEST_HIT_PROGRAM(name) extern "C" void ref_closesthit_##name
#define OPTIX_ANY_HIT_PROGRAM(name) extern "C" void ref_anyhit_##name
#define OPTIX_MISS_PROGRAM(name) extern "C" void ref_miss_##name

extern "C" void ref_closesthit_hairCH();
extern "C" void ref_closesthit_triangleMeshCH();
extern "C" void ref_anyhit_hairAHShadow();
extern "C" void ref_anyhit_triangleMeshAHShadow();
extern "C" void ref_anyhit_hairAHMultiScatter();
extern "C" void ref_anyhit_triangleMeshAHMultiScatter();
extern "C" void ref_miss_miss();

namespace owl {

inline float linear_to_srgb(float x) {
    if (x <= 0.0031308f) return 12.92f * x;
    return 1.055f * powf(x, 1.f / 2.4f) - 0.055f;
}
inline uint32_t make_8bit(const float f) { return std::min(255, std::max(0, int(f * 256.f))); }
inline uint32_t make_rgba(const vec3f color) {
    return (make_8bit(color.x) << 0) + (make_8bit(color.y) << 8) + (make_8bit(color.z) << 16) + (0xffU << 24);
}

inline vec2i getLaunchIndex() { return vec2i(refemu::g_ctx.launch_x, refemu::g_ctx.launch_y); }

template <typename T>
inline const T& getProgramData() { return *(const T*)refemu::g_ctx.program_data; }
template <typename T>
inline T& getPRD() { return *(T*)refemu::g_ctx.prd; }

template <int _rayType = 0, int _numRayTypes = 1>
struct RayT {
    enum { rayType = _rayType };
    enum { numRayTypes = _numRayTypes };
    vec3f origin, direction;
    float tmin = 0.f, tmax = 1e30f, time = 0.f;
    OptixVisibilityMask visibilityMask = (OptixVisibilityMask)-1;
};

template <typename RayType, typename PRD>
inline void traceRay(OptixTraversableHandle, const RayType& ray, PRD& prd, uint32_t = 0u) {
    using namespace refemu;
    HitContext saved = g_ctx;
    hm::V3 o(ray.origin.x, ray.origin.y, ray.origin.z), d(ray.direction.x, ray.direction.y, ray.direction.z);
    g_ctx.org = make_float3(o.x, o.y, o.z);
    g_ctx.dir = make_float3(d.x, d.y, d.z);
    g_ctx.prd = (void*)&prd;
    g_ctx.terminated = false;
    const int ns = g_geom.num_segments;
    if ((int)RayType::rayType == 0) {
        hm::Hit h = hm::trace<false>(g_geom, o, d, ray.tmin, ray.tmax);
        if (h.prim < 0) {
            g_ctx.program_data = g_miss_program_data;
            ref_miss_miss();
        } else if (h.prim < ns) {
            g_ctx.tmax = h.t; g_ctx.prim = (unsigned)h.prim; g_ctx.u = h.u;
            g_ctx_cp0 = g_seg_cp[h.prim];
            g_ctx.program_data = nullptr;
            ref_closesthit_hairCH();
        } else {
            g_ctx.tmax = h.t; g_ctx.prim = (unsigned)(h.prim - ns);
            g_ctx.bary = make_float2(h.u, h.v);
            g_ctx.program_data = g_triangle_program_data;
            ref_closesthit_triangleMeshCH();
        }
    } else if ((int)RayType::rayType == 1) {
        hm::Hit h = hm::trace<true>(g_geom, o, d, ray.tmin, ray.tmax);
        if (h.prim >= 0) {
            g_ctx.tmax = h.t;
            if (h.prim < ns) ref_anyhit_hairAHShadow();
            else { g_ctx.program_data = g_triangle_program_data; ref_anyhit_triangleMeshAHShadow(); }
        }
        // shadow rays have no miss program bound (OWL installs `miss` for ray type 0 only)
    } else {
        // multiscatter rays are never traced by the shipped renderers
    }
    g_ctx = saved;
}

}  // namespace owl
