// hm_rng.h — per-(pixel, frame) sample streams.
//
// Bit-exact with the reference stream definition (cuda_headers/lcg_random.cuh:11-65):
// a murmur3-mixed seed of (pixel index in the FULL frame, frame id) followed by a
// 32-bit LCG.  The float draw is ldexp((float)u32, -32): the u32 -> float cast
// rounds to nearest, so the draw can be exactly 1.0f — kept as is.
//
// Draw ORDER is part of the contract (SURVEY §7 "RNG draw order"): every caller in
// this code base pulls draws into named temporaries, first call = first component.
#pragma once
#include "hm_math.h"

namespace hm {

struct Rng {
    uint32_t state;
};

HM_HD uint32_t rotl32(uint32_t v, int r) { return (v << r) | (v >> (32 - r)); }

HM_HD uint32_t mm3_mix(uint32_t h, uint32_t k) {
    k *= 0xcc9e2d51u;
    k = rotl32(k, 15);
    k *= 0x1b873593u;
    h ^= k;
    h = rotl32(h, 13) * 5u + 0xe6546b64u;
    return h;
}

HM_HD uint32_t mm3_final(uint32_t h) {
    h ^= h >> 16;
    h *= 0x85ebca6bu;
    h ^= h >> 13;
    h *= 0xc2b2ae35u;
    h ^= h >> 16;
    return h;
}

// frame_id is accumId + 10007 at every call site of the reference
// (cuda/path_tracing.cu:27, cuda/hair_msnn.cu:194, cuda/nrc.cu:321).
HM_HD Rng rng_seed(int frame_id, uint32_t px, uint32_t py, uint32_t width) {
    Rng r;
    r.state = mm3_mix(0u, px + py * width);
    r.state = mm3_mix(r.state, (uint32_t)frame_id);
    r.state = mm3_final(r.state);
    return r;
}

HM_HD uint32_t rng_next_u32(Rng& r) {
    r.state = r.state * 1664525u + 1013904223u;
    return r.state;
}

HM_HD float rng_next(Rng& r) {
    // (float)u32 * 2^-32 is exact scaling == ldexp((float)u32, -32)
    return (float)rng_next_u32(r) * 2.3283064365386963e-10f;
}

}  // namespace hm
