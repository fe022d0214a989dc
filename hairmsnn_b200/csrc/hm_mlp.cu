// hm_mlp.cu — radiance-cache network kernels (sm_100a).
//
// One network configuration: Composite[HashGrid 16x2 | OneBlob 6x4 |
// Identity] -> 64 -> 64 -> 64 -> 16(3), ReLU, no biases; RelativeL2Luminance; Adam under
// exponential decay.  tiny-cuda-nn spreads this over ~10 kernels (encodings/grid.h:221,
// encodings/oneblob.h:99, encodings/identity.h:46, src/fully_fused_mlp.cu:499/150,
// losses/relative_l2_luminance.h:40, 3 CUTLASS split-K GEMMs, grid.h:395,
// optimizers/adam.h:48); here it is
//   k_mlp_forward<TRAIN=false> : encode + 3 layers + fp32 AoS output, one pass over HBM
//   k_mlp_forward<TRAIN=true>  : same + saves e/h1/h2, loss value and dL/dy
//   k_mlp_backward             : dgrad through the 3 layers + hash-grid scatter
//   k_wgrad / k_wgrad_out      : weight gradients
//   k_adam                     : optimizer step (+ fp16 copy, + gradient clear)
//
// Tiling: a CTA owns 128 rows (4 warps x 32 rows).  Activations live in shared memory as
// fp16 [128][72] (72-half stride = conflict-free fragment loads); weights are resident
// in shared memory for the CTA's lifetime; every warp works on its own 32 rows, so after
// the weight load the only synchronisation is __syncwarp.  Accumulation is fp32 (the
// reference accumulates in fp16 inside wmma — tolerance documented in tests/test_gpu_mlp.py).
#include "hm_mlp.h"

#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <random>
#include <stdexcept>

#include "hm_io.h"

namespace hm {

#define HM_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            throw std::runtime_error(std::string("CUDA: ") + cudaGetErrorString(e_) + " at " + \
                                     __FILE__ + ":" + std::to_string(__LINE__));               \
    } while (0)

namespace {

constexpr int kW = 64;          // network width
constexpr int kOutPad = 16;     // padded output width
constexpr int kStride = 72;     // smem row stride in halves
constexpr int kTile = 128;      // rows per CTA
constexpr float kLossScale = 128.f;

struct NetShape {
    GridLayout grid;
    int in_ch;
    int blob_dims, blob_bins, identity_dims;
    int feats;
    uint32_t hashed_mask;       // bit l set: level l is hashed
    uint32_t pow2_mask;         // bit l set: level l's size is a power of two (index % size == index & (size - 1))
};

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint32_t lds32(const __half* p) { return *reinterpret_cast<const uint32_t*>(p); }
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

// acc[mt][nt][4] = act(32 x 64, this warp's rows) * W^T where W is [NT*8][64] (row stride kStride)
template <int NT>
__device__ __forceinline__ void warp_gemm(const __half* act, const __half* W, float (&acc)[2][NT][4], int k_steps = 4) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        if (kk >= k_steps) break;
        const int k0 = kk * 16 + t * 2;
        uint32_t a[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            const __half* r0 = act + (mt * 16 + g) * kStride + k0;
            a[mt][0] = lds32(r0);
            a[mt][1] = lds32(r0 + 8 * kStride);
            a[mt][2] = lds32(r0 + 8);
            a[mt][3] = lds32(r0 + 8 * kStride + 8);
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const __half* w = W + (nt * 8 + g) * kStride + k0;
            uint32_t b0 = lds32(w), b1 = lds32(w + 8);
            mma16816(acc[0][nt], a[0], b0, b1);
            mma16816(acc[1][nt], a[1], b0, b1);
        }
    }
}

// writes fragment values (optionally ReLU'd) back into the warp's activation rows
template <int NT, bool RELU>
__device__ __forceinline__ void warp_store_act(__half* act, const float (&acc)[2][NT][4]) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            float c0 = acc[mt][nt][0], c1 = acc[mt][nt][1], c2 = acc[mt][nt][2], c3 = acc[mt][nt][3];
            if (RELU) { c0 = fmaxf(c0, 0.f); c1 = fmaxf(c1, 0.f); c2 = fmaxf(c2, 0.f); c3 = fmaxf(c3, 0.f); }
            __half* p = act + (mt * 16 + g) * kStride + nt * 8 + t * 2;
            *reinterpret_cast<uint32_t*>(p) = pack2(c0, c1);
            *reinterpret_cast<uint32_t*>(p + 8 * kStride) = pack2(c2, c3);
        }
}

// copies a [rows][64] fp16 matrix from global into shared memory with row stride kStride
__device__ __forceinline__ void load_matrix(__half* dst, const __half* src, int rows) {
    for (int i = threadIdx.x; i < rows * 8; i += blockDim.x) {
        int r = i >> 3, c = i & 7;
        *reinterpret_cast<uint4*>(dst + r * kStride + c * 8) = __ldg(reinterpret_cast<const uint4*>(src + r * kW + c * 8));
    }
}
// transposed copy: dst[n][k] = src[k][n], src is [rows_k][64]; dst row stride kStride
__device__ __forceinline__ void load_matrix_t(__half* dst, const __half* src, int rows_k) {
    for (int i = threadIdx.x; i < rows_k * kW; i += blockDim.x) {
        int k = i / kW, n = i % kW;
        dst[n * kStride + k] = src[i];
    }
}

// ---- encoding --------------------------------------------------------------------
__device__ __forceinline__ uint32_t grid_cell_index(uint32_t x, uint32_t y, uint32_t z, uint32_t res, uint32_t size, bool hashed) {
    uint32_t idx = hashed ? (x ^ (y * 2654435761u) ^ (z * 805459861u)) : (x + y * res + z * res * res);
    return idx % size;
}

__device__ __forceinline__ float quartic_cdf(float x, float inv_radius) {
    const float u = x * inv_radius;
    const float u2 = u * u;
    const float u4 = u2 * u2;
    return fmaxf(0.0f, fminf(1.0f, (15.f / 16.f) * u * (1 - (2.f / 3.f) * u2 + (1.f / 5.f) * u4) + 0.5f));
}

// One level of the hash grid for one row (kernel_grid, encodings/grid.h:221-351): the 8 corner values are
// fetched first (8 independent loads), then accumulated exactly as tiny-cuda-nn does — fp32 product, rounded to
// fp16, added in fp16 (grid.h:337-341) — so the encoded features are bit-identical to the reference's.
template <bool POW2>
__device__ __forceinline__ uint32_t encode_level(const NetShape& S, const __half2* __restrict__ table, float x0, float x1, float x2, int l) {
    const float scale = S.grid.scale[l];
    const uint32_t res = S.grid.resolution[l];
    const uint32_t size = S.grid.offset[l + 1] - S.grid.offset[l];
    const bool hashed = (S.hashed_mask >> l) & 1u;
    const __half2* tl = table + S.grid.offset[l];
    float p0 = x0 * scale + 0.5f, p1 = x1 * scale + 0.5f, p2 = x2 * scale + 0.5f;      // pos_fract, common_device.h:434-445
    const float f0 = floorf(p0), f1 = floorf(p1), f2 = floorf(p2);
    const uint32_t q0 = (uint32_t)(int)f0, q1 = (uint32_t)(int)f1, q2 = (uint32_t)(int)f2;
    p0 -= f0; p1 -= f1; p2 -= f2;
    // per-axis index terms: dense = x + y*res + z*res^2, hashed = x ^ y*2654435761 ^ z*805459861 (grid.h:171-187)
    uint32_t ix[2], iy[2], iz[2];
    ix[0] = q0; ix[1] = q0 + 1u;
    if (hashed) { iy[0] = q1 * 2654435761u; iy[1] = iy[0] + 2654435761u; iz[0] = q2 * 805459861u; iz[1] = iz[0] + 805459861u; }
    else { iy[0] = q1 * res; iy[1] = iy[0] + res; iz[0] = q2 * res * res; iz[1] = iz[0] + res * res; }
    __half2 v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        uint32_t idx = hashed ? (ix[c & 1] ^ iy[(c >> 1) & 1] ^ iz[(c >> 2) & 1]) : (ix[c & 1] + iy[(c >> 1) & 1] + iz[(c >> 2) & 1]);
        idx = (POW2 || ((S.pow2_mask >> l) & 1u)) ? (idx & (size - 1u)) : (idx % size);
        v[c] = __ldg(tl + idx);
    }
    const float wx[2] = {1 - p0, p0}, wy[2] = {1 - p1, p1}, wz[2] = {1 - p2, p2};
    __half2 acc = __floats2half2_rn(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float w = (wx[c & 1] * wy[(c >> 1) & 1]) * wz[(c >> 2) & 1];
        const float2 f = __half22float2(v[c]);
        acc = __hadd2(acc, __floats2half2_rn(w * f.x, w * f.y));
    }
    return *reinterpret_cast<uint32_t*>(&acc);
}

// Encodes (part `part` of `parts` of) one input row into 64 halves; `at(col)` gives the address of column `col`
// (even columns are 4-byte aligned and followed by their odd neighbour).  With parts = 2 two threads share a row:
// part p takes the grid levels and OneBlob dimensions of parity p, part parts-1 the identity columns and the padding.
template <bool POW2, class At>
__device__ __forceinline__ void encode_row(const NetShape& S, const __half2* __restrict__ table, const float* x, At at, int part = 0, int parts = 1) {
    const int nl = S.grid.n_levels;
#pragma unroll 2
    for (int l = part; l < nl; l += parts) *reinterpret_cast<uint32_t*>(at(2 * l)) = encode_level<POW2>(S, table, x[0], x[1], x[2], l);
    int c0 = nl * 2;
    const int nb = S.blob_bins;
    const float inv_r = (float)nb;
    const bool last = part == parts - 1;
    const bool nrc_pad = S.identity_dims == 0;
    for (int j = part; j < S.blob_dims; j += parts) {
        const float xv = x[3 + j];
        float left = quartic_cdf(-xv, inv_r) + quartic_cdf(-xv - 1.0f, inv_r) + quartic_cdf(-xv + 1.0f, inv_r);
        for (int k = 0; k < nb; ++k) {
            const float rb = (float)(k + 1) / (float)nb;
            const float right = quartic_cdf(rb - xv, inv_r) + quartic_cdf(rb - xv - 1.0f, inv_r) + quartic_cdf(rb - xv + 1.0f, inv_r);
            const int col = c0 + j * nb + k;
            // render_nrc: columns blob_dims .. blob_dims + n_pad - 1 of the OneBlob slice are overwritten with ones (below)
            if (!(nrc_pad && col >= c0 + S.blob_dims && col < c0 + S.blob_dims + (kW - c0 - S.blob_dims * nb))) *at(col) = __float2half_rn(right - left);
            left = right;
        }
    }
    if (!last) return;
    const int blob0 = c0;
    c0 += S.blob_dims * nb;
    for (int j = 0; j < S.identity_dims; ++j) *at(c0 + j) = __float2half_rn(x[3 + S.blob_dims + j]);
    if (!nrc_pad) {
        // Identity is the last nested encoding and pads the network input with ones (identity.h:46-66)
        for (int c = c0 + S.identity_dims; c < kW; ++c) *at(c) = __float2half_rn(1.0f);
    } else {
        // render_nrc (9 inputs): the Identity encoding is dropped (composite.h:181) and OneBlob pads.  Its SoA
        // padding writes the ones at element offset N * n_dims_to_encode of its own slice (oneblob.h:221-225),
        // i.e. over its rows blob_dims .. blob_dims + n_pad - 1 (network inputs 38..45), and never writes the real
        // padding rows, which keep the allocation's contents — zero in a fresh one.  Pinned by the reference's
        // own tiny-cuda-nn: tests/golden/tcnn_9.npz.
        const int n_pad = kW - c0;
        for (int c = 0; c < n_pad; ++c) *at(blob0 + S.blob_dims + c) = __float2half_rn(1.0f);
        for (int c = c0; c < kW; ++c) *at(c) = __float2half_rn(0.0f);
    }
}

struct FwdArgs {
    const float* in;        // [n][in_ch]
    float* out;             // [n][3]           (inference)
    const float* target;    // [n][3]           (training)
    __half* e; __half* h1; __half* h2;   // [n][64] (training)
    __half* dy;             // [n][4]           (training) loss gradient, scaled
    float* loss;            // [1]
    float inv_n_total;      // 1 / (records * 3)
    int n_tiles;
    const int* tile_mask;   // optional [n_tiles]: tiles whose flag is 0 are skipped (their outputs are not written)
};

template <bool TRAIN>
__global__ void __launch_bounds__(kTile) k_mlp_forward(const NetShape S, const __half* __restrict__ params, FwdArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __half* sW0 = reinterpret_cast<__half*>(smem_raw);
    __half* sW1 = sW0 + kW * kStride;
    __half* sWo = sW1 + kW * kStride;
    __half* sAct = sWo + kOutPad * kStride;
    load_matrix(sW0, params, kW);
    load_matrix(sW1, params + kW * kW, kW);
    load_matrix(sWo, params + 2 * kW * kW, kOutPad);
    __syncthreads();
    const __half2* table = reinterpret_cast<const __half2*>(params + 2 * kW * kW + kOutPad * kW);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    __half* wAct = sAct + warp * 32 * kStride;

    for (int tile = blockIdx.x; tile < A.n_tiles; tile += gridDim.x) {
        if (A.tile_mask && __ldg(A.tile_mask + tile) == 0) continue;
        const size_t row = (size_t)tile * kTile + threadIdx.x;
        {
            float x[12];
            const float* src = A.in + row * S.in_ch;
            if (S.in_ch == 12) {
                const float4* s4 = reinterpret_cast<const float4*>(src);
                float4 a = __ldg(s4), b = __ldg(s4 + 1), c = __ldg(s4 + 2);
                x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
                x[8] = c.x; x[9] = c.y; x[10] = c.z; x[11] = c.w;
            } else {
                for (int i = 0; i < 12; ++i) x[i] = i < S.in_ch ? __ldg(src + i) : 0.f;
            }
            __half* rowp = sAct + threadIdx.x * kStride;
            encode_row<false>(S, table, x, [rowp](int col) { return rowp + col; });
        }
        __syncwarp();
        if (TRAIN) {
            const uint4* s = reinterpret_cast<const uint4*>(sAct + threadIdx.x * kStride);
            uint4* d = reinterpret_cast<uint4*>(A.e + row * kW);
#pragma unroll
            for (int i = 0; i < 8; ++i) d[i] = s[i];
        }
        float acc[2][8][4];
        warp_gemm<8>(wAct, sW0, acc);
        __syncwarp();
        warp_store_act<8, true>(wAct, acc);
        __syncwarp();
        if (TRAIN) {
            const uint4* s = reinterpret_cast<const uint4*>(sAct + threadIdx.x * kStride);
            uint4* d = reinterpret_cast<uint4*>(A.h1 + row * kW);
#pragma unroll
            for (int i = 0; i < 8; ++i) d[i] = s[i];
        }
        warp_gemm<8>(wAct, sW1, acc);
        __syncwarp();
        warp_store_act<8, true>(wAct, acc);
        __syncwarp();
        if (TRAIN) {
            const uint4* s = reinterpret_cast<const uint4*>(sAct + threadIdx.x * kStride);
            uint4* d = reinterpret_cast<uint4*>(A.h2 + row * kW);
#pragma unroll
            for (int i = 0; i < 8; ++i) d[i] = s[i];
        }
        float o[2][1][4];
        warp_gemm<1>(wAct, sWo, o);
        __syncwarp();
        // fragment: lanes t == 0 hold columns 0,1; t == 1 holds column 2 (and unused 3)
        const size_t wrow = (size_t)tile * kTile + warp * 32;
        float lsum = 0.f;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const size_t r = wrow + mt * 16 + hh * 8 + g;
                // network output is held in fp16 by the reference (trim_and_cast_from)
                float v0 = __half2float(__float2half_rn(o[mt][0][hh * 2 + 0]));
                float v1 = __half2float(__float2half_rn(o[mt][0][hh * 2 + 1]));
                if (!TRAIN) {
                    if (t == 0) { A.out[r * 3 + 0] = v0; A.out[r * 3 + 1] = v1; }
                    else if (t == 1) A.out[r * 3 + 2] = v0;
                } else {
                    // gather r,g,b of the row into every lane of the quad
                    const int base = lane & ~3;
                    float pr = __shfl_sync(0xffffffffu, v0, base), pg = __shfl_sync(0xffffffffu, v1, base);
                    float pb = __shfl_sync(0xffffffffu, v0, base + 1);
                    if (t < 2) {
                        const float lum = 0.299f * pr + 0.587f * pg + 0.114f * pb;
                        const float denom = lum * lum + 0.01f;
                        float d0 = 0.f, d1 = 0.f;
                        if (t == 0) {
                            float df0 = pr - __ldg(A.target + r * 3 + 0), df1 = pg - __ldg(A.target + r * 3 + 1);
                            lsum += df0 * df0 / denom * A.inv_n_total + df1 * df1 / denom * A.inv_n_total;
                            d0 = kLossScale * (2 * df0 / denom) * A.inv_n_total;
                            d1 = kLossScale * (2 * df1 / denom) * A.inv_n_total;
                        } else {
                            float df2 = pb - __ldg(A.target + r * 3 + 2);
                            lsum += df2 * df2 / denom * A.inv_n_total;
                            d0 = kLossScale * (2 * df2 / denom) * A.inv_n_total;
                        }
                        *reinterpret_cast<uint32_t*>(A.dy + r * 4 + t * 2) = pack2(d0, d1);
                    }
                }
            }
        if (TRAIN) {
#pragma unroll
            for (int ofs = 16; ofs > 0; ofs >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, ofs);
            if (lane == 0) atomicAdd(A.loss, lsum);
        }
        __syncwarp();
    }
}

// ---- tcgen05 / TMEM forward -----------------------------------------------------------
// Same computation as k_mlp_forward, with the three layers on the 5th-generation tensor
// cores: a CTA owns 128 rows; the activation tile A [128 x 64] fp16 and the weights
// W0, W1 [64 x 64], Wout [16 x 64] sit in shared memory in the canonical K-major
// SWIZZLE_128B layout (row = 128 B; 8-row groups of 1024 B; 16-byte chunk c of row r is
// stored at chunk c ^ (r & 7)); one elected thread issues 4 x tcgen05.mma (M=128, N=64|16,
// K=16) per layer into a 64-column TMEM accumulator and commits to an mbarrier; the
// epilogue is row-per-thread: tcgen05.ld 32x32b gives thread i the 64 fp32 accumulators of
// row i, which are ReLU'd, packed to fp16 and written back into A (swizzled) as the next
// layer's operand.  The encoded inputs are produced in-kernel, so A never touches HBM.
namespace tc {

constexpr uint32_t kABytes = 128 * 128, kWBytes = 64 * 128, kWoBytes = 16 * 128;
constexpr uint32_t kSmemBytes = kABytes + 2 * kWBytes + kWoBytes + 64 /*barrier + tmem ptr*/ + 1024 /*alignment slack*/;
constexpr uint32_t kTmemCols = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t sw128(uint32_t row, uint32_t chunk) {
    return (row >> 3) * 1024u + (row & 7u) * 128u + ((chunk ^ (row & 7u)) << 4);
}
// K-major, SWIZZLE_128B, 128-byte rows: LBO = 1 (16 B), SBO = 64 (1024 B), version 1
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffffu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)64 << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint32_t make_idesc(int n, uint32_t a_mn = 0, uint32_t b_mn = 0) {
    // c_format F32 (bit 4), A/B F16, a_major (bit 15) / b_major (bit 16): 0 = K-major, 1 = MN-major,
    // N >> 3 at bit 17, M >> 4 at bit 24
    return (1u << 4) | (a_mn << 15) | (b_mn << 16) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
// The SAME bytes read MN-major: a [rows][64] fp16 tile in the K-major SWIZZLE_128B layout above (128-byte rows,
// 8-row groups of 1024 B, 16-byte chunk c of row r at c ^ (r & 7)) is also the canonical MN-major SWIZZLE_128B
// layout of its transpose — MN = the 64 contiguous elements of a row, K = the rows: SBO = 1024 B between 8-row
// (K) groups, LBO = distance between 64-element blocks along MN (two tiles stacked for M or N = 128).  This is
// what lets the backward pass reuse the forward pass's weight tiles (dX = dY * W) and the activation tiles as
// both operands of the weight-gradient products (dW = dY^T * X) without any transposed copy.
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffffu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)64 << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void wait_bar(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// weights [rows][64] fp16 (row-major = K-major) -> swizzled tile
__device__ __forceinline__ void load_weights_sw(unsigned char* dst, const __half* src, int rows) {
    for (int i = threadIdx.x; i < rows * 8; i += blockDim.x) {
        const int r = i >> 3, c = i & 7;
        *reinterpret_cast<uint4*>(dst + sw128(r, c)) = __ldg(reinterpret_cast<const uint4*>(src + r * kW + c * 8));
    }
}

}  // namespace tc

// 256 threads per CTA, two per row of the 128-row tile: thread (row, part) encodes the grid levels / OneBlob
// dimensions of parity `part` (the 32 lanes of a warp are 32 adjacent rows working on the SAME level, so their
// gathers stay as coherent as the rows are), and in the hidden-layer epilogues drains columns [32 part, 32 part + 32)
// of its row's accumulator.  Warps w and w + 4 share TMEM lanes 32 (w & 3) .. + 31.
constexpr int kFwdThreads = 256;

template <bool TRAIN, bool POW2>
__global__ void __launch_bounds__(kFwdThreads, 4) k_mlp_forward_tc(const NetShape S, const __half* __restrict__ params, FwdArgs A) {
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t raw = tc::smem_u32(smem_dyn);
    unsigned char* base = smem_dyn + ((1024u - (raw & 1023u)) & 1023u);
    unsigned char* sA = base;
    unsigned char* sW0 = sA + tc::kABytes;
    unsigned char* sW1 = sW0 + tc::kWBytes;
    unsigned char* sWo = sW1 + tc::kWBytes;
    uint64_t* bar_ptr = reinterpret_cast<uint64_t*>(sWo + tc::kWoBytes);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_ptr + 1);
    const uint32_t bar = tc::smem_u32(bar_ptr);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int part = warp >> 2;
    const uint32_t row = (uint32_t)((warp & 3) * 32 + lane);

    tc::load_weights_sw(sW0, params, kW);
    tc::load_weights_sw(sW1, params + kW * kW, kW);
    tc::load_weights_sw(sWo, params + 2 * kW * kW, kOutPad);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tc::smem_u32(tmem_slot)), "r"(tc::kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc::fence_async_smem();          // weight tiles were written through the generic proxy
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tmem_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint64_t dA = tc::make_desc(tc::smem_u32(sA));
    const uint64_t dW0 = tc::make_desc(tc::smem_u32(sW0));
    const uint64_t dW1 = tc::make_desc(tc::smem_u32(sW1));
    const uint64_t dWo = tc::make_desc(tc::smem_u32(sWo));
    const uint32_t idesc64 = tc::make_idesc(64), idesc16 = tc::make_idesc(16);
    const __half2* table = reinterpret_cast<const __half2*>(params + 2 * kW * kW + kOutPad * kW);
    uint32_t parity = 0;

    // one layer on the tensor core: D[128 x n] = A[128 x 64] * W[n x 64]^T, K in 4 steps of 16
    auto layer = [&](uint64_t dW, uint32_t idesc) {
        tc::fence_async_smem();      // this thread's A-tile writes -> async proxy
        tc::fence_before();
        __syncthreads();
        if (threadIdx.x == 0) {
            tc::fence_after();
#pragma unroll
            for (uint32_t kk = 0; kk < 4; ++kk) tc::mma_f16(tmem, dA + 2 * kk, dW + 2 * kk, idesc, kk);
            tc::commit(bar);
        }
        tc::wait_bar(bar, parity);
        parity ^= 1u;
        tc::fence_after();
    };
    // this thread's 32 accumulator columns of its row -> ReLU -> fp16 -> A tile (and optionally global)
    auto epilogue_hidden = [&](__half* gdst) {
        uint32_t v[32];
        tc::ld32(tmem_lane + (uint32_t)part * 32u, v);
        tc::wait_ld();
        uint4 packed[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            uint32_t w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                w[j] = pack2(fmaxf(__uint_as_float(v[c * 8 + 2 * j]), 0.f), fmaxf(__uint_as_float(v[c * 8 + 2 * j + 1]), 0.f));
            packed[c] = make_uint4(w[0], w[1], w[2], w[3]);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(sA + tc::sw128(row, (uint32_t)(4 * part + c))) = packed[c];
        if (gdst) {
            uint4* d = reinterpret_cast<uint4*>(gdst) + 4 * part;
#pragma unroll
            for (int c = 0; c < 4; ++c) d[c] = packed[c];
        }
    };

    for (int tile = blockIdx.x; tile < A.n_tiles; tile += gridDim.x) {
        if (A.tile_mask && __ldg(A.tile_mask + tile) == 0) continue;
        const size_t grow = (size_t)tile * kTile + row;
        {
            float x[12];
            const float* src = A.in + grow * S.in_ch;
            if (S.in_ch == 12) {
                const float4* s4 = reinterpret_cast<const float4*>(src);
                float4 a = __ldg(s4), b = __ldg(s4 + 1), c = __ldg(s4 + 2);
                x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
                x[8] = c.x; x[9] = c.y; x[10] = c.z; x[11] = c.w;
            } else {
                for (int i = 0; i < 12; ++i) x[i] = i < S.in_ch ? __ldg(src + i) : 0.f;
            }
            encode_row<POW2>(S, table, x, [sA, row](int col) {
                return reinterpret_cast<__half*>(sA + tc::sw128(row, (uint32_t)col >> 3) + ((uint32_t)col & 7u) * 2u);
            }, part, 2);
        }
        if (TRAIN) {
            // the encoded row goes to global memory for the backward pass: each thread copies its half once both
            // parts are in shared memory
            __syncthreads();
            uint4* d = reinterpret_cast<uint4*>(A.e + grow * kW) + 4 * part;
#pragma unroll
            for (int c = 0; c < 4; ++c) d[c] = *reinterpret_cast<const uint4*>(sA + tc::sw128(row, (uint32_t)(4 * part + c)));
        }
        layer(dW0, idesc64);
        epilogue_hidden(TRAIN ? A.h1 + grow * kW : nullptr);
        layer(dW1, idesc64);
        epilogue_hidden(TRAIN ? A.h2 + grow * kW : nullptr);
        layer(dWo, idesc16);
        if (part == 0) {
            uint32_t v[16];
            tc::ld16(tmem_lane, v);
            tc::wait_ld();
            // network output is held in fp16 by the reference (trim_and_cast_from)
            const float pr = __half2float(__float2half_rn(__uint_as_float(v[0])));
            const float pg = __half2float(__float2half_rn(__uint_as_float(v[1])));
            const float pb = __half2float(__float2half_rn(__uint_as_float(v[2])));
            if (!TRAIN) {
                A.out[grow * 3 + 0] = pr; A.out[grow * 3 + 1] = pg; A.out[grow * 3 + 2] = pb;
            } else {
                const float lum = 0.299f * pr + 0.587f * pg + 0.114f * pb;
                const float denom = lum * lum + 0.01f;
                const float df0 = pr - __ldg(A.target + grow * 3 + 0);
                const float df1 = pg - __ldg(A.target + grow * 3 + 1);
                const float df2 = pb - __ldg(A.target + grow * 3 + 2);
                float lsum = df0 * df0 / denom * A.inv_n_total + df1 * df1 / denom * A.inv_n_total + df2 * df2 / denom * A.inv_n_total;
                uint2 dy;
                dy.x = pack2(kLossScale * (2 * df0 / denom) * A.inv_n_total, kLossScale * (2 * df1 / denom) * A.inv_n_total);
                dy.y = pack2(kLossScale * (2 * df2 / denom) * A.inv_n_total, 0.f);
                *reinterpret_cast<uint2*>(A.dy + grow * 4) = dy;
#pragma unroll
                for (int ofs = 16; ofs > 0; ofs >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, ofs);
                if (lane == 0) atomicAdd(A.loss, lsum);
            }
        }
        // the next tile's encode overwrites A and its first MMA overwrites the accumulator:
        // order them after every thread's TMEM reads
        tc::fence_before();
        __syncthreads();
    }
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(tc::kTmemCols) : "memory");
    }
}

// ---- fused training step (forward + loss + backward + weight gradients), all on tcgen05 -----------------------
// tiny-cuda-nn runs kernel_mlp_fused (forward, src/fully_fused_mlp.cu:499-557), the loss kernel, kernel_mlp_fused_backward
// (:150-259), three split-K CUTLASS GEMMs for the weight gradients (:784-836) and kernel_grid_backward
// (encodings/grid.h:395-516).  Here one persistent CTA per SM takes a 128-row tile through all of it without the
// activations ever leaving shared memory:
//   e  --W0-->  h1  --W1-->  h2  --Wout-->  y  -> loss, dy
//   dy --Wout(MN-major)--> dh2 (masked by h2 > 0) --W1(MN-major)--> dh1 (masked by h1 > 0) --W0(MN-major)--> de -> grid scatter
//   G1[128 x 128] += [dh2 | dh1]^T * [h1 | e]   (both operands MN-major views of the activation tiles; the diagonal
//                                                64 x 64 blocks are dW1 and dW0, the off-diagonal blocks are not used)
//   G2[128 x 64]  += [dy | dy]^T * h2           (rows 0..2 are dWout)
// Nine tcgen05.mma groups per tile, every one M = 128; G1 / G2 stay in TMEM (fp32) for the CTA's lifetime and are
// added to the global gradient buffer once at the end.
namespace tc {
constexpr uint32_t kTile16K = 128 * 128;
constexpr uint32_t kTrainSmem = 6 * kTile16K + 2 * kWBytes + kWoBytes + 64 + 1024;
constexpr uint32_t kTrainTmemCols = 256;   // acc 64 | G1 128 | G2 64
}

template <bool POW2>
__global__ void __launch_bounds__(kFwdThreads, 1) k_mlp_train_fused(const NetShape S, const __half* __restrict__ params, FwdArgs A,
                                                                     float* __restrict__ grads, float* __restrict__ grid_grads) {
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t raw = tc::smem_u32(smem_dyn);
    unsigned char* base = smem_dyn + ((1024u - (raw & 1023u)) & 1023u);
    unsigned char* sH1 = base;                          // [h1 | e] : B operand of G1 (blocks 16 KB apart)
    unsigned char* sE = sH1 + tc::kTile16K;
    unsigned char* sDH2 = sE + tc::kTile16K;            // [dh2 | dh1] : A operand of G1
    unsigned char* sDH1 = sDH2 + tc::kTile16K;
    unsigned char* sH2 = sDH1 + tc::kTile16K;
    unsigned char* sDY = sH2 + tc::kTile16K;
    unsigned char* sW0 = sDY + tc::kTile16K;
    unsigned char* sW1 = sW0 + tc::kWBytes;
    unsigned char* sWo = sW1 + tc::kWBytes;
    uint64_t* bar_ptr = reinterpret_cast<uint64_t*>(sWo + tc::kWoBytes);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_ptr + 1);
    const uint32_t bar = tc::smem_u32(bar_ptr);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int part = warp >> 2;
    const uint32_t row = (uint32_t)((warp & 3) * 32 + lane);

    tc::load_weights_sw(sW0, params, kW);
    tc::load_weights_sw(sW1, params + kW * kW, kW);
    tc::load_weights_sw(sWo, params + 2 * kW * kW, kOutPad);
    // dy tile: only columns 0..15 are ever written; the rest must read as zero (it is an operand of G2)
    for (uint32_t i = threadIdx.x; i < tc::kTile16K / 16; i += blockDim.x) reinterpret_cast<uint4*>(sDY)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tc::smem_u32(tmem_slot)), "r"(tc::kTrainTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc::fence_async_smem();
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tmem_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t tG1 = tmem + 64, tG2 = tmem + 192;
    // K-major views (forward operands, and the A operands of the dgrad products)
    const uint64_t kE = tc::make_desc(tc::smem_u32(sE)), kH1 = tc::make_desc(tc::smem_u32(sH1)), kH2 = tc::make_desc(tc::smem_u32(sH2));
    const uint64_t kDY = tc::make_desc(tc::smem_u32(sDY)), kDH2 = tc::make_desc(tc::smem_u32(sDH2)), kDH1 = tc::make_desc(tc::smem_u32(sDH1));
    const uint64_t kW0 = tc::make_desc(tc::smem_u32(sW0)), kW1 = tc::make_desc(tc::smem_u32(sW1)), kWo = tc::make_desc(tc::smem_u32(sWo));
    // MN-major views: weights as B of the dgrad products, activation tiles as both operands of the weight gradients
    const uint64_t mW0 = tc::make_desc_mn(tc::smem_u32(sW0), 0), mW1 = tc::make_desc_mn(tc::smem_u32(sW1), 0), mWo = tc::make_desc_mn(tc::smem_u32(sWo), 0);
    const uint64_t mD = tc::make_desc_mn(tc::smem_u32(sDH2), tc::kTile16K);     // [dh2 | dh1], M = 128
    const uint64_t mX = tc::make_desc_mn(tc::smem_u32(sH1), tc::kTile16K);      // [h1 | e],   N = 128
    const uint64_t mDY = tc::make_desc_mn(tc::smem_u32(sDY), 0);                // [dy | dy],  M = 128 (both blocks the same tile)
    const uint64_t mH2 = tc::make_desc_mn(tc::smem_u32(sH2), 0);
    const uint32_t i64 = tc::make_idesc(64), i16 = tc::make_idesc(16);
    const uint32_t i64_bmn = tc::make_idesc(64, 0, 1);
    const uint32_t i128_mn = tc::make_idesc(128, 1, 1), i64_mn = tc::make_idesc(64, 1, 1);
    const __half2* table = reinterpret_cast<const __half2*>(params + 2 * kW * kW + kOutPad * kW);
    uint32_t parity = 0;
    int tiles_done = 0;

    // publish this thread's shared-memory writes to the async proxy, meet, let one thread issue, wait for the commit
    auto run = [&](auto&& issue) {
        tc::fence_async_smem();
        tc::fence_before();
        __syncthreads();
        if (threadIdx.x == 0) {
            tc::fence_after();
            issue();
            tc::commit(bar);
        }
        tc::wait_bar(bar, parity);
        parity ^= 1u;
        tc::fence_after();
    };
    auto pack_row = [&](unsigned char* dst, const uint32_t (&v)[32], const unsigned char* mask_tile, bool relu) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint32_t chunk = (uint32_t)(4 * part + c);
            uint32_t w[4];
            uint4 m = make_uint4(0, 0, 0, 0);
            if (mask_tile) m = *reinterpret_cast<const uint4*>(mask_tile + tc::sw128(row, chunk));
            const uint32_t mw[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float lo = __uint_as_float(v[c * 8 + 2 * j]), hi = __uint_as_float(v[c * 8 + 2 * j + 1]);
                if (relu) { lo = fmaxf(lo, 0.f); hi = fmaxf(hi, 0.f); }
                if (mask_tile) {
                    const __half2 h = *reinterpret_cast<const __half2*>(&mw[j]);
                    if (!(__low2float(h) > 0.f)) lo = 0.f;
                    if (!(__high2float(h) > 0.f)) hi = 0.f;
                }
                w[j] = pack2(lo, hi);
            }
            *reinterpret_cast<uint4*>(dst + tc::sw128(row, chunk)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    };
    auto drain_to = [&](unsigned char* dst, const unsigned char* mask_tile, bool relu) {
        uint32_t v[32];
        tc::ld32(tmem_lane + (uint32_t)part * 32u, v);
        tc::wait_ld();
        pack_row(dst, v, mask_tile, relu);
    };

    for (int tile = blockIdx.x; tile < A.n_tiles; tile += gridDim.x, ++tiles_done) {
        const size_t grow = (size_t)tile * kTile + row;
        float x[12];
        {
            const float* src = A.in + grow * S.in_ch;
            if (S.in_ch == 12) {
                const float4* s4 = reinterpret_cast<const float4*>(src);
                float4 a = __ldg(s4), b = __ldg(s4 + 1), c = __ldg(s4 + 2);
                x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
                x[8] = c.x; x[9] = c.y; x[10] = c.z; x[11] = c.w;
            } else {
                for (int i = 0; i < 12; ++i) x[i] = i < S.in_ch ? __ldg(src + i) : 0.f;
            }
            encode_row<POW2>(S, table, x, [sE, row](int col) {
                return reinterpret_cast<__half*>(sE + tc::sw128(row, (uint32_t)col >> 3) + ((uint32_t)col & 7u) * 2u);
            }, part, 2);
        }
        // ---- forward ----
        run([&] {
#pragma unroll
            for (uint32_t kk = 0; kk < 4; ++kk) tc::mma_f16(tmem, kE + 2 * kk, kW0 + 2 * kk, i64, kk);
        });
        drain_to(sH1, nullptr, true);
        run([&] {
#pragma unroll
            for (uint32_t kk = 0; kk < 4; ++kk) tc::mma_f16(tmem, kH1 + 2 * kk, kW1 + 2 * kk, i64, kk);
        });
        drain_to(sH2, nullptr, true);
        run([&] {
#pragma unroll
            for (uint32_t kk = 0; kk < 4; ++kk) tc::mma_f16(tmem, kH2 + 2 * kk, kWo + 2 * kk, i16, kk);
        });
        // ---- loss (relative_l2_luminance.h:40-87) and dL/dy ----
        if (part == 0) {
            uint32_t v[16];
            tc::ld16(tmem_lane, v);
            tc::wait_ld();
            const float pr = __half2float(__float2half_rn(__uint_as_float(v[0])));
            const float pg = __half2float(__float2half_rn(__uint_as_float(v[1])));
            const float pb = __half2float(__float2half_rn(__uint_as_float(v[2])));
            const float lum = 0.299f * pr + 0.587f * pg + 0.114f * pb;
            const float denom = lum * lum + 0.01f;
            const float df0 = pr - __ldg(A.target + grow * 3 + 0);
            const float df1 = pg - __ldg(A.target + grow * 3 + 1);
            const float df2 = pb - __ldg(A.target + grow * 3 + 2);
            float lsum = df0 * df0 / denom * A.inv_n_total + df1 * df1 / denom * A.inv_n_total + df2 * df2 / denom * A.inv_n_total;
            uint4 d0;
            d0.x = pack2(kLossScale * (2 * df0 / denom) * A.inv_n_total, kLossScale * (2 * df1 / denom) * A.inv_n_total);
            d0.y = pack2(kLossScale * (2 * df2 / denom) * A.inv_n_total, 0.f);
            d0.z = 0u; d0.w = 0u;
            *reinterpret_cast<uint4*>(sDY + tc::sw128(row, 0)) = d0;      // columns 0..7; columns 8..63 stay zero
#pragma unroll
            for (int ofs = 16; ofs > 0; ofs >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, ofs);
            if (lane == 0) atomicAdd(A.loss, lsum);
        }
        // ---- backward: dh2 = (dy * Wout) . [h2 > 0] ----
        run([&] { tc::mma_f16(tmem, kDY, mWo, i64_bmn, 0); });           // K = 16: the 16 rows of the Wout tile
        drain_to(sDH2, sH2, false);
        // dh1 = (dh2 * W1) . [h1 > 0]
        run([&] {
#pragma unroll
            for (uint32_t kk = 0; kk < 4; ++kk) tc::mma_f16(tmem, kDH2 + 2 * kk, mW1 + 128 * kk, i64_bmn, kk);   // B: 16 rows = 2048 B per K step
        });
        drain_to(sDH1, sH1, false);
        // de = dh1 * W0, and the weight gradients of this tile (all operands are complete now)
        const uint32_t acc_first = tiles_done > 0 ? 1u : 0u;
        run([&] {
#pragma unroll
            for (uint32_t kk = 0; kk < 4; ++kk) tc::mma_f16(tmem, kDH1 + 2 * kk, mW0 + 128 * kk, i64_bmn, kk);
#pragma unroll
            for (uint32_t kk = 0; kk < 8; ++kk) tc::mma_f16(tG1, mD + 128 * kk, mX + 128 * kk, i128_mn, kk ? 1u : acc_first);
#pragma unroll
            for (uint32_t kk = 0; kk < 8; ++kk) tc::mma_f16(tG2, mDY + 128 * kk, mH2 + 128 * kk, i64_mn, kk ? 1u : acc_first);
        });
        // ---- hash-grid scatter (kernel_grid_backward, grid.h:395-516): this thread's levels of its row ----
        {
            uint32_t v[32];
            tc::ld32(tmem_lane, v);          // columns 0..31 = the 16 levels x 2 features
            tc::wait_ld();
            const int nl = S.grid.n_levels;
#pragma unroll
            for (int l2 = 0; l2 < kMlpMaxLevels / 2; ++l2) {
                const int l = 2 * l2 + part;
                if (l >= nl) continue;
                // the reference holds dL/d(encoded) in fp16 (fc_multiply output)
                const float g0 = __half2float(__float2half_rn(__uint_as_float(v[2 * l])));
                const float g1 = __half2float(__float2half_rn(__uint_as_float(v[2 * l + 1])));
                if (g0 == 0.f && g1 == 0.f) continue;
                const float scale = S.grid.scale[l];
                const uint32_t res = S.grid.resolution[l];
                const uint32_t size = S.grid.offset[l + 1] - S.grid.offset[l];
                const bool hashed = (S.hashed_mask >> l) & 1u;
                float* gl = grid_grads + 2 * (size_t)S.grid.offset[l];
                float p0 = x[0] * scale + 0.5f, p1 = x[1] * scale + 0.5f, p2 = x[2] * scale + 0.5f;
                const float f0 = floorf(p0), f1 = floorf(p1), f2 = floorf(p2);
                const uint32_t q0 = (uint32_t)(int)f0, q1 = (uint32_t)(int)f1, q2 = (uint32_t)(int)f2;
                p0 -= f0; p1 -= f1; p2 -= f2;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    float w = ((c & 1) ? p0 : 1 - p0) * ((c & 2) ? p1 : 1 - p1);
                    w *= (c & 4) ? p2 : 1 - p2;
                    const uint32_t idx = grid_cell_index(q0 + (c & 1), q1 + ((c >> 1) & 1), q2 + ((c >> 2) & 1), res, size, hashed);
                    // per-contribution fp16 rounding as in kernel_grid_backward's half2 atomics
                    atomicAdd(gl + 2 * (size_t)idx + 0, __half2float(__float2half_rn(g0 * w)));
                    atomicAdd(gl + 2 * (size_t)idx + 1, __half2float(__float2half_rn(g1 * w)));
                }
            }
        }
        // the next tile's encode overwrites e and its MMAs the accumulator: order them after every thread's TMEM reads
        tc::fence_before();
        __syncthreads();
    }
    // ---- weight gradients: TMEM -> global (fp32 atomics; one add per CTA and element) ----
    if (tiles_done > 0) {
        tc::fence_after();
        // G1 row m, columns n: m < 64, n < 64 -> dW1[m][n];  m >= 64, n >= 64 -> dW0[m - 64][n - 64]
        const bool want = (row < 64u) == (part == 0);
        if (want) {
            float* dst = row < 64u ? grads + kW * kW + (size_t)row * kW : grads + (size_t)(row - 64u) * kW;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t v[32];
                tc::ld32(tG1 + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)part * 64u + (uint32_t)h * 32u, v);
                tc::wait_ld();
#pragma unroll
                for (int j = 0; j < 32; ++j) atomicAdd(dst + h * 32 + j, __uint_as_float(v[j]));
            }
        }
        // G2 rows 0..2 = dWout (rows 3..15 of the padded matrix get no gradient: their dy columns are zero)
        if (warp == 0 || warp == 4) {
            uint32_t v[32];
            tc::ld32(tG2 + (uint32_t)part * 32u, v);      // warps 0 and 4 both own lanes 0..31; columns by part
            tc::wait_ld();
            if (lane < 3) {
                float* dst = grads + 2 * kW * kW + (size_t)lane * kW + part * 32;
#pragma unroll
                for (int j = 0; j < 32; ++j) atomicAdd(dst + j, __uint_as_float(v[j]));
            }
        }
    }
    tc::fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(tc::kTrainTmemCols) : "memory");
    }
}

// ---- backward ----------------------------------------------------------------------
struct BwdArgs {
    const float* in;       // [n][in_ch] (positions for the grid scatter)
    const __half* dy;      // [n][4]
    const __half* h1; const __half* h2;   // [n][64]
    __half* dh1; __half* dh2;             // [n][64] out
    float* grid_grads;     // fp32 [entries][2]
    int n_tiles;
};

__global__ void __launch_bounds__(kTile) k_mlp_backward(const NetShape S, const __half* __restrict__ params, BwdArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __half* sW0T = reinterpret_cast<__half*>(smem_raw);       // [in 64][out 64]
    __half* sW1T = sW0T + kW * kStride;                         // [in 64][out 64]
    __half* sWoT = sW1T + kW * kStride;                         // [hidden 64][out 16]
    __half* sAct = sWoT + kW * kStride;
    load_matrix_t(sW0T, params, kW);
    load_matrix_t(sW1T, params + kW * kW, kW);
    load_matrix_t(sWoT, params + 2 * kW * kW, kOutPad);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    __half* wAct = sAct + warp * 32 * kStride;

    for (int tile = blockIdx.x; tile < A.n_tiles; tile += gridDim.x) {
        const size_t wrow = (size_t)tile * kTile + warp * 32;
        // stage dy into the activation tile: columns 0..3 from memory, 4..15 zero
        {
            const size_t r = wrow + lane;
            uint2 v = __ldg(reinterpret_cast<const uint2*>(A.dy + r * 4));
            uint32_t* d = reinterpret_cast<uint32_t*>(wAct + lane * kStride);
            d[0] = v.x; d[1] = v.y;
#pragma unroll
            for (int i = 2; i < 8; ++i) d[i] = 0u;
        }
        __syncwarp();
        float acc[2][8][4];
        warp_gemm<8>(wAct, sWoT, acc, 1);    // dh2 = dy * Wout   (K = 16)
        __syncwarp();
        // ReLU mask from h2, store
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const size_t r = wrow + mt * 16 + hh * 8 + g;
                    __half2 h = *reinterpret_cast<const __half2*>(A.h2 + r * kW + nt * 8 + t * 2);
                    if (!(__low2float(h) > 0.f)) acc[mt][nt][hh * 2 + 0] = 0.f;
                    if (!(__high2float(h) > 0.f)) acc[mt][nt][hh * 2 + 1] = 0.f;
                }
        warp_store_act<8, false>(wAct, acc);
        __syncwarp();
        {
            const uint4* s = reinterpret_cast<const uint4*>(wAct + lane * kStride);
            uint4* d = reinterpret_cast<uint4*>(A.dh2 + (wrow + lane) * kW);
#pragma unroll
            for (int i = 0; i < 8; ++i) d[i] = s[i];
        }
        warp_gemm<8>(wAct, sW1T, acc);       // dh1 = dh2 * W1
        __syncwarp();
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const size_t r = wrow + mt * 16 + hh * 8 + g;
                    __half2 h = *reinterpret_cast<const __half2*>(A.h1 + r * kW + nt * 8 + t * 2);
                    if (!(__low2float(h) > 0.f)) acc[mt][nt][hh * 2 + 0] = 0.f;
                    if (!(__high2float(h) > 0.f)) acc[mt][nt][hh * 2 + 1] = 0.f;
                }
        warp_store_act<8, false>(wAct, acc);
        __syncwarp();
        {
            const uint4* s = reinterpret_cast<const uint4*>(wAct + lane * kStride);
            uint4* d = reinterpret_cast<uint4*>(A.dh1 + (wrow + lane) * kW);
#pragma unroll
            for (int i = 0; i < 8; ++i) d[i] = s[i];
        }
        // de = dh1 * W0, only the hash-grid columns (first 2 * n_levels) are trainable inputs
        float de[2][4][4];
        warp_gemm<4>(wAct, sW0T, de);
        __syncwarp();
        // scatter: this lane holds de for rows (mt, hh) and level = nt * 4 + t (two features)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const size_t r = wrow + mt * 16 + hh * 8 + g;
                const float* src = A.in + r * S.in_ch;
                const float x0 = __ldg(src), x1 = __ldg(src + 1), x2 = __ldg(src + 2);
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    const int l = nt * 4 + t;
                    if (l >= S.grid.n_levels) continue;
                    // the reference holds dL/d(encoded) in fp16 (fc_multiply output)
                    const float g0 = __half2float(__float2half_rn(de[mt][nt][hh * 2 + 0]));
                    const float g1 = __half2float(__float2half_rn(de[mt][nt][hh * 2 + 1]));
                    if (g0 == 0.f && g1 == 0.f) continue;
                    const float scale = S.grid.scale[l];
                    const uint32_t res = S.grid.resolution[l];
                    const uint32_t size = S.grid.offset[l + 1] - S.grid.offset[l];
                    const bool hashed = (S.hashed_mask >> l) & 1u;
                    float* gl = A.grid_grads + 2 * (size_t)S.grid.offset[l];
                    float p0 = x0 * scale + 0.5f, p1 = x1 * scale + 0.5f, p2 = x2 * scale + 0.5f;
                    float f0 = floorf(p0), f1 = floorf(p1), f2 = floorf(p2);
                    uint32_t q0 = (uint32_t)(int)f0, q1 = (uint32_t)(int)f1, q2 = (uint32_t)(int)f2;
                    p0 -= f0; p1 -= f1; p2 -= f2;
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        float w = ((c & 1) ? p0 : 1 - p0) * ((c & 2) ? p1 : 1 - p1);
                        w *= (c & 4) ? p2 : 1 - p2;
                        uint32_t idx = grid_cell_index(q0 + (c & 1), q1 + ((c >> 1) & 1), q2 + ((c >> 2) & 1), res, size, hashed);
                        // per-contribution fp16 rounding as in kernel_grid_backward's half2 atomics
                        atomicAdd(gl + 2 * (size_t)idx + 0, __half2float(__float2half_rn(g0 * w)));
                        atomicAdd(gl + 2 * (size_t)idx + 1, __half2float(__float2half_rn(g1 * w)));
                    }
                }
            }
        __syncwarp();
    }
}

// dW[o][i] += sum_r D[r][o] * Aact[r][i] over a chunk of rows; D, Aact are [n][64] fp16.
__global__ void __launch_bounds__(256) k_wgrad(const __half* __restrict__ D, const __half* __restrict__ Aact, float* __restrict__ dW,
                                               int n, int rows_per_cta) {
    __shared__ __align__(16) __half sD[64][kW + 8];
    __shared__ __align__(16) __half sA[64][kW + 8];
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;   // 16 x 16 threads, 4 x 4 outputs each
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int r_begin = blockIdx.x * rows_per_cta;
    const int r_end = min(n, r_begin + rows_per_cta);
    for (int r0 = r_begin; r0 < r_end; r0 += 64) {
        for (int i = threadIdx.x; i < 64 * 8; i += 256) {
            int r = i >> 3, c = i & 7;
            uint4 vd = make_uint4(0, 0, 0, 0), va = make_uint4(0, 0, 0, 0);
            if (r0 + r < r_end) {
                vd = __ldg(reinterpret_cast<const uint4*>(D + (size_t)(r0 + r) * kW + c * 8));
                va = __ldg(reinterpret_cast<const uint4*>(Aact + (size_t)(r0 + r) * kW + c * 8));
            }
            *reinterpret_cast<uint4*>(&sD[r][c * 8]) = vd;
            *reinterpret_cast<uint4*>(&sA[r][c * 8]) = va;
        }
        __syncthreads();
#pragma unroll 8
        for (int r = 0; r < 64; ++r) {
            float2 d01 = __half22float2(*reinterpret_cast<const __half2*>(&sD[r][ty * 4]));
            float2 d23 = __half22float2(*reinterpret_cast<const __half2*>(&sD[r][ty * 4 + 2]));
            float2 a01 = __half22float2(*reinterpret_cast<const __half2*>(&sA[r][tx * 4]));
            float2 a23 = __half22float2(*reinterpret_cast<const __half2*>(&sA[r][tx * 4 + 2]));
            const float d[4] = {d01.x, d01.y, d23.x, d23.y};
            const float a[4] = {a01.x, a01.y, a23.x, a23.y};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] += d[i] * a[j];
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) atomicAdd(dW + (ty * 4 + i) * kW + tx * 4 + j, acc[i][j]);
}

// dWout[o][i] += sum_r dy[r][o] * h2[r][i], o < 4 (rows 3.. of the padded matrix get no gradient)
__global__ void __launch_bounds__(256) k_wgrad_out(const __half* __restrict__ dy, const __half* __restrict__ h2, float* __restrict__ dW,
                                                   int n, int rows_per_cta) {
    const int o = threadIdx.x >> 6, i = threadIdx.x & 63;
    const int r_begin = blockIdx.x * rows_per_cta;
    const int r_end = min(n, r_begin + rows_per_cta);
    float acc = 0.f;
    for (int r = r_begin; r < r_end; ++r)
        acc += __half2float(__ldg(dy + (size_t)r * 4 + o)) * __half2float(__ldg(h2 + (size_t)r * kW + i));
    atomicAdd(dW + o * kW + i, acc);
}

// adam_step (optimizers/adam.h:48-120) on fp32 gradients rounded to fp16 first (tcnn's gradient
// buffer is __half); clears the gradient for the next step.
__global__ void __launch_bounds__(256) k_adam(size_t n, size_t n_matrix, float lr, float beta1, float beta2, float eps, float l2_reg,
                                              float* __restrict__ master, __half* __restrict__ half_w, float* __restrict__ grads,
                                              float* __restrict__ m1, float* __restrict__ m2, uint32_t* __restrict__ steps) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float gradient = __half2float(__float2half_rn(grads[i])) / kLossScale;
        grads[i] = 0.f;
        if (i >= n_matrix && gradient == 0.f) continue;
        const float w = master[i];
        if (i < n_matrix) gradient += l2_reg * w;
        const float g2 = gradient * gradient;
        const float fm = m1[i] = beta1 * m1[i] + (1 - beta1) * gradient;
        const float sm = m2[i] = beta2 * m2[i] + (1 - beta2) * g2;
        const uint32_t step = ++steps[i];
        const float lr_t = lr * sqrtf(1 - powf(beta2, (float)step)) / (1 - powf(beta1, (float)step));
        const float eff = lr_t / (sqrtf(sm) + eps);
        const float nw = w - eff * fm;
        master[i] = nw;
        half_w[i] = __float2half_rn(nw);
    }
}

__global__ void k_float_to_half(const float* __restrict__ src, __half* __restrict__ dst, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = __float2half_rn(src[i]);
}

// pcg32 (dependencies/pcg32/pcg32.h), restated
struct Pcg32 {
    uint64_t state = 0, inc = 0;
    explicit Pcg32(uint64_t initstate, uint64_t initseq = 1) {
        state = 0; inc = (initseq << 1u) | 1u;
        next_uint(); state += initstate; next_uint();
    }
    uint32_t next_uint() {
        uint64_t old = state;
        state = old * 0x5851f42d4c957f2dULL + inc;
        uint32_t xs = (uint32_t)(((old >> 18u) ^ old) >> 27u);
        uint32_t rot = (uint32_t)(old >> 59u);
        return (xs >> rot) | (xs << ((~rot + 1u) & 31));
    }
    float next_float() {
        union { uint32_t u; float f; } x;
        x.u = (next_uint() >> 9) | 0x3f800000u;
        return x.f - 1.0f;
    }
};

int sm_count() {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms > 0 ? sms : 148;
}

constexpr size_t kFwdSmem = (size_t)(2 * kW + kOutPad + kTile) * kStride * sizeof(__half);
constexpr size_t kBwdSmem = (size_t)(3 * kW + kTile) * kStride * sizeof(__half);

NetShape make_shape(const MlpConfig& c, const GridLayout& g) {
    NetShape s;
    s.grid = g;
    s.in_ch = c.in_ch;
    s.blob_dims = c.blob_dims; s.blob_bins = c.blob_bins;
    s.identity_dims = c.in_ch - c.grid_dims - c.blob_dims;
    s.feats = c.feats;
    s.hashed_mask = 0;
    s.pow2_mask = 0;
    for (int l = 0; l < g.n_levels; ++l) {
        // grid_index (encodings/grid.h:171-187): dense while the running stride fits the level
        uint64_t stride = 1;
        const uint32_t size = g.offset[l + 1] - g.offset[l];
        for (int d = 0; d < 3 && stride <= size; ++d) stride *= g.resolution[l];
        if (size < stride) s.hashed_mask |= 1u << l;
        if ((size & (size - 1)) == 0) s.pow2_mask |= 1u << l;
    }
    return s;
}

}  // namespace

// ---- configuration -------------------------------------------------------------------
MlpConfig mlp_config_from_json(const std::string& path, int in_ch, int out_ch) {
    MlpConfig c;
    c.in_ch = in_ch; c.out_ch = out_ch;
    Json j = parse_json_file(path);
    auto otype = [](const Json& o) { return o.has("otype") ? o.at("otype").string() : std::string(); };
    if (const Json* loss = j.find("loss"))
        if (otype(*loss) != "RelativeL2Luminance") throw std::invalid_argument("tcnn config: unsupported loss '" + otype(*loss) + "'");
    if (const Json* opt = j.find("optimizer")) {
        const Json* adam = opt;
        if (otype(*opt) == "ExponentialDecay") {
            if (opt->has("decay_start")) c.decay_start = (int)opt->at("decay_start").number();
            if (opt->has("decay_interval")) c.decay_interval = (int)opt->at("decay_interval").number();
            if (c.decay_interval <= 0) throw std::invalid_argument("tcnn optimizer.decay_interval must be positive");
            if (opt->has("decay_base")) c.decay_base = (float)opt->at("decay_base").number();
            adam = &opt->at("nested");
        } else {
            c.decay_start = 1 << 30;
        }
        if (otype(*adam) != "Adam") throw std::invalid_argument("tcnn config: unsupported optimizer '" + otype(*adam) + "'");
        if (adam->has("learning_rate")) c.lr = (float)adam->at("learning_rate").number();
        if (adam->has("beta1")) c.beta1 = (float)adam->at("beta1").number();
        if (adam->has("beta2")) c.beta2 = (float)adam->at("beta2").number();
        if (adam->has("epsilon")) c.eps = (float)adam->at("epsilon").number();
        if (adam->has("l2_reg")) c.l2_reg = (float)adam->at("l2_reg").number();
    }
    if (const Json* enc = j.find("encoding")) {
        if (otype(*enc) != "Composite") throw std::invalid_argument("tcnn config: unsupported encoding '" + otype(*enc) + "'");
        const Json& nested = enc->at("nested");
        if (nested.arr.size() < 2) throw std::invalid_argument("tcnn config: expected HashGrid + OneBlob [+ Identity]");
        const Json& g = nested.arr[0];
        if (otype(g) != "HashGrid") throw std::invalid_argument("tcnn config: first nested encoding must be HashGrid");
        if (g.has("n_dims_to_encode")) c.grid_dims = (int)g.at("n_dims_to_encode").number();
        if (g.has("n_levels")) c.n_levels = (int)g.at("n_levels").number();
        if (g.has("n_features_per_level")) c.feats = (int)g.at("n_features_per_level").number();
        if (g.has("log2_hashmap_size")) c.log2_hashmap = (int)g.at("log2_hashmap_size").number();
        if (g.has("base_resolution")) c.base_res = (int)g.at("base_resolution").number();
        if (g.has("per_level_scale")) c.per_level_scale = (float)g.at("per_level_scale").number();
        const Json& b = nested.arr[1];
        if (otype(b) != "OneBlob") throw std::invalid_argument("tcnn config: second nested encoding must be OneBlob");
        if (b.has("n_dims_to_encode")) c.blob_dims = (int)b.at("n_dims_to_encode").number();
        if (b.has("n_bins")) c.blob_bins = (int)b.at("n_bins").number();
        if (nested.arr.size() > 2 && otype(nested.arr[2]) != "Identity")
            throw std::invalid_argument("tcnn config: third nested encoding must be Identity");
    }
    if (const Json* net = j.find("network")) {
        if (otype(*net) != "FullyFusedMLP") throw std::invalid_argument("tcnn config: unsupported network '" + otype(*net) + "'");
        if (net->has("n_neurons")) c.width = (int)net->at("n_neurons").number();
        if (net->has("n_hidden_layers")) c.hidden_layers = (int)net->at("n_hidden_layers").number();
        if (net->has("activation") && net->at("activation").string() != "ReLU") throw std::invalid_argument("tcnn config: activation must be ReLU");
        if (net->has("output_activation") && net->at("output_activation").string() != "None")
            throw std::invalid_argument("tcnn config: output_activation must be None");
    }
    return c;
}

GridLayout Mlp::grid_layout(const MlpConfig& c) {
    GridLayout g;
    memset(&g, 0, sizeof(g));
    g.n_levels = c.n_levels;
    const float log2_pls = log2f(c.per_level_scale);
    uint32_t offset = 0;
    for (int l = 0; l < c.n_levels; ++l) {
        const float scale = exp2f(l * log2_pls) * c.base_res - 1.0f;          // grid_scale, grid.h:195-200
        const uint32_t res = (uint32_t)ceilf(scale) + 1;                       // grid_resolution, grid.h:202-204
        const uint32_t max_params = 0xffffffffu / 2;
        uint32_t n = powf((float)res, 3.f) > (float)max_params ? max_params : res * res * res;
        n = (n + 7u) / 8u * 8u;
        n = std::min(n, 1u << c.log2_hashmap);
        g.offset[l] = offset;
        g.scale[l] = scale;
        g.resolution[l] = res;
        offset += n;
    }
    g.offset[c.n_levels] = offset;
    return g;
}

void Mlp::initial_params(const MlpConfig& c, std::vector<float>& out, size_t& n_matrix) {
    GridLayout g = grid_layout(c);
    n_matrix = (size_t)kW * kW * 2 + (size_t)kOutPad * kW;
    const size_t n_grid = (size_t)g.offset[c.n_levels] * c.feats;
    out.assign(n_matrix + n_grid, 0.f);
    // Trainer ctor (trainer.h:53-62)
    std::seed_seq seq{c.seed};
    std::vector<uint32_t> seeds(2);
    seq.generate(seeds.begin(), seeds.end());
    Pcg32 rng(seeds.front());
    size_t pos = 0;
    const int shapes[3][2] = {{kW, kW}, {kW, kW}, {kOutPad, kW}};
    for (auto& s : shapes) {
        const float scale = std::sqrt(6.0f / (float)(s[0] + s[1]));               // xavier uniform, gpu_matrix.h:291-305
        for (int i = 0; i < s[0] * s[1]; ++i) out[pos++] = rng.next_float() * 2.0f * scale - scale;
    }
    // generate_random_kernel (random.h:66-94): thread i advances 4*i and writes elements i + T*j
    const size_t n_thr = (n_grid + 3) / 4;
    const size_t T = (n_thr + 127) / 128 * 128;
    std::vector<float> stream(4 * T);
    for (auto& f : stream) f = rng.next_float();
    for (size_t i = 0; i < T; ++i)
        for (size_t j = 0; j < 4; ++j) {
            const size_t idx = i + T * j;
            if (idx >= n_grid) break;
            // val * (upper - lower) + lower: one FMA in tcnn's device code (nvcc contracts it), pinned by tests/golden/tcnn_*.npz
            out[n_matrix + idx] = std::fmaf(stream[4 * i + j], 1e-4f - -1e-4f, -1e-4f);
        }
}

// ---- lifetime ----------------------------------------------------------------------------
Mlp::Mlp(const MlpConfig& cfg, cudaStream_t stream) : cfg_(cfg), stream_(stream) {
    if (cfg.width != kW || cfg.hidden_layers != 2) throw std::invalid_argument("only the 64-neuron, 2-hidden-layer network is built");
    if (cfg.feats != 2 || cfg.grid_dims != 3 || cfg.n_levels > kMlpMaxLevels) throw std::invalid_argument("unsupported hash-grid shape");
    if (cfg.out_ch != 3) throw std::invalid_argument("the network has 3 outputs");
    if (cfg.in_ch > 12 || cfg.in_ch < cfg.grid_dims + cfg.blob_dims) throw std::invalid_argument("unsupported number of input channels");
    if (2 * cfg.n_levels + cfg.blob_dims * cfg.blob_bins + (cfg.in_ch - cfg.grid_dims - cfg.blob_dims) > kW)
        throw std::invalid_argument("encoded width exceeds 64");
    layout_ = grid_layout(cfg);
    std::vector<float> init;
    initial_params(cfg, init, n_matrix_);
    n_params_ = init.size();
    HM_CUDA(cudaMalloc(&d_master_, n_params_ * 4));
    HM_CUDA(cudaMalloc(&d_half_, n_params_ * 2));
    HM_CUDA(cudaMalloc(&d_grads_, n_params_ * 4));
    HM_CUDA(cudaMalloc(&d_m1_, n_params_ * 4));
    HM_CUDA(cudaMalloc(&d_m2_, n_params_ * 4));
    HM_CUDA(cudaMalloc(&d_steps_, n_params_ * 4));
    HM_CUDA(cudaMalloc(&d_loss_, 4));
    HM_CUDA(cudaFuncSetAttribute(k_mlp_forward<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdSmem));
    HM_CUDA(cudaFuncSetAttribute(k_mlp_forward<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdSmem));
    HM_CUDA(cudaFuncSetAttribute(k_mlp_backward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmem));
    HM_CUDA(cudaFuncSetAttribute(k_mlp_train_fused<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kTrainSmem));
    HM_CUDA(cudaFuncSetAttribute(k_mlp_train_fused<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kTrainSmem));
    HM_CUDA(cudaFuncSetAttribute(k_mlp_forward_tc<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kSmemBytes));
    HM_CUDA(cudaFuncSetAttribute(k_mlp_forward_tc<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kSmemBytes));
    HM_CUDA(cudaFuncSetAttribute(k_mlp_forward_tc<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kSmemBytes));
    HM_CUDA(cudaFuncSetAttribute(k_mlp_forward_tc<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::kSmemBytes));
    // forward implementation: tcgen05/TMEM by default; HM_MLP_IMPL=mma selects the mma.sync kernel
    const char* impl = getenv("HM_MLP_IMPL");
    use_tc_ = !(impl && std::string(impl) == "mma");
    // training step: one fused tcgen05 kernel by default; HM_MLP_TRAIN=split selects the forward + mma.sync backward +
    // CUDA-core weight-gradient kernels it replaced (kept for A/B and as the reference point of profiles/)
    const char* tr = getenv("HM_MLP_TRAIN");
    fused_train_ = use_tc_ && !(tr && std::string(tr) == "split");
    reinitialize();
}

Mlp::~Mlp() {
    cudaFree(d_master_); cudaFree(d_half_); cudaFree(d_grads_); cudaFree(d_m1_); cudaFree(d_m2_); cudaFree(d_steps_);
    cudaFree(d_loss_); cudaFree(d_x_); cudaFree(d_h1_); cudaFree(d_h2_); cudaFree(d_dy_);
    cudaFree(d_wpack_);
}

void Mlp::sync_half_params(bool) {
    k_float_to_half<<<sm_count() * 4, 256, 0, stream_>>>(d_master_, (__half*)d_half_, n_params_);
    launches_++;
}

void Mlp::reinitialize() {
    std::vector<float> init;
    size_t nm;
    initial_params(cfg_, init, nm);
    HM_CUDA(cudaMemcpyAsync(d_master_, init.data(), n_params_ * 4, cudaMemcpyHostToDevice, stream_));
    HM_CUDA(cudaMemsetAsync(d_grads_, 0, n_params_ * 4, stream_));
    HM_CUDA(cudaMemsetAsync(d_m1_, 0, n_params_ * 4, stream_));
    HM_CUDA(cudaMemsetAsync(d_m2_, 0, n_params_ * 4, stream_));
    HM_CUDA(cudaMemsetAsync(d_steps_, 0, n_params_ * 4, stream_));
    HM_CUDA(cudaMemsetAsync(d_loss_, 0, 4, stream_));
    sync_half_params(true);
    HM_CUDA(cudaStreamSynchronize(stream_));
    step_ = 0;
    lr_factor_ = 1.f;
}

void Mlp::reset_weights() {
    // TINY_MLP::reset (cuda/neural_network.cu:17-21): set_params(zeros) — weights only, optimizer state kept
    HM_CUDA(cudaMemsetAsync(d_master_, 0, n_params_ * 4, stream_));
    HM_CUDA(cudaMemsetAsync(d_half_, 0, n_params_ * 2, stream_));
}

void Mlp::get_params(float* host, size_t count) {
    if (count > n_params_) throw std::invalid_argument("parameter count out of range");
    HM_CUDA(cudaStreamSynchronize(stream_));
    HM_CUDA(cudaMemcpy(host, d_master_, count * 4, cudaMemcpyDeviceToHost));
}

void Mlp::set_params(const float* host, size_t count) {
    if (count != n_params_) throw std::invalid_argument("parameter count mismatch");
    HM_CUDA(cudaMemcpyAsync(d_master_, host, count * 4, cudaMemcpyHostToDevice, stream_));
    sync_half_params(true);
    HM_CUDA(cudaStreamSynchronize(stream_));
}

void Mlp::ensure_train_buffers(int n) {
    if (n <= train_cap_) return;
    cudaFree(d_x_); cudaFree(d_h1_); cudaFree(d_h2_); cudaFree(d_dy_);
    // e, h1, h2 forward activations; dh1/dh2 reuse a second half of each allocation
    HM_CUDA(cudaMalloc(&d_x_, (size_t)n * kW * 2));
    HM_CUDA(cudaMalloc(&d_h1_, (size_t)n * kW * 2 * 2));
    HM_CUDA(cudaMalloc(&d_h2_, (size_t)n * kW * 2 * 2));
    HM_CUDA(cudaMalloc(&d_dy_, (size_t)n * 4 * 2));
    train_cap_ = n;
}

// ---- compute -------------------------------------------------------------------------------
// resident CTAs per SM of the persistent forward grid (HM_MLP_CTAS overrides; swept on B200)
static int fwd_ctas_per_sm() {
    static int v = 0;
    if (!v) {
        v = 4;
        if (const char* e = getenv("HM_MLP_CTAS")) v = std::max(1, std::min(16, atoi(e)));
    }
    return v;
}

void Mlp::inference(const float* d_in, float* d_out, int n, const int* d_tile_mask) {
    if (n % kTile != 0) throw std::invalid_argument("batch size must be a multiple of 128");
    NetShape S = make_shape(cfg_, layout_);
    FwdArgs A;
    memset(&A, 0, sizeof(A));
    A.in = d_in; A.out = d_out; A.n_tiles = n / kTile;
    A.tile_mask = d_tile_mask;
    int grid = std::min(A.n_tiles, sm_count() * fwd_ctas_per_sm());
    const bool pow2 = S.pow2_mask == (S.grid.n_levels >= 32 ? 0xffffffffu : (1u << S.grid.n_levels) - 1u);
    if (use_tc_ && pow2) k_mlp_forward_tc<false, true><<<grid, kFwdThreads, tc::kSmemBytes, stream_>>>(S, (const __half*)d_half_, A);
    else if (use_tc_) k_mlp_forward_tc<false, false><<<grid, kFwdThreads, tc::kSmemBytes, stream_>>>(S, (const __half*)d_half_, A);
    else k_mlp_forward<false><<<grid, kTile, kFwdSmem, stream_>>>(S, (const __half*)d_half_, A);
    launches_++;
    HM_CUDA(cudaGetLastError());
}

void Mlp::forward_backward(const float* d_in, const float* d_target, int n, int n_total_records) {
    if (n % kTile != 0) throw std::invalid_argument("batch size must be a multiple of 128");
    NetShape S = make_shape(cfg_, layout_);
    if (fused_train_) {
        HM_CUDA(cudaMemsetAsync(d_loss_, 0, 4, stream_));
        FwdArgs A;
        memset(&A, 0, sizeof(A));
        A.in = d_in; A.target = d_target; A.loss = d_loss_;
        A.inv_n_total = 1.f / (float)((size_t)n_total_records * cfg_.out_ch);
        A.n_tiles = n / kTile;
        const int grid = std::min(A.n_tiles, sm_count());
        const bool pow2 = S.pow2_mask == (S.grid.n_levels >= 32 ? 0xffffffffu : (1u << S.grid.n_levels) - 1u);
        if (pow2) k_mlp_train_fused<true><<<grid, kFwdThreads, tc::kTrainSmem, stream_>>>(S, (const __half*)d_half_, A, d_grads_, d_grads_ + n_matrix_);
        else k_mlp_train_fused<false><<<grid, kFwdThreads, tc::kTrainSmem, stream_>>>(S, (const __half*)d_half_, A, d_grads_, d_grads_ + n_matrix_);
        launches_++;
        HM_CUDA(cudaGetLastError());
        return;
    }
    ensure_train_buffers(n);
    __half* e = (__half*)d_x_;
    __half* h1 = (__half*)d_h1_; __half* dh1 = h1 + (size_t)train_cap_ * kW;
    __half* h2 = (__half*)d_h2_; __half* dh2 = h2 + (size_t)train_cap_ * kW;
    HM_CUDA(cudaMemsetAsync(d_loss_, 0, 4, stream_));
    FwdArgs A;
    memset(&A, 0, sizeof(A));
    A.in = d_in; A.target = d_target; A.e = e; A.h1 = h1; A.h2 = h2; A.dy = (__half*)d_dy_; A.loss = d_loss_;
    A.inv_n_total = 1.f / (float)((size_t)n_total_records * cfg_.out_ch);
    A.n_tiles = n / kTile;
    int grid = std::min(A.n_tiles, sm_count() * 4);
    const bool pow2 = S.pow2_mask == (S.grid.n_levels >= 32 ? 0xffffffffu : (1u << S.grid.n_levels) - 1u);
    if (use_tc_ && pow2) k_mlp_forward_tc<true, true><<<grid, kFwdThreads, tc::kSmemBytes, stream_>>>(S, (const __half*)d_half_, A);
    else if (use_tc_) k_mlp_forward_tc<true, false><<<grid, kFwdThreads, tc::kSmemBytes, stream_>>>(S, (const __half*)d_half_, A);
    else k_mlp_forward<true><<<grid, kTile, kFwdSmem, stream_>>>(S, (const __half*)d_half_, A);
    BwdArgs B;
    B.in = d_in; B.dy = (const __half*)d_dy_; B.h1 = h1; B.h2 = h2; B.dh1 = dh1; B.dh2 = dh2;
    B.grid_grads = d_grads_ + n_matrix_;
    B.n_tiles = n / kTile;
    k_mlp_backward<<<grid, kTile, kBwdSmem, stream_>>>(S, (const __half*)d_half_, B);
    const int rows_per_cta = 256;
    const int ctas = (n + rows_per_cta - 1) / rows_per_cta;
    k_wgrad<<<ctas, 256, 0, stream_>>>(dh1, e, d_grads_, n, rows_per_cta);
    k_wgrad<<<ctas, 256, 0, stream_>>>(dh2, h1, d_grads_ + kW * kW, n, rows_per_cta);
    k_wgrad_out<<<ctas, 256, 0, stream_>>>((const __half*)d_dy_, h2, d_grads_ + 2 * kW * kW, n, rows_per_cta);
    launches_ += 5;
    HM_CUDA(cudaGetLastError());
}

void Mlp::optimizer_step() {
    // ExponentialDecayOptimizer::step (optimizers/exponential_decay.h:60-71)
    if (step_ == 0) lr_factor_ = 1.f;
    if (step_ >= cfg_.decay_start && (step_ - cfg_.decay_start) % cfg_.decay_interval == 0) lr_factor_ *= cfg_.decay_base;
    const float lr = cfg_.lr * lr_factor_;
    step_++;
    k_adam<<<sm_count() * 8, 256, 0, stream_>>>(n_params_, n_matrix_, lr, cfg_.beta1, cfg_.beta2, cfg_.eps, cfg_.l2_reg, d_master_,
                                                  (__half*)d_half_, d_grads_, d_m1_, d_m2_, d_steps_);
    launches_++;
    HM_CUDA(cudaGetLastError());
}

float Mlp::loss() {
    float v = 0.f;
    HM_CUDA(cudaMemcpyAsync(&v, d_loss_, 4, cudaMemcpyDeviceToHost, stream_));
    HM_CUDA(cudaStreamSynchronize(stream_));
    return v;
}

}  // namespace hm
