// Softmax building blocks shared by the single-key-tile tcgen05 attention kernels (attention_tc.cu: N <= 208;
// attention_tc257.cu: CLIP's 257 tokens): 16-bit packing, the chunk-wise row maximum and the chunk-wise exponential that
// writes the probabilities back to tensor memory, and the FMA-pipe exp2.
#pragma once

#include <math.h>

#include <type_traits>

#include "ptx.cuh"

namespace vidil {
namespace attn {

template <typename T>
__device__ __forceinline__ uint32_t pack2(float a, float b);
template <>
__device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
template <>
__device__ __forceinline__ uint32_t pack2<__half>(float a, float b) {
    __half2 v = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}

// Row maximum over W (16 or 32) S values starting at key column c0; columns >= nv are masked or padding.  Four
// independent accumulators: a warp issues in order, so one dependent FMNMX3 chain would expose its latency 16 times a chunk.
template <int W>
__device__ __forceinline__ void chunk_max(const uint32_t (&r)[W], int c0, int nv, float (&mx)[4]) {
    if (c0 + W <= nv) {
#pragma unroll
        for (int i = 0; i < W; i += 8) {
#pragma unroll
            for (int j = 0; j < 4; ++j) mx[j] = ptx::max3(mx[j], __uint_as_float(r[i + 2 * j]), __uint_as_float(r[i + 2 * j + 1]));
        }
    } else {
#pragma unroll
        for (int i = 0; i < W; ++i) mx[i & 3] = fmaxf(mx[i & 3], (c0 + i < nv) ? __uint_as_float(r[i]) : -INFINITY);
    }
}

// exp2(c s - c max) of W (16 or 32) S values -> W/2 columns of packed 16-bit P in TMEM; accumulates the row sums.
// A warp issues in order and the MUFU pipe takes one warp-wide ex2 every 8 clocks, so the consumers of an exponential (the
// row-sum add and the 16-bit pack) are placed LAG pairs behind it in program order: with one softmax warp per sub-partition
// nothing else would cover the MUFU latency, and a consumer right behind its producer stalls the whole warp (measured: 540-600
// clocks per 32-column chunk with the add/pack or the mask select directly after each pair, against the 256 the MUFU pipe
// needs).  LAG is in elements: 8 = 4 pairs = 8 MUFU slots = 64 clocks.
// 2^a for a pair of arguments a <= 0 on the FMA / ALU pipes (no MUFU): a = n + f with n = round(a), f in [-0.5, 0.5];
// 2^f by a degree-4 near-minimax polynomial (rel. err 3.7e-6, two orders below the rounding of a 16-bit P), 2^n by adding n
// to the exponent field (the magic constant 1.5 * 2^23 leaves n in the low mantissa bits of t).  The MUFU pipe takes one
// warp-wide ex2 every 8 clocks and is what bounds the softmax phase once the loop overhead is gone, while a single softmax
// warp per sub-partition leaves most issue slots empty: evaluating POLY of every 16 pairs here moves work from the
// saturated pipe to the idle one (the FlashAttention-4 trick).  ~11 instructions per pair, all packed fp32x2 but the clamp
// and the exponent insert.
__device__ __forceinline__ void exp2_poly2(float& a0, float& a1) {
    a0 = fmaxf(a0, -126.0f);
    a1 = fmaxf(a1, -126.0f);
    float t0 = a0, t1 = a1, r0, r1, f0, f1, p0, p1;
    ptx::add2(t0, t1, 12582912.0f, 12582912.0f);          // t = a + magic
    r0 = t0, r1 = t1;
    ptx::add2(r0, r1, -12582912.0f, -12582912.0f);        // r = round(a)
    ptx::fma2v(f0, f1, r0, r1, -1.0f, -1.0f, a0, a1);     // f = a - r
    ptx::fma2(p0, p1, f0, f1, 9.676037098e-03f, 5.592203565e-02f);
    ptx::fma2v(p0, p1, f0, f1, p0, p1, 2.402210736e-01f, 2.402210736e-01f);
    ptx::fma2v(p0, p1, f0, f1, p0, p1, 6.931210340e-01f, 6.931210340e-01f);
    ptx::fma2v(p0, p1, f0, f1, p0, p1, 1.000000075e+00f, 1.000000075e+00f);
    a0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
    a1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}
// pair p (0..15) of a chunk goes to the polynomial iff this is true: POLY pairs of every 16, evenly spread
template <int POLY>
__host__ __device__ constexpr bool poly_pair(int p) {
    return (p * POLY) % 16 < POLY;
}

template <typename T, int W, bool MASKED, int POLY = 0, int LAG = 8>
__device__ __forceinline__ void chunk_exp_impl(const uint32_t (&r)[W], int c0, int nv, float c, float neg_mxs, uint32_t taddr_p,
                                               float (&sum)[4]) {
    uint32_t pk[W / 2];
    float a[W];
#pragma unroll
    for (int i = 0; i < W; i += 2) ptx::fma2(a[i], a[i + 1], __uint_as_float(r[i]), __uint_as_float(r[i + 1]), c, neg_mxs);
#pragma unroll
    for (int i = 0; i < W + LAG; i += 2) {
        if (i < W) {
            if (poly_pair<POLY>(i >> 1)) {
                exp2_poly2(a[i], a[i + 1]);
            } else {
                a[i] = ptx::ex2_approx(a[i]);
                a[i + 1] = ptx::ex2_approx(a[i + 1]);
            }
        }
        if (i >= LAG) {
            const int j = i - LAG;
            if (MASKED) {  // the select is a consumer too: it sits LAG behind the exponential, not right after it
                if (c0 + j >= nv) a[j] = 0.f;
                if (c0 + j + 1 >= nv) a[j + 1] = 0.f;
            }
            ptx::add2(sum[j & 2], sum[(j & 2) + 1], a[j], a[j + 1]);
            pk[j >> 1] = pack2<T>(a[j], a[j + 1]);
        }
    }
    if constexpr (W == 32)
        ptx::tmem_st_32x32b_x16(taddr_p + (c0 >> 1), pk);
    else
        ptx::tmem_st_32x32b_x8(taddr_p + (c0 >> 1), pk);
}
template <typename T, int W, int POLY>
__device__ __forceinline__ void chunk_exp(const uint32_t (&r)[W], int c0, int nv, float c, float neg_mxs, uint32_t taddr_p,
                                          float (&sum)[4]) {
    // warp-uniform choice: both variants end in a tcgen05.st.sync.aligned, which the whole warp must execute together (under a
    // causal mask nv differs from lane to lane; the masked variant is correct for every lane)
    if (__all_sync(0xffffffffu, c0 + W <= nv))
        chunk_exp_impl<T, W, false, POLY>(r, c0, nv, c, neg_mxs, taddr_p, sum);
    else
        chunk_exp_impl<T, W, true, POLY>(r, c0, nv, c, neg_mxs, taddr_p, sum);
}

}  // namespace attn
}  // namespace vidil
