// Fused softmax attention on tcgen05 for sequences that fit one key tile (N <= 208 tokens, head_dim 64) — the
// ViT-L/16 @224 shape (N = 197).  Reference: models/vit.py:72-83 (Attention.forward); the reference materialises
// [B,H,N,N] scores in HBM, here S and P never leave the SM.
//
// One persistent CTA per SM walks (frame, head) items.  Per item the queries form up to two 128-row tiles:
//   S = Q_t K^T          tcgen05.mma M=128, N=ceil16(keys), K=64;  fp32 S in 208 of a lane's 256 TMEM columns
//   P = exp2(c S - c max)   two threads per query row (one per half of the keys) read S from TMEM twice (max, then
//                           exp/sum), exchange max/sum through shared memory, and write 16-bit P to shared memory in
//                           the no-swizzle K-major UMMA layout
//   O = P V              tcgen05.mma M=128, N=64, K=keys; V is consumed straight from its TMA tile as an MN-major B
//                        operand (no transpose anywhere); fp32 O over S's first 64 TMEM columns
//   out = O / rowsum     -> 16-bit -> swizzled staging -> 3-D TMA store (rows past the frame's last token clipped)
// The CTA runs two independent "lanes" L = 0,1, each with its own MMA-issuing thread, 256 TMEM columns, P buffer and
// 8 softmax warps; lane L takes tile (L + item) & 1, so the 128-row and the 69-row tile alternate between the lanes and
// each lane's S -> softmax -> PV -> output chain overlaps the other lane's.  Warp 0 is the TMA producer (Q/K are
// reloaded as soon as both S MMAs of an item retire, V as soon as both PV MMAs do).
// q/k/v are read in place from the fused-QKV GEMM output [B, N, 3, H, 64]; out is [B, N, H*64] (vit.py:83's
// transpose(1,2).reshape), ready to be the proj GEMM's A operand.
#include <math.h>

#include <type_traits>

#include "kernels.h"
#include "ptx.cuh"

namespace vidil {
namespace {

constexpr int HD = 64;
constexpr int QT = 128;          // query rows per tile (UMMA M)
constexpr int MAX_KEYS = 208;    // 13 x 16: S (fp32) fits 256 TMEM columns, K/V tiles fit one TMA box
constexpr int ATC_THREADS = 640; // warp 0 producer, 1-2 MMA issuers, 3 TMEM allocator, 4-11 / 12-19 softmax groups

constexpr int Q_BYTES = QT * HD * 2;               // 16 KB
constexpr int KV_BYTES = MAX_KEYS * HD * 2;        // 26 KB
constexpr int P_BYTES = (MAX_KEYS / 8) * QT * 16;  // 26 chunks of [128 rows][8 keys]: 52 KB
constexpr int OFF_Q = 0;
constexpr int OFF_K = 2 * Q_BYTES;
constexpr int OFF_V = OFF_K + KV_BYTES;
constexpr int OFF_P = OFF_V + KV_BYTES;
constexpr int OFF_OUT = OFF_P + 2 * P_BYTES;       // 2 lanes x 4 quarters x [32 rows][128 B]
constexpr int OFF_XCH = OFF_OUT + 8 * 4096;        // max / sum exchange: [2 lanes][2 halves][128 rows] x 2 floats
constexpr int OFF_BAR = OFF_XCH + 2 * 2 * 128 * 8;
constexpr int ATC_SMEM = OFF_BAR + 128 + 1024;     // + alignment slack
static_assert(OFF_K % 1024 == 0 && OFF_V % 1024 == 0 && OFF_P % 1024 == 0 && OFF_OUT % 1024 == 0, "tile alignment");
static_assert(ATC_SMEM <= 227 * 1024, "shared memory budget");

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

// Row maximum over `n` (16 or 32) S values starting at column c0; columns >= N are padding.
template <int W>
__device__ __forceinline__ float chunk_max(const uint32_t (&r)[32], int c0, int N, float mx) {
    if (c0 + W <= N) {
#pragma unroll
        for (int i = 0; i < W; i += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(r[i]), __uint_as_float(r[i + 1])));
    } else {
#pragma unroll
        for (int i = 0; i < W; ++i) mx = fmaxf(mx, (c0 + i < N) ? __uint_as_float(r[i]) : -INFINITY);
    }
    return mx;
}

// 2^a for a <= 0 on the FMA / ALU pipes (no MUFU): a = n + f with n = round(a), f in [-0.5, 0.5]; 2^f by a degree-4
// near-minimax polynomial (rel. err 3.7e-6), 2^n by adding n to the exponent field (the magic constant 1.5 * 2^23 leaves n
// in the low mantissa bits of t).  Measured on B200 (ViT-L shape, ncu): evaluating 0 / 1 / 2 / 3 of every 8 exponentials
// this way gives 136.3 / 136.7 / 138.4 / 143.3 us per layer — the MUFU unit (4 exponentials per clock per sub-partition)
// is NOT what bounds the softmax phase, the extra instructions cost more than the MUFU slots they free — so POLY_PER_8 = 0.
__device__ __forceinline__ float exp2_poly(float a) {
    a = fmaxf(a, -126.0f);
    const float t = a + 12582912.0f;
    const float f = a - (t - 12582912.0f);
    float p = fmaf(f, 9.676037098e-03f, 5.592203565e-02f);
    p = fmaf(f, p, 2.402210736e-01f);
    p = fmaf(f, p, 6.931210340e-01f);
    p = fmaf(f, p, 1.000000075e+00f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
constexpr int POLY_PER_8 = 0;

// exp2(c s - c max) of W (16 or 32) S values -> 16-bit P chunks in shared memory; returns the partial row sums.
template <typename T, int W>
__device__ __forceinline__ void chunk_exp(const uint32_t (&r)[32], int c0, int N, float c, float neg_mxs, uint32_t prow,
                                          float& sum0, float& sum1) {
    const bool nomask = (c0 + W <= N);
#pragma unroll
    for (int q = 0; q < W / 8; ++q) {  // 8 keys = one 16-byte chunk of the UMMA A layout
        float p[8];
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
            float a0, a1;
            ptx::fma2(a0, a1, __uint_as_float(r[8 * q + i]), __uint_as_float(r[8 * q + i + 1]), c, neg_mxs);
            p[i] = (i < POLY_PER_8) ? exp2_poly(a0) : ptx::ex2_approx(a0);
            p[i + 1] = (i + 1 < POLY_PER_8) ? exp2_poly(a1) : ptx::ex2_approx(a1);
        }
        if (!nomask) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (c0 + 8 * q + i >= N) p[i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < 8; i += 2) ptx::add2(sum0, sum1, p[i], p[i + 1]);
        ptx::st_shared_v4(prow + ((c0 >> 3) + q) * 2048, pack2<T>(p[0], p[1]), pack2<T>(p[2], p[3]), pack2<T>(p[4], p[5]),
                          pack2<T>(p[6], p[7]));
    }
}

template <typename T>
__global__ void __launch_bounds__(ATC_THREADS, 1)
    attention_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv,
                        const __grid_constant__ CUtensorMap map_out, int n_items, int N, int H, float scale_log2e,
                        long long* trace, int causal) {
    // Developer timeline (vidil_debug_set_trace): CTA 0 stamps clock64 at its synchronisation points for 16 items.
#define ATC_TRACE(role, ev)                                                                    \
    do {                                                                                       \
        if (trace != nullptr && blockIdx.x == 0 && it < 16) trace[((role) * 16 + it) * 8 + (ev)] = clock64(); \
    } while (0)
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    const uint32_t sbase = ptx::smem_u32(smem);

    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* qk_full = bars + 0;
    uint64_t* v_full = bars + 1;
    uint64_t* qk_empty = bars + 2;
    uint64_t* v_empty = bars + 3;
    uint64_t* s_full = bars + 4;   // [2] per lane
    uint64_t* p_full = bars + 6;   // [2]
    uint64_t* o_full = bars + 8;   // [2]
    uint64_t* s_free = bars + 10;  // [2]
    uint64_t* tok = bars + 12;     // [2] "this lane is in the last chunk of its exponentials"
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 14);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int nk16 = (N + 15) & ~15;          // keys padded to the UMMA N / K granularity
    const int n_tiles = (N + QT - 1) / QT;    // 1 or 2
    const int D = H * HD;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&map_q);
        ptx::prefetch_tensormap(&map_kv);
        ptx::prefetch_tensormap(&map_out);
        ptx::mbar_init(qk_full, 1);
        ptx::mbar_init(v_full, 1);
        ptx::mbar_init(qk_empty, n_tiles);  // one commit per lane that issued an S MMA for the item
        ptx::mbar_init(v_empty, n_tiles);
        for (int l = 0; l < 2; ++l) {
            ptx::mbar_init(&s_full[l], 1);
            ptx::mbar_init(&p_full[l], 8);  // lane 0 of each of the lane's 8 softmax warps
            ptx::mbar_init(&o_full[l], 1);
            ptx::mbar_init(&s_free[l], 4);  // one per row quarter, after the pair of warps sharing it has synchronised
            ptx::mbar_init(&tok[l], 8);     // lane 0 of each softmax warp, when it enters its last chunk
        }
        ptx::fence_mbar_init();
    }
    if (warp == 3) ptx::tmem_alloc<1>(tmem_ptr_smem, 512);
    ptx::tcgen05_fence_before();
    __syncthreads();
    ptx::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    // PDL: set-up done; let the next kernel start its own, then wait for the predecessor's qkv before any global access
    ptx::griddep_launch();
    ptx::griddep_wait();

    if (warp == 0) {
        if (lane == 0) {
            // ===================== TMA producer =====================
            int it = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const uint32_t ph = it & 1;
                const int b = item / H, h = item - b * H;
                const int row0 = b * N;
                ATC_TRACE(0, 0);
                ptx::mbar_wait(qk_empty, ph ^ 1);  // previous item's S MMAs have consumed Q and K
                ATC_TRACE(0, 1);
                ptx::mbar_arrive_expect_tx(qk_full, n_tiles * Q_BYTES + nk16 * HD * 2);
                ptx::tma_load_2d(&map_q, qk_full, smem + OFF_Q, h * HD, row0);
                if (n_tiles == 2) ptx::tma_load_2d(&map_q, qk_full, smem + OFF_Q + Q_BYTES, h * HD, row0 + QT);
                ptx::tma_load_2d(&map_kv, qk_full, smem + OFF_K, D + h * HD, row0);
                ptx::mbar_wait(v_empty, ph ^ 1);  // previous item's PV MMAs have consumed V
                ATC_TRACE(0, 2);
                ptx::mbar_arrive_expect_tx(v_full, nk16 * HD * 2);
                ptx::tma_load_2d(&map_kv, v_full, smem + OFF_V, 2 * D + h * HD, row0);
            }
        }
        __syncwarp();
    } else if (warp == 1 || warp == 2) {
        if (lane == 0) {
            // ===================== MMA issuer of lane L =====================
            const int L = warp - 1;
            constexpr bool kIsBf16 = std::is_same<T, __nv_bfloat16>::value;
            const uint32_t idesc_s = ptx::make_idesc_f16(kIsBf16, QT, nk16);
            const uint32_t idesc_o = ptx::make_idesc_f16_bmn(kIsBf16, QT, HD);
            const int ksteps = nk16 / 16;
            const uint32_t tmem_d = tmem_base + L * 256;
            const uint32_t sp = sbase + OFF_P + L * P_BYTES;
            int it = 0, n = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const int t = (n_tiles == 2) ? ((L + it) & 1) : L;  // one tile: lane 0 does every item, lane 1 idles
                if (t >= n_tiles) continue;
                const uint32_t ph = it & 1, par = n & 1;
                ++n;
                ATC_TRACE(1 + L, 0);
                ptx::mbar_wait(qk_full, ph);
                ATC_TRACE(1 + L, 1);
                ptx::mbar_wait(&s_free[L], par ^ 1);  // this lane's previous O (aliasing S) has been read out
                ATC_TRACE(1 + L, 2);
                ptx::tcgen05_fence_after();
                const uint64_t dk = ptx::make_kmajor_sw128_desc(sbase + OFF_K);
                const uint64_t dq = ptx::make_kmajor_sw128_desc(sbase + OFF_Q + t * Q_BYTES);
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) ptx::umma_f16<1>(tmem_d, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
                ptx::umma_commit<1>(&s_full[L]);
                ptx::umma_commit<1>(qk_empty);
                ATC_TRACE(1 + L, 3);
                ptx::mbar_wait(v_full, ph);
                ATC_TRACE(1 + L, 4);
                ptx::mbar_wait(&p_full[L], par);  // P is in shared memory, S has been consumed
                ATC_TRACE(1 + L, 5);
                ptx::tcgen05_fence_after();
                for (int j = 0; j < ksteps; ++j) {
                    // A: P k-step j = two [128 rows][8 keys] chunks 2048 B apart, 8-row groups 128 B apart
                    const uint64_t da = ptx::make_smem_desc(sp + j * 4096, 2048, 128, 0);
                    // B: V rows 16j..16j+15 (two 8-row 1024-byte swizzle groups), 64 contiguous channels per row
                    const uint64_t db = ptx::make_smem_desc(sbase + OFF_V + j * 2048, 0, 1024, 2);
                    ptx::umma_f16<1>(tmem_d, da, db, idesc_o, j != 0);
                }
                ptx::umma_commit<1>(&o_full[L]);
                ptx::umma_commit<1>(v_empty);
                ATC_TRACE(1 + L, 6);
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ===================== softmax + output group of lane L =====================
        const int L = (warp - 4) >> 3;
        const int half = ((warp - 4) >> 2) & 1;  // which half of the keys / of the 64 output channels
        const int quarter = warp & 3;            // TMEM lanes this warp may touch: 32 * (warp % 4) ...
        const uint32_t pair_bar = 1 + L * 4 + quarter;  // named barrier shared with the warp handling the other half
        const uint32_t group_bar = 9 + L;               // named barrier of the lane's 8 softmax warps
        const bool poller = ((warp - 4) & 7) == 0;
        // A warp spinning on an mbarrier issues a few instructions every ~20 clocks; with 8 warps of one lane doing that
        // while the other lane computes, a third of the SM's issue slots went to polling.  One warp per lane polls, the
        // other seven sleep in a hardware named barrier.
        auto group_wait = [&](uint64_t* bar, uint32_t parity) {
            if (poller) ptx::mbar_wait(bar, parity);
            ptx::named_bar_sync(group_bar, 256);
        };
        const int row_in_tile = quarter * 32 + lane;
        const uint32_t stage = sbase + OFF_OUT + (L * 4 + quarter) * 4096;
        const void* stage_ptr = smem + OFF_OUT + (L * 4 + quarter) * 4096;
        float2* xch_mine = reinterpret_cast<float2*>(smem + OFF_XCH) + (L * 2 + half) * 128 + row_in_tile;
        float2* xch_other = reinterpret_cast<float2*>(smem + OFF_XCH) + (L * 2 + (half ^ 1)) * 128 + row_in_tile;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + L * 256;
        const uint32_t prow = sbase + OFF_P + L * P_BYTES + row_in_tile * 16;
        // this thread's key columns [cb, ce): the first ceil(nchunks/2) 16-column chunks go to half 0
        const int split = ((nk16 / 16 + 1) / 2) * 16;
        const int cb = half ? split : 0, ce = half ? nk16 : split;
        int it = 0, n = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const int t = (n_tiles == 2) ? ((L + it) & 1) : L;  // one tile: lane 0 does every item, lane 1 idles
            if (t >= n_tiles) continue;
            const uint32_t par = n & 1;
            ++n;
            const int b = item / H, h = item - b * H;
            const bool warp_valid = (t * QT + quarter * 32) < N;
            // keys this query row may see: all N, or 0..row under a causal mask (padding rows see nothing that is kept)
            const int NV = causal ? min(N, t * QT + row_in_tile + 1) : N;
            float sum0 = 0.f, sum1 = 0.f;

#define ATC_TRACE_S(ev)                                     \
    do {                                                    \
        if (quarter == 0 && half == 0 && lane == 0) ATC_TRACE(3 + L, ev); \
    } while (0)
            ATC_TRACE_S(0);
            group_wait(&s_full[L], par);
            ATC_TRACE_S(1);
            ptx::tcgen05_fence_after();
            float mx = -INFINITY;
            if (warp_valid) {
                for (int c0 = cb; c0 < ce; c0 += 32) {
                    uint32_t r[32];
                    if (c0 + 32 <= ce) {
                        ptx::tmem_ld_32x32b_x32(taddr + c0, r);
                        ptx::tmem_ld_wait();
                        mx = chunk_max<32>(r, c0, NV, mx);
                    } else {
                        ptx::tmem_ld_32x32b_x16(taddr + c0, *reinterpret_cast<uint32_t(*)[16]>(&r[0]));
                        ptx::tmem_ld_wait();
                        mx = chunk_max<16>(r, c0, NV, mx);
                    }
                }
                xch_mine->x = mx;
                // the staging tile is about to be reused: its previous TMA store (issued by the half-0 warp) must have read it
                if (half == 0 && lane == 0) ptx::bulk_wait_group_read<0>();
            }
            ATC_TRACE_S(2);
            ptx::named_bar_sync(pair_bar, 64);
            // Pass the exponential phase back and forth between the lanes, so that one lane's softmax runs under the other
            // lane's MMAs / barriers / output instead of both lanes doing the same phase at the same time (they fall into
            // lockstep otherwise, because they share the Q/K/V buffers).  Measured: 162 -> 136 us per layer.
            if (n_tiles == 2) group_wait(&tok[L ^ 1], L == 0 ? (par ^ 1) : par);
            ATC_TRACE_S(3);
            if (warp_valid) {
                mx = fmaxf(mx, xch_other->x);  // half 0 always holds key 0, so the row maximum is finite
                const float neg_mxs = -mx * scale_log2e;
                bool handed = false;
                for (int c0 = cb; c0 < ce; c0 += 32) {
                    uint32_t r[32];
                    // hand the exponential phase to the other lane while this warp still has its last chunk to do (measured:
                    // 136 -> 132 us per layer; handing over one chunk earlier, 141 us): its
                    // pack / store tail and the other lane's first loads then overlap instead of leaving the MUFU idle
                    if (c0 + 32 >= ce) {
                        if (lane == 0) ptx::mbar_arrive(&tok[L]);
                        handed = true;
                    }
                    if (c0 + 32 <= ce) {
                        ptx::tmem_ld_32x32b_x32(taddr + c0, r);
                        ptx::tmem_ld_wait();
                        chunk_exp<T, 32>(r, c0, NV, scale_log2e, neg_mxs, prow, sum0, sum1);
                    } else {
                        ptx::tmem_ld_32x32b_x16(taddr + c0, *reinterpret_cast<uint32_t(*)[16]>(&r[0]));
                        ptx::tmem_ld_wait();
                        chunk_exp<T, 16>(r, c0, NV, scale_log2e, neg_mxs, prow, sum0, sum1);
                    }
                }
                if (!handed && lane == 0) ptx::mbar_arrive(&tok[L]);  // this half has no key columns at all (tiny N)
                xch_mine->y = sum0 + sum1;
            } else if (lane == 0) {
                ptx::mbar_arrive(&tok[L]);  // rows past the sequence: nothing to do, pass the turn on
            }
            ptx::fence_proxy_async_smem();  // P (generic-proxy stores) before the MMA's async-proxy reads
            ptx::tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&p_full[L]);
            ATC_TRACE_S(4);
            group_wait(&o_full[L], par);
            ATC_TRACE_S(5);
            ptx::tcgen05_fence_after();
            if (warp_valid) {
                uint32_t r[32];
                ptx::tmem_ld_32x32b_x32(taddr + half * 32, r);
                const float inv_sum = 1.0f / (sum0 + sum1 + xch_other->y);
                ptx::tmem_ld_wait();
                const uint32_t srow = stage + lane * 128;
                const uint32_t swz = static_cast<uint32_t>(lane & 7);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t u[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float a0, a1;
                        ptx::mul2(a0, a1, __uint_as_float(r[8 * c + 2 * i]), __uint_as_float(r[8 * c + 2 * i + 1]), inv_sum);
                        u[i] = pack2<T>(a0, a1);
                    }
                    ptx::st_shared_v4(srow + ((static_cast<uint32_t>(4 * half + c) ^ swz) << 4), u[0], u[1], u[2], u[3]);
                }
                ptx::fence_proxy_async_smem();
            }
            ptx::tcgen05_fence_before();
            ATC_TRACE_S(6);
            ptx::named_bar_sync(pair_bar, 64);  // both halves of the staging tile written, both warps done with TMEM
            ATC_TRACE_S(7);
            if (half == 0 && lane == 0) {
                ptx::mbar_arrive(&s_free[L]);  // the lane's TMEM columns may take the next S
                if (warp_valid) {
                    ptx::tma_store_3d(&map_out, stage_ptr, h * HD, t * QT + quarter * 32, b);
                    ptx::bulk_commit_group();
                }
            }
        }
        if (half == 0 && lane == 0) ptx::bulk_wait_group<0>();
    }

    ptx::tcgen05_fence_before();
    __syncthreads();
    if (warp == 3) ptx::tmem_dealloc<1>(tmem_base, 512);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

long long* g_trace = nullptr;  // developer hook, see attention_set_trace

template <typename T>
int launch_tc(const AttentionMaps& m, int B, int N, int H, float scale_log2e, cudaStream_t stream) {
    auto kern = attention_tc_kernel<T>;
    static bool configured = false;
    if (!configured) {
        VIDIL_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ATC_SMEM));
        configured = true;
    }
    const int n_items = B * H;
    int grid = gemm_num_sms();
    if (grid > n_items) grid = n_items;
    if (grid < 1) return 1;
    VIDIL_CUDA_OK(launch_pdl(kern, dim3(grid), dim3(ATC_THREADS), ATC_SMEM, stream, m.q, m.kv, m.out, n_items, N, H, scale_log2e, g_trace,
                             m.causal ? 1 : 0));
    count_launches(1);
    return 0;
}

}  // namespace

void attention_set_trace(long long* dev_buf) { g_trace = dev_buf; }

bool attention_tc_supported(int N) { return N >= 1 && N <= MAX_KEYS; }

int attention_tc_prepare(AttentionMaps& m, const void* qkv, void* out, DType dt, int B, int N, int H) {
    EncodeTiledFn fn = encode_fn();
    if (fn == nullptr) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return 1;
    }
    if ((reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) {
        set_error("attention: qkv and out must be 16-byte aligned");
        return 1;
    }
    const CUtensorMapDataType cdt = (dt == DT_BF16) ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    const int nk16 = (N + 15) & ~15;
    const cuuint32_t estr3[3] = {1, 1, 1};
    {
        const cuuint64_t dims[2] = {static_cast<cuuint64_t>(3) * H * HD, static_cast<cuuint64_t>(B) * N};
        const cuuint64_t strides[1] = {static_cast<cuuint64_t>(3) * H * HD * 2};
        const cuuint32_t box_q[2] = {HD, QT};
        const cuuint32_t box_kv[2] = {HD, static_cast<cuuint32_t>(nk16)};
        CUresult r = fn(&m.q, cdt, 2, const_cast<void*>(qkv), dims, strides, box_q, estr3, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r == CUDA_SUCCESS)
            r = fn(&m.kv, cdt, 2, const_cast<void*>(qkv), dims, strides, box_kv, estr3, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("attention: cuTensorMapEncodeTiled(qkv) failed with CUresult %d", static_cast<int>(r));
            return 1;
        }
    }
    {
        const cuuint64_t dims[3] = {static_cast<cuuint64_t>(H) * HD, static_cast<cuuint64_t>(N), static_cast<cuuint64_t>(B)};
        const cuuint64_t strides[2] = {static_cast<cuuint64_t>(H) * HD * 2, static_cast<cuuint64_t>(N) * H * HD * 2};
        const cuuint32_t box[3] = {HD, 32, 1};
        CUresult r = fn(&m.out, cdt, 3, out, dims, strides, box, estr3, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("attention: cuTensorMapEncodeTiled(out) failed with CUresult %d", static_cast<int>(r));
            return 1;
        }
    }
    m.qkv = qkv;
    m.out_ptr = out;
    m.B = B;
    m.N = N;
    m.H = H;
    m.dt = dt;
    return 0;
}

int attention_tc_run(const AttentionMaps& m, float scale, cudaStream_t stream) {
    if (gemm_num_sms() == 0) return 1;
    const float sl2 = scale * 1.4426950408889634f;
    if (m.dt == DT_BF16) return launch_tc<__nv_bfloat16>(m, m.B, m.N, m.H, sl2, stream);
    return launch_tc<__half>(m, m.B, m.N, m.H, sl2, stream);
}

}  // namespace vidil
