// Fused softmax attention on tcgen05 for CLIP ViT-L/14's 257 tokens (head_dim 64) — the same single-pass design as
// attention_tc.cu (S and P never leave tensor memory, one thread per query row, P written back with tcgen05.st and fed to
// the PV MMA as a TMEM A operand), stretched over the one token that does not fit:
//   queries  rows 0..255 are two 128-row tcgen05 tiles (one per lane); row 256 is computed by four SIMT "tail" warps from the
//            K / V tiles that are in shared memory anyway (split over the keys, merged through shared memory)
//   keys     0..255 are scored by the MMA (S = Q K^T, N = 256: exactly the lane's 256 TMEM columns); key 256 is scored by
//            the softmax threads themselves (one 64-element dot product per row, while the S MMA runs) and enters the PV
//            product as a 17th k-step whose other 15 keys carry P = 0
// Reference: the attention transformers' CLIPAttention.forward performs for run_visual_tokenization.py:138-142
// (modeling_clip.py:300-336, eager_attention_forward :261-279), scale 1/8, no mask.
// TMEM columns of a lane: S [0,256) -> P of keys [0,128) in [0,64), O in [64,128), P of keys [128,256) in [128,192), P of
// keys 256..271 in [192,200).  The first eight PV k-steps are issued while the second half of the row is still being
// exponentiated (p_half), as in the N = 197 kernel.
// The previous kernel for this shape (attention_tcl.cu: key-block loop, S computed twice, P through shared memory) stays for
// the other long sequences (ViT-B/16 @384: 577 tokens).
#include <math.h>
#include <stdlib.h>

#include <type_traits>

#include "attention_common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace vidil {
namespace {

using namespace attn;

constexpr int HD = 64;
constexpr int QT = 128;
constexpr int NTOK = 257;
constexpr int NKM = 256;            // keys scored by the MMA
constexpr int KV_ROWS = 272;        // K / V rows staged per item: 256 + one 16-row box holding token 256
constexpr int T257_THREADS = 512;   // warp 0 producer, 1-2 MMA issuers, 3 TMEM allocator, 4-7 / 8-11 softmax, 12-15 tail row
constexpr int POLY = 6;             // pairs of every 16 exponentiated on the FMA pipe (see attention_common.cuh)

constexpr int Q_BYTES = QT * HD * 2;
constexpr int KV_BYTES = KV_ROWS * HD * 2;
// Q (two tiles) and K are double-buffered — the TMA producer runs one item ahead, so their HBM latency is hidden behind the
// current item; V (needed only from the middle of the softmax phase on) keeps one buffer: two of everything would be 232 KB
constexpr int STAGES = 2;
constexpr int OFF_Q = 0;
constexpr int OFF_K = STAGES * 2 * Q_BYTES;
constexpr int OFF_V = OFF_K + STAGES * KV_BYTES;
constexpr int OFF_OUT = OFF_V + KV_BYTES;
constexpr int OFF_TAIL = OFF_OUT + 8 * 4096;      // tail row: q [64] fp32, then per-warp partials [4][66] fp32
constexpr int OFF_BAR = OFF_TAIL + 1536;
constexpr int T257_SMEM = OFF_BAR + 256 + 1024;
static_assert(OFF_K % 1024 == 0 && OFF_V % 1024 == 0 && OFF_OUT % 1024 == 0, "tile alignment");
static_assert(T257_SMEM <= 227 * 1024, "shared memory budget");

constexpr int O_COL = 64, PHI_COL = 128, PX_COL = 192;
__host__ __device__ constexpr int pcol(int chunk) { return chunk < 4 ? 16 * chunk : PHI_COL + 16 * (chunk - 4); }

template <typename T>
__device__ __forceinline__ void unpack2(uint32_t w, float& lo, float& hi) {
    if constexpr (std::is_same<T, __nv_bfloat16>::value) {
        lo = __uint_as_float(w << 16);
        hi = __uint_as_float(w & 0xffff0000u);
    } else {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w));
        lo = f.x;
        hi = f.y;
    }
}

template <typename T>
__global__ void __launch_bounds__(T257_THREADS, 1)
    attention_tc257_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv,
                           const __grid_constant__ CUtensorMap map_kv16, const __grid_constant__ CUtensorMap map_out, int n_items,
                           int H, float scale_log2e, const T* __restrict__ qkv, T* __restrict__ out, long long* trace) {
#define T257_TRACE(role, ev)                                                                                  \
    do {                                                                                                      \
        if (trace != nullptr && blockIdx.x == 0 && it < 16) trace[((role) * 16 + it) * 8 + (ev)] = clock64(); \
    } while (0)
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    const uint32_t sbase = ptx::smem_u32(smem);

    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* qk_full = bars + 0;   // [2] per stage
    uint64_t* qk_empty = bars + 2;  // [2]
    uint64_t* v_full = bars + 4;
    uint64_t* v_empty = bars + 5;
    uint64_t* s_full = bars + 6;    // [2] per lane
    uint64_t* p_full = bars + 8;    // [2]
    uint64_t* o_full = bars + 10;   // [2]
    uint64_t* s_free = bars + 12;   // [2]
    uint64_t* tok = bars + 14;      // [2]
    uint64_t* p_half = bars + 16;   // [2]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 18);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int D = H * HD;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&map_q);
        ptx::prefetch_tensormap(&map_kv);
        ptx::prefetch_tensormap(&map_kv16);
        ptx::prefetch_tensormap(&map_out);
        for (int st = 0; st < STAGES; ++st) {
            ptx::mbar_init(&qk_full[st], 1);
            ptx::mbar_init(&qk_empty[st], 2 + 8 + 1);  // both S MMAs, the 8 softmax warps (key 256 scored from Q and K), the tail group
        }
        ptx::mbar_init(v_full, 1);
        ptx::mbar_init(v_empty, 2 + 1);       // both PV MMAs, the tail group
        for (int l = 0; l < 2; ++l) {
            ptx::mbar_init(&s_full[l], 1);
            ptx::mbar_init(&p_full[l], 4);
            ptx::mbar_init(&o_full[l], 1);
            ptx::mbar_init(&s_free[l], 4);
            ptx::mbar_init(&tok[l], 4);
            ptx::mbar_init(&p_half[l], 4);
        }
        ptx::fence_mbar_init();
    }
    if (warp == 3) ptx::tmem_alloc<1>(tmem_ptr_smem, 512);
    ptx::tcgen05_fence_before();
    __syncthreads();
    ptx::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    ptx::griddep_launch();
    ptx::griddep_wait();

    if (warp == 0) {
        if (lane == 0) {
            // ===================== TMA producer =====================
            int it = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const int st = it & 1;
                const uint32_t ph = it & 1, ph2 = (it >> 1) & 1;
                const int ritem = n_items - 1 - item;   // last frames first, as in attention_tc.cu: the end of the qkv matrix is what L2 still holds
                const int b = ritem / H, h = ritem - b * H;
                const int row0 = b * NTOK;
                uint8_t* sq = smem + OFF_Q + st * 2 * Q_BYTES;
                uint8_t* skk = smem + OFF_K + st * KV_BYTES;
                T257_TRACE(0, 0);
                ptx::mbar_wait(&qk_empty[st], ph2 ^ 1);
                T257_TRACE(0, 1);
                ptx::mbar_arrive_expect_tx(&qk_full[st], 2 * Q_BYTES + KV_BYTES);
                ptx::tma_load_2d(&map_q, &qk_full[st], sq, h * HD, row0);
                ptx::tma_load_2d(&map_q, &qk_full[st], sq + Q_BYTES, h * HD, row0 + QT);
                ptx::tma_load_2d(&map_kv, &qk_full[st], skk, D + h * HD, row0);
                ptx::tma_load_2d(&map_kv16, &qk_full[st], skk + NKM * HD * 2, D + h * HD, row0 + NKM);
                ptx::mbar_wait(v_empty, ph ^ 1);
                T257_TRACE(0, 2);
                ptx::mbar_arrive_expect_tx(v_full, KV_BYTES);
                ptx::tma_load_2d(&map_kv, v_full, smem + OFF_V, 2 * D + h * HD, row0);
                ptx::tma_load_2d(&map_kv16, v_full, smem + OFF_V + NKM * HD * 2, 2 * D + h * HD, row0 + NKM);
            }
        }
        __syncwarp();
    } else if (warp == 1 || warp == 2) {
        if (lane == 0) {
            // ===================== MMA issuer of lane L =====================
            const int L = warp - 1;
            constexpr bool kIsBf16 = std::is_same<T, __nv_bfloat16>::value;
            constexpr uint32_t idesc_s = ptx::make_idesc_f16(kIsBf16, QT, NKM);
            constexpr uint32_t idesc_o = ptx::make_idesc_f16_bmn(kIsBf16, QT, HD);
            const uint32_t tmem_s = tmem_base + L * 256;
            const uint32_t tmem_o = tmem_s + O_COL;
            int it = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const int t = (L + it) & 1, st = it & 1;
                const uint32_t ph = it & 1, ph2 = (it >> 1) & 1;
                T257_TRACE(1 + L, 0);
                ptx::mbar_wait(&qk_full[st], ph2);
                T257_TRACE(1 + L, 1);
                ptx::mbar_wait(&s_free[L], ph ^ 1);
                T257_TRACE(1 + L, 2);
                ptx::tcgen05_fence_after();
                const uint64_t dk = ptx::make_kmajor_sw128_desc(sbase + OFF_K + st * KV_BYTES);
                const uint64_t dq = ptx::make_kmajor_sw128_desc(sbase + OFF_Q + (st * 2 + t) * Q_BYTES);
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) ptx::umma_f16<1>(tmem_s, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
                ptx::umma_commit<1>(&s_full[L]);
                ptx::umma_commit<1>(&qk_empty[st]);
                T257_TRACE(1 + L, 3);
                ptx::mbar_wait(v_full, ph);
                T257_TRACE(1 + L, 4);
                ptx::mbar_wait(&p_half[L], ph);
                ptx::tcgen05_fence_after();
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint64_t db = ptx::make_smem_desc(sbase + OFF_V + j * 2048, 0, 1024, 2);
                    ptx::umma_f16_tmem_a(tmem_o, tmem_s + j * 8, db, idesc_o, j != 0);
                }
                ptx::mbar_wait(&p_full[L], ph);
                T257_TRACE(1 + L, 5);
                ptx::tcgen05_fence_after();
#pragma unroll
                for (int j = 8; j < 17; ++j) {
                    const uint64_t db = ptx::make_smem_desc(sbase + OFF_V + j * 2048, 0, 1024, 2);
                    ptx::umma_f16_tmem_a(tmem_o, tmem_s + (j < 16 ? PHI_COL + (j - 8) * 8 : PX_COL), db, idesc_o, 1);
                }
                ptx::umma_commit<1>(&o_full[L]);
                ptx::umma_commit<1>(v_empty);
                T257_TRACE(1 + L, 6);
            }
        }
        __syncwarp();
    } else if (warp >= 4 && warp < 12) {
        // ===================== softmax + output warps of lane L =====================
        const int L = (warp - 4) >> 2;
        const int quarter = warp & 3;
        const uint32_t group_bar = 1 + L;
        const bool poller = quarter == 0;
        auto group_wait = [&](uint64_t* bar, uint32_t parity) {
            if (poller) ptx::mbar_wait(bar, parity);
            ptx::named_bar_sync(group_bar, 128);
        };
        const int row_in_tile = quarter * 32 + lane;
        const uint32_t stage = sbase + OFF_OUT + (L * 4 + quarter) * 4096;
        const void* stage_ptr = smem + OFF_OUT + (L * 4 + quarter) * 4096;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + L * 256;
        int it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const int t = (L + it) & 1, st = it & 1;
            const uint32_t ph = it & 1, ph2 = (it >> 1) & 1;
            const int ritem = n_items - 1 - item;
            const int b = ritem / H, h = ritem - b * H;
            float sum[4] = {0.f, 0.f, 0.f, 0.f};

            // ---- key 256: this row's score from the Q and K tiles in shared memory, under the S MMA ----
#define T257_TRACE_S(ev)                                     \
    do {                                                     \
        if (quarter == 0 && lane == 0) T257_TRACE(3 + L, ev); \
    } while (0)
            T257_TRACE_S(0);
            group_wait(&qk_full[st], ph2);
            float sx = 0.f;
            {
                const uint32_t qrow = sbase + OFF_Q + (st * 2 + t) * Q_BYTES + row_in_tile * 128;
                const uint32_t krow = sbase + OFF_K + st * KV_BYTES + NKM * 128;  // row 256: 256 & 7 == 0, so its 16-byte chunks are in order
                float s0 = 0.f, s1 = 0.f;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    uint32_t qw[4], kw[4];
                    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                                 : "=r"(qw[0]), "=r"(qw[1]), "=r"(qw[2]), "=r"(qw[3])
                                 : "r"(qrow + ((c ^ (row_in_tile & 7)) << 4)));
                    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                                 : "=r"(kw[0]), "=r"(kw[1]), "=r"(kw[2]), "=r"(kw[3])
                                 : "r"(krow + (c << 4)));
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float ql, qh, kl, kh;
                        unpack2<T>(qw[i], ql, qh);
                        unpack2<T>(kw[i], kl, kh);
                        s0 = fmaf(ql, kl, s0);
                        s1 = fmaf(qh, kh, s1);
                    }
                }
                sx = s0 + s1;
            }
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&qk_empty[st]);

            T257_TRACE_S(1);
            group_wait(&s_full[L], ph);
            T257_TRACE_S(2);
            ptx::tcgen05_fence_after();
            // ---- pass 1: row maximum over the 256 MMA scores and key 256 ----
            float mx[4] = {sx, -INFINITY, -INFINITY, -INFINITY};
            {
                uint32_t ra[32], rb[32];
#pragma unroll
                for (int c = 0; c < 8; c += 2) {
                    ptx::tmem_ld_32x32b_x32(taddr + 32 * c, ra);
                    ptx::tmem_ld_32x32b_x32(taddr + 32 * c + 32, rb);
                    ptx::tmem_ld_wait();
                    chunk_max<32>(ra, 32 * c, NKM, mx);
                    chunk_max<32>(rb, 32 * c + 32, NKM, mx);
                }
            }
            const float neg_mxs = -fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * scale_log2e;
            // the exponential phase alternates between the lanes (see attention_tc.cu)
            T257_TRACE_S(3);
            group_wait(&tok[L ^ 1], L == 0 ? (ph ^ 1) : ph);
            T257_TRACE_S(4);
            // ---- pass 2: exponentials, row sum, P -> TMEM ----
            {
                uint32_t ra[32], rb[32];
                ptx::tmem_ld_32x32b_x32(taddr, ra);
#pragma unroll
                for (int c = 0; c < 8; c += 2) {
                    ptx::tmem_ld_wait();
                    ptx::tmem_ld_32x32b_x32(taddr + 32 * (c + 1), rb);
                    chunk_exp_impl<T, 32, false, POLY>(ra, 32 * c, NKM, scale_log2e, neg_mxs, taddr + pcol(c) - 16 * c, sum);
                    if (c == 4) {
                        // keys [0,128) = chunks 0..3 were stored a chunk ago: the wait is short, and the MMA issuer may start
                        // the first eight PV k-steps while this warp exponentiates the rest of the row
                        ptx::tmem_st_wait();
                        ptx::tcgen05_fence_before();
                        __syncwarp();
                        if (lane == 0) ptx::mbar_arrive(&p_half[L]);
                    }
                    ptx::tmem_ld_wait();
                    if (c + 2 < 8) ptx::tmem_ld_32x32b_x32(taddr + 32 * (c + 2), ra);
                    if (c + 1 == 7 && lane == 0) ptx::mbar_arrive(&tok[L]);  // entering the last chunk: hand the MUFU phase over
                    chunk_exp_impl<T, 32, false, POLY>(rb, 32 * (c + 1), NKM, scale_log2e, neg_mxs, taddr + pcol(c + 1) - 16 * (c + 1), sum);
                }
                // key 256 (and 15 keys of padding with P = 0): one more 16-key k-step
                const float px = ptx::ex2_approx(fmaf(sx, scale_log2e, neg_mxs));
                sum[0] += px;
                uint32_t pk8[8] = {pack2<T>(px, 0.f), 0u, 0u, 0u, 0u, 0u, 0u, 0u};
                ptx::tmem_st_32x32b_x8(taddr + PX_COL, pk8);
                ptx::tmem_st_wait();
            }
            ptx::tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&p_full[L]);
            T257_TRACE_S(5);

            group_wait(&o_full[L], ph);
            T257_TRACE_S(6);
            ptx::tcgen05_fence_after();
            const uint32_t srow = stage + lane * 128;
            const uint32_t swz = static_cast<uint32_t>(lane & 7);
            const float inv_sum = 1.0f / ((sum[0] + sum[1]) + (sum[2] + sum[3]));
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t o[32];
                ptx::tmem_ld_32x32b_x32(taddr + O_COL + 32 * half, o);
                if (half == 0 && lane == 0) ptx::bulk_wait_group_read<0>();  // the staging tile's previous store has read it
                ptx::tmem_ld_wait();
                if (half == 0) __syncwarp();
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t u[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float a0, a1;
                        ptx::mul2(a0, a1, __uint_as_float(o[8 * c + 2 * i]), __uint_as_float(o[8 * c + 2 * i + 1]), inv_sum);
                        u[i] = pack2<T>(a0, a1);
                    }
                    ptx::st_shared_v4(srow + ((static_cast<uint32_t>(4 * half + c) ^ swz) << 4), u[0], u[1], u[2], u[3]);
                }
            }
            ptx::tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&s_free[L]);
            T257_TRACE_S(7);
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                ptx::tma_store_3d(&map_out, stage_ptr, h * HD, t * QT + quarter * 32, b);
                ptx::bulk_commit_group();
            }
        }
        if (lane == 0) ptx::bulk_wait_group<0>();
    } else if (warp >= 12) {
        // ===================== query row 256 of each (frame, head): four SIMT warps, a quarter of the keys each =====================
        const int tw = warp - 12;
        float* tq = reinterpret_cast<float*>(smem + OFF_TAIL);   // [64] the query row, pre-scaled by scale * log2 e
        float* tpart = tq + HD;                                  // [4 warps][66]: partial output, running max, running sum
        constexpr int PER_WARP = (NTOK + 3) / 4;                 // 65 keys per warp
        int it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const int st = it & 1;
            const uint32_t ph = it & 1, ph2 = (it >> 1) & 1;
            const int ritem = n_items - 1 - item;
            const int b = ritem / H, h = ritem - b * H;
            if (threadIdx.x - 12 * 32 < HD)
                tq[threadIdx.x - 12 * 32] =
                    static_cast<float>(qkv[(static_cast<int64_t>(b) * NTOK + NKM) * (3 * D) + h * HD + (threadIdx.x - 12 * 32)]) * scale_log2e;
            // ---- phase A (K only): this warp's <= 65 scores, three per lane, then the warp's own maximum / exponentials / sum;
            //      K is released at once — the next item's Q / K load must not wait for the V phase ----
            if (tw == 0 && lane == 0) T257_TRACE(5, 0);
            if (tw == 0) ptx::mbar_wait(&qk_full[st], ph2);
            ptx::named_bar_sync(11, 128);
            if (tw == 0 && lane == 0) T257_TRACE(5, 1);
            const int k_begin = tw * PER_WARP, k_end = min(NTOK, k_begin + PER_WARP);
            const uint32_t sk = sbase + OFF_K + st * KV_BYTES, sv = sbase + OFF_V;
            const float4* q4 = reinterpret_cast<const float4*>(tq);
            float sc[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const int k = k_begin + 32 * r + lane;
                sc[r] = -INFINITY;
                if (k < k_end) {
                    float s0 = 0.f, s1 = 0.f;
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        uint32_t w[4];
                        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                                     : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3])
                                     : "r"(sk + k * 128 + ((c ^ (k & 7)) << 4)));
                        const float4 qa = q4[2 * c], qb = q4[2 * c + 1];
                        const float qq[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            float lo, hi;
                            unpack2<T>(w[i], lo, hi);
                            s0 = fmaf(qq[2 * i], lo, s0);
                            s1 = fmaf(qq[2 * i + 1], hi, s1);
                        }
                    }
                    sc[r] = s0 + s1;
                }
            }
            ptx::named_bar_sync(11, 128);  // all four warps have read their K rows
            if (tw == 0 && lane == 0) ptx::mbar_arrive(&qk_empty[st]);
            if (tw == 0 && lane == 0) T257_TRACE(5, 2);
            float m_run = fmaxf(sc[0], fmaxf(sc[1], sc[2]));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m_run = fmaxf(m_run, __shfl_xor_sync(0xffffffffu, m_run, o));
            float pr[3], l_run = 0.f;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                pr[r] = (k_begin + 32 * r + lane < k_end) ? ptx::ex2_approx(sc[r] - m_run) : 0.f;
                l_run += pr[r];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) l_run += __shfl_xor_sync(0xffffffffu, l_run, o);
            // ---- phase B (V): out += p v over this warp's keys.  Lane = (key group g = lane >> 3, channel octet c8 = lane & 7):
            //      one 16-byte V chunk per key, keys g, g + 4, ... of the warp's range, then a sum over the 4 key groups ----
            if (tw == 0) ptx::mbar_wait(v_full, ph);
            ptx::named_bar_sync(11, 128);
            if (tw == 0 && lane == 0) T257_TRACE(5, 3);
            const int g = lane >> 3, c8 = lane & 7;
            float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int r = 0; r < 3; ++r) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int kk = 4 * j + g;                         // key offset within this block of 32
                    const float pk = __shfl_sync(0xffffffffu, pr[r], kk);
                    const int key = k_begin + 32 * r + kk;
                    if (key < k_end) {
                        uint32_t w[4];
                        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                                     : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3])
                                     : "r"(sv + key * 128 + ((c8 ^ (key & 7)) << 4)));
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            float lo, hi;
                            unpack2<T>(w[i], lo, hi);
                            acc[2 * i] = fmaf(pk, lo, acc[2 * i]);
                            acc[2 * i + 1] = fmaf(pk, hi, acc[2 * i + 1]);
                        }
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8);
                acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
            }
            float* pp = tpart + tw * 66;
            if (lane < 8) {
#pragma unroll
                for (int i = 0; i < 8; ++i) pp[8 * c8 + i] = acc[i];
            }
            if (lane == 0) {
                pp[64] = m_run;
                pp[65] = l_run;
            }
            ptx::named_bar_sync(11, 128);  // all four warps are done with this item's V tile; partials are visible
            if (tw == 0) {
                if (lane == 0) ptx::mbar_arrive(v_empty);
                if (lane == 0) T257_TRACE(5, 4);
                float M = -INFINITY;
#pragma unroll
                for (int w = 0; w < 4; ++w) M = fmaxf(M, tpart[w * 66 + 64]);
                float den = 0.f, o0 = 0.f, o1 = 0.f;
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const float* p4 = tpart + w * 66;
                    const float f = ptx::ex2_approx(p4[64] - M);
                    den = fmaf(f, p4[65], den);
                    o0 = fmaf(f, p4[2 * lane], o0);
                    o1 = fmaf(f, p4[2 * lane + 1], o1);
                }
                const float inv = 1.0f / den;
                *reinterpret_cast<uint32_t*>(out + (static_cast<int64_t>(b) * NTOK + NKM) * D + h * HD + 2 * lane) =
                    pack2<T>(o0 * inv, o1 * inv);
            }
            ptx::named_bar_sync(11, 128);  // tq / tpart are reused by the next item
        }
    }

    ptx::tcgen05_fence_before();
    __syncthreads();
    if (warp == 3) ptx::tmem_dealloc<1>(tmem_base, 512);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn257() {
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

long long* g_trace257 = nullptr;

template <typename T>
int launch257(const AttentionMaps& m, float scale_log2e, cudaStream_t stream) {
    auto kern = attention_tc257_kernel<T>;
    // per launch: the attribute belongs to the current device's copy of the function
    VIDIL_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, T257_SMEM));
    const int n_items = m.B * m.H;
    int grid = gemm_num_sms();
    if (grid > n_items) grid = n_items;
    if (grid < 1) return 1;
    VIDIL_CUDA_OK(launch_pdl(kern, dim3(grid), dim3(T257_THREADS), T257_SMEM, stream, m.q, m.kv, m.kv16, m.out, n_items, m.H,
                             scale_log2e, reinterpret_cast<const T*>(m.qkv), reinterpret_cast<T*>(m.out_ptr), g_trace257));
    count_launches(1);
    return 0;
}

}  // namespace

void attention_tc257_set_trace(long long* dev_buf) { g_trace257 = dev_buf; }

bool attention_tc257_supported(int N) {
    static const bool off = [] { const char* e = getenv("VIDIL_ATTN257"); return e != nullptr && e[0] == '0'; }();  // developer A/B
    return N == NTOK && !off;
}

int attention_tc257_prepare(AttentionMaps& m, const void* qkv, void* out, DType dt, int B, int N, int H) {
    EncodeTiledFn fn = encode_fn257();
    if (fn == nullptr) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return 1;
    }
    if (N != NTOK) {
        set_error("attention_tc257: built for %d tokens, got %d", NTOK, N);
        return 1;
    }
    if ((reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) {
        set_error("attention: qkv and out must be 16-byte aligned");
        return 1;
    }
    const CUtensorMapDataType cdt = (dt == DT_BF16) ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    const cuuint32_t estr3[3] = {1, 1, 1};
    {
        const cuuint64_t dims[2] = {static_cast<cuuint64_t>(3) * H * HD, static_cast<cuuint64_t>(B) * N};
        const cuuint64_t strides[1] = {static_cast<cuuint64_t>(3) * H * HD * 2};
        const cuuint32_t box_q[2] = {HD, QT};
        const cuuint32_t box_kv[2] = {HD, NKM};
        const cuuint32_t box_16[2] = {HD, KV_ROWS - NKM};
        CUresult r = fn(&m.q, cdt, 2, const_cast<void*>(qkv), dims, strides, box_q, estr3, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r == CUDA_SUCCESS)
            r = fn(&m.kv, cdt, 2, const_cast<void*>(qkv), dims, strides, box_kv, estr3, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r == CUDA_SUCCESS)
            r = fn(&m.kv16, cdt, 2, const_cast<void*>(qkv), dims, strides, box_16, estr3, CU_TENSOR_MAP_INTERLEAVE_NONE,
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
    m.is_long = false;
    m.is_257 = true;
    return 0;
}

int attention_tc257_run(const AttentionMaps& m, float scale, cudaStream_t stream) {
    if (gemm_num_sms() == 0) return 1;
    const float sl2 = scale * 1.4426950408889634f;
    if (m.dt == DT_BF16) return launch257<__nv_bfloat16>(m, sl2, stream);
    return launch257<__half>(m, sl2, stream);
}

}  // namespace vidil
