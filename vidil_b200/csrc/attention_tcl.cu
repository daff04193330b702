// tcgen05 softmax attention for sequences longer than one key tile (N > 208: CLIP ViT-L/14 has 257 tokens, BLIP
// ViT-B/16 @384 has 577), head_dim 64.  Same building blocks as attention_tc.cu (two lanes per CTA, each with its
// MMA-issuing thread, 256 TMEM columns, P buffer and 8 softmax warps; Q/K/V by TMA in place from the fused-QKV GEMM
// output; V as an MN-major UMMA operand; 3-D TMA store of the head-merged result), plus a loop over key blocks:
//
//   work item   = (frame, head, round r): query tiles 2r and 2r+1 (128 rows each), one per lane
//   key blocks  = nkb blocks of KB <= 144 keys (257 -> 2 x 144, 577 -> 5 x 128)
//   sweep 0     for every key block j:  S = Q K_j^T  -> running row maximum            (no P, no V)
//   sweep 1     for every key block j:  S = Q K_j^T again -> P_j = exp2(c S - c max) -> O += P_j V_j, row sums
//   finally     out = O / rowsum
// Recomputing S costs tensor time that is idle anyway (the kernel is bound by the softmax warps) and removes the
// accumulator rescaling of an online softmax: O lives in its own 64 TMEM columns and is only ever accumulated into.
// A lane whose tile index runs past the last tile recomputes the last tile and does not store it, so both lanes always
// take part in every barrier.  Query rows beyond the last full tile are not worth a whole extra round when there are only
// a few of them (the 257th token of CLIP): up to 4 such rows are computed by four extra "tail" warps of the same CTA with
// plain fp32 SIMT arithmetic straight from the K / V tiles that are in shared memory anyway (no extra HBM traffic);
// 5..16 rows go to attention_rows_kernel below.
#include <math.h>

#include <type_traits>

#include "kernels.h"
#include "ptx.cuh"

namespace vidil {
namespace {

constexpr int HD = 64;
constexpr int QT = 128;
constexpr int MAX_KB = 144;        // keys per block: S (fp32) in TMEM columns [0, 144), O in [192, 256) of the lane's 256
constexpr int O_COL = 192;
constexpr int ATL_THREADS = 768;   // warp 0 producer, 1-2 MMA issuers, 3 TMEM allocator, 4-11 / 12-19 softmax groups, 20-23 tail rows
constexpr int MAX_TAIL = 4;
constexpr int Q_BYTES = QT * HD * 2;

struct Layout {  // shared-memory byte offsets for a given key-block size
    int k, v, p, out, xch, tail, bar, total, kbytes, pbytes;
};
__host__ __device__ inline Layout make_layout(int KB) {
    Layout l;
    l.kbytes = KB * HD * 2;        // [KB rows][128 B], SWIZZLE_128B
    l.pbytes = (KB / 8) * QT * 16; // KB/8 chunks of [128 rows][8 keys]
    l.k = 2 * Q_BYTES;
    l.v = l.k + 2 * l.kbytes;
    l.p = l.v + 2 * l.kbytes;
    l.out = l.p + 2 * l.pbytes;
    l.xch = l.out + 8 * 4096;
    l.tail = l.xch + 2 * 2 * 128 * 8;                      // tail rows: q [4][64] fp32 + per-warp partials [4][4][66] fp32
    l.bar = l.tail + MAX_TAIL * HD * 4 + 4 * MAX_TAIL * 66 * 4;
    l.total = l.bar + 256 + 1024;
    return l;
}

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

// Row maximum over 16 S values starting at key column c0 of the block; columns >= nv are padding.
__device__ __forceinline__ float max16(const uint32_t (&r)[16], int c0, int nv, float mx) {
    if (c0 + 16 <= nv) {
#pragma unroll
        for (int i = 0; i < 16; i += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(r[i]), __uint_as_float(r[i + 1])));
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) mx = fmaxf(mx, (c0 + i < nv) ? __uint_as_float(r[i]) : -INFINITY);
    }
    return mx;
}

// exp2(c s - c max) of 16 S values -> two 16-byte P chunks in shared memory; accumulates the row sums.
template <typename T>
__device__ __forceinline__ void exp16(const uint32_t (&r)[16], int c0, int nv, float c, float neg_mxs, uint32_t prow,
                                      float& sum0, float& sum1) {
    const bool nomask = (c0 + 16 <= nv);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        float p[8];
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
            float a0, a1;
            ptx::fma2(a0, a1, __uint_as_float(r[8 * q + i]), __uint_as_float(r[8 * q + i + 1]), c, neg_mxs);
            p[i] = ptx::ex2_approx(a0);
            p[i + 1] = ptx::ex2_approx(a1);
        }
        if (!nomask) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (c0 + 8 * q + i >= nv) p[i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < 8; i += 2) ptx::add2(sum0, sum1, p[i], p[i + 1]);
        ptx::st_shared_v4(prow + ((c0 >> 3) + q) * 2048, pack2<T>(p[0], p[1]), pack2<T>(p[2], p[3]), pack2<T>(p[4], p[5]),
                          pack2<T>(p[6], p[7]));
    }
}

template <typename T>
__global__ void __launch_bounds__(ATL_THREADS, 1)
    attention_tcl_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv,
                         const __grid_constant__ CUtensorMap map_out, int n_items, int N, int H, int KB, int nkb, int n_qt,
                         int n_tail, float scale_log2e, const T* __restrict__ qkv, T* __restrict__ out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    const uint32_t sbase = ptx::smem_u32(smem);
    const Layout lay = make_layout(KB);

    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + lay.bar);
    uint64_t* q_full = bars + 0;
    uint64_t* q_empty = bars + 1;
    uint64_t* k_full = bars + 2;    // [2]
    uint64_t* k_empty = bars + 4;   // [2]
    uint64_t* v_full = bars + 6;    // [2]
    uint64_t* v_empty = bars + 8;   // [2]
    uint64_t* s_full = bars + 10;   // [2] per lane
    uint64_t* s_free = bars + 12;   // [2]
    uint64_t* p_full = bars + 14;   // [2]
    uint64_t* p_empty = bars + 16;  // [2]
    uint64_t* o_full = bars + 18;   // [2]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 20);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int D = H * HD;
    const int rounds = (n_qt + 1) / 2;
    const int nsteps = 2 * nkb;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&map_q);
        ptx::prefetch_tensormap(&map_kv);
        ptx::prefetch_tensormap(&map_out);
        ptx::mbar_init(q_full, 1);
        ptx::mbar_init(q_empty, 2);  // one commit per lane
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&k_full[i], 1);
            ptx::mbar_init(&k_empty[i], 3);  // both lanes' MMA commits + the tail group
            ptx::mbar_init(&v_full[i], 1);
            ptx::mbar_init(&v_empty[i], 3);
            ptx::mbar_init(&s_full[i], 1);
            ptx::mbar_init(&s_free[i], 1);   // elected thread, after the lane's 8 softmax warps have synchronised
            ptx::mbar_init(&p_full[i], 1);
            ptx::mbar_init(&p_empty[i], 1);
            ptx::mbar_init(&o_full[i], 1);
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
            int it = 0, kc = 0, vc = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const int ritem = n_items - 1 - item;   // last frames first (L2 residency of the freshly written qkv rows, see attention_tc.cu)
                const int r = ritem % rounds, bh = ritem / rounds;
                const int b = bh / H, h = bh - b * H;
                const int row0 = b * N;
                ptx::mbar_wait(q_empty, (it & 1) ^ 1);
                ptx::mbar_arrive_expect_tx(q_full, 2 * Q_BYTES);
                ptx::tma_load_2d(&map_q, q_full, smem, h * HD, row0 + min(2 * r, n_qt - 1) * QT);
                ptx::tma_load_2d(&map_q, q_full, smem + Q_BYTES, h * HD, row0 + min(2 * r + 1, n_qt - 1) * QT);
                for (int s = 0; s < nsteps; ++s) {
                    const int j = (s >= nkb) ? s - nkb : s;
                    const int kb = kc & 1;
                    ptx::mbar_wait(&k_empty[kb], ((kc >> 1) & 1) ^ 1);
                    ptx::mbar_arrive_expect_tx(&k_full[kb], lay.kbytes);
                    ptx::tma_load_2d(&map_kv, &k_full[kb], smem + lay.k + kb * lay.kbytes, D + h * HD, row0 + j * KB);
                    ++kc;
                    if (s >= nkb) {
                        const int vb = vc & 1;
                        ptx::mbar_wait(&v_empty[vb], ((vc >> 1) & 1) ^ 1);
                        ptx::mbar_arrive_expect_tx(&v_full[vb], lay.kbytes);
                        ptx::tma_load_2d(&map_kv, &v_full[vb], smem + lay.v + vb * lay.kbytes, 2 * D + h * HD, row0 + j * KB);
                        ++vc;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1 || warp == 2) {
        if (lane == 0) {
            // ===================== MMA issuer of lane L =====================
            const int L = warp - 1;
            constexpr bool kIsBf16 = std::is_same<T, __nv_bfloat16>::value;
            const uint32_t idesc_o = ptx::make_idesc_f16_bmn(kIsBf16, QT, HD);
            const uint32_t tmem_s = tmem_base + L * 256;
            const uint32_t tmem_o = tmem_base + L * 256 + O_COL;
            const uint32_t sp = sbase + lay.p + L * lay.pbytes;
            int it = 0, kc = 0, vc = 0, sc = 0, pc = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const int slot = (L + it) & 1;  // which of the round's two query tiles
                ptx::mbar_wait(q_full, it & 1);
                const uint64_t dq = ptx::make_kmajor_sw128_desc(sbase + slot * Q_BYTES);
                for (int s = 0; s < nsteps; ++s) {
                    const int j = (s >= nkb) ? s - nkb : s;
                    const int n16 = (min(KB, N - j * KB) + 15) & ~15;  // keys of this block, padded to the UMMA granularity
                    const int kb = kc & 1;
                    ptx::mbar_wait(&k_full[kb], (kc >> 1) & 1);
                    ptx::mbar_wait(&s_free[L], (sc & 1) ^ 1);  // the softmax warps are done with the previous S
                    ptx::tcgen05_fence_after();
                    const uint64_t dk = ptx::make_kmajor_sw128_desc(sbase + lay.k + kb * lay.kbytes);
                    const uint32_t idesc_s = ptx::make_idesc_f16(kIsBf16, QT, n16);
#pragma unroll
                    for (int k = 0; k < HD / 16; ++k) ptx::umma_f16<1>(tmem_s, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
                    ptx::umma_commit<1>(&s_full[L]);
                    ptx::umma_commit<1>(&k_empty[kb]);
                    if (s == nsteps - 1) ptx::umma_commit<1>(q_empty);
                    ++kc;
                    ++sc;
                    if (s >= nkb) {
                        const int vb = vc & 1;
                        ptx::mbar_wait(&v_full[vb], (vc >> 1) & 1);
                        ptx::mbar_wait(&p_full[L], pc & 1);  // P_j is in shared memory
                        ptx::tcgen05_fence_after();
                        const uint32_t sv = sbase + lay.v + vb * lay.kbytes;
                        for (int ks = 0; ks < n16 / 16; ++ks) {
                            const uint64_t da = ptx::make_smem_desc(sp + ks * 4096, 2048, 128, 0);
                            const uint64_t db = ptx::make_smem_desc(sv + ks * 2048, 0, 1024, 2);
                            ptx::umma_f16<1>(tmem_o, da, db, idesc_o, (j | ks) != 0);
                        }
                        ptx::umma_commit<1>(&p_empty[L]);
                        ptx::umma_commit<1>(&v_empty[vb]);
                        if (j == nkb - 1) ptx::umma_commit<1>(&o_full[L]);
                        ++vc;
                        ++pc;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp >= 4 && warp < 20) {
        // ===================== softmax + output group of lane L =====================
        const int L = (warp - 4) >> 3;
        const int half = ((warp - 4) >> 2) & 1;
        const int quarter = warp & 3;
        const uint32_t pair_bar = 1 + L * 4 + quarter;
        const uint32_t group_bar = 9 + L;
        const bool poller = ((warp - 4) & 7) == 0;
        const int row_in_tile = quarter * 32 + lane;
        const uint32_t stage = sbase + lay.out + (L * 4 + quarter) * 4096;
        const void* stage_ptr = smem + lay.out + (L * 4 + quarter) * 4096;
        float2* xch_mine = reinterpret_cast<float2*>(smem + lay.xch) + (L * 2 + half) * 128 + row_in_tile;
        float2* xch_other = reinterpret_cast<float2*>(smem + lay.xch) + (L * 2 + (half ^ 1)) * 128 + row_in_tile;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + L * 256;
        const uint32_t prow = sbase + lay.p + L * lay.pbytes + row_in_tile * 16;
        int it = 0, sc = 0, pc = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const int ritem = n_items - 1 - item;
            const int r = ritem % rounds, bh = ritem / rounds;
            const int b = bh / H, h = bh - b * H;
            const int slot = (L + it) & 1;
            const int t_raw = 2 * r + slot;
            const int t = min(t_raw, n_qt - 1);
            const bool store_ok = (t_raw < n_qt);  // otherwise this lane only keeps the barriers company
            const bool warp_valid = (t * QT + quarter * 32) < N;
            float mx = -INFINITY, neg_mxs = 0.f, sum0 = 0.f, sum1 = 0.f;

            for (int s = 0; s < nsteps; ++s) {
                const int j = (s >= nkb) ? s - nkb : s;
                const int nv = min(KB, N - j * KB);  // valid keys of this block
                const int n16 = (nv + 15) & ~15;
                // this thread's key columns of the block: the first ceil(chunks/2) 16-column chunks go to half 0
                const int split = ((n16 / 16 + 1) / 2) * 16;
                const int cb = half ? split : 0, ce = half ? n16 : split;
                if (s < nkb) {
                    // ---- sweep 0: running row maximum ----
                    if (poller) ptx::mbar_wait(&s_full[L], sc & 1);
                    ptx::named_bar_sync(group_bar, 256);
                    ptx::tcgen05_fence_after();
                    if (warp_valid) {
                        for (int c0 = cb; c0 < ce; c0 += 16) {
                            uint32_t rr[16];
                            ptx::tmem_ld_32x32b_x16(taddr + c0, rr);
                            ptx::tmem_ld_wait();
                            mx = max16(rr, c0, nv, mx);
                        }
                    }
                    ptx::tcgen05_fence_before();
                    ptx::named_bar_sync(group_bar, 256);  // all 8 warps have read S
                    if (poller && lane == 0) ptx::mbar_arrive(&s_free[L]);
                    if (s == nkb - 1) {
                        // exchange the two halves' maxima (half 0 of block 0 always holds key 0: the result is finite)
                        xch_mine->x = mx;
                        ptx::named_bar_sync(pair_bar, 64);
                        mx = fmaxf(mx, xch_other->x);
                        neg_mxs = -mx * scale_log2e;
                    }
                } else {
                    // ---- sweep 1: exponentials, P, row sums ----
                    if (poller) {
                        ptx::mbar_wait(&s_full[L], sc & 1);
                        ptx::mbar_wait(&p_empty[L], (pc & 1) ^ 1);                        // previous PV has read P
                        ptx::mbar_wait(&p_full[L ^ 1], L == 0 ? ((pc & 1) ^ 1) : (pc & 1));  // the lanes take turns
                    }
                    ptx::named_bar_sync(group_bar, 256);
                    ptx::tcgen05_fence_after();
                    if (warp_valid) {
                        for (int c0 = cb; c0 < ce; c0 += 16) {
                            uint32_t rr[16];
                            ptx::tmem_ld_32x32b_x16(taddr + c0, rr);
                            ptx::tmem_ld_wait();
                            exp16<T>(rr, c0, nv, scale_log2e, neg_mxs, prow, sum0, sum1);
                        }
                    }
                    ptx::fence_proxy_async_smem();
                    ptx::tcgen05_fence_before();
                    ptx::named_bar_sync(group_bar, 256);
                    if (poller && lane == 0) {
                        ptx::mbar_arrive(&p_full[L]);
                        ptx::mbar_arrive(&s_free[L]);
                    }
                    ++pc;
                }
                ++sc;
            }

            // ---- output: O / rowsum -> 16-bit -> staging -> TMA store ----
            xch_mine->y = sum0 + sum1;
            if (half == 0 && lane == 0) ptx::bulk_wait_group_read<0>();  // the staging tile's previous store has read it
            if (poller) ptx::mbar_wait(&o_full[L], it & 1);
            ptx::named_bar_sync(group_bar, 256);
            ptx::tcgen05_fence_after();
            if (warp_valid) {
                uint32_t rr[32];
                ptx::tmem_ld_32x32b_x32(taddr + O_COL + half * 32, rr);
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
                        ptx::mul2(a0, a1, __uint_as_float(rr[8 * c + 2 * i]), __uint_as_float(rr[8 * c + 2 * i + 1]), inv_sum);
                        u[i] = pack2<T>(a0, a1);
                    }
                    ptx::st_shared_v4(srow + ((static_cast<uint32_t>(4 * half + c) ^ swz) << 4), u[0], u[1], u[2], u[3]);
                }
                ptx::fence_proxy_async_smem();
            }
            ptx::tcgen05_fence_before();
            ptx::named_bar_sync(pair_bar, 64);  // both halves of the staging tile written, O read by both warps
            if (half == 0 && lane == 0 && warp_valid && store_ok) {
                ptx::tma_store_3d(&map_out, stage_ptr, h * HD, t * QT + quarter * 32, b);
                ptx::bulk_commit_group();
            }
        }
        if (half == 0 && lane == 0) ptx::bulk_wait_group<0>();
    }

    if (warp >= 20) {
        // ===================== tail rows: queries n_qt*128 .. N-1 of each (frame, head), SIMT =====================
        // The four warps take a quarter of each key block each.  Per warp and row: running max m, sum l and the output
        // accumulator (two channels per lane), merged across the warps at the end of the item.  Keys and values are read
        // from the SWIZZLE_128B tiles the TMA producer filled for the tensor-core lanes, during sweep 1 (when both the K
        // and the V block of a step are resident); every step's buffers are released through the same k_empty / v_empty
        // barriers the MMA issuers commit to.
        const int tw = warp - 20;
        float* tq = reinterpret_cast<float*>(smem + lay.tail);                 // [MAX_TAIL][64], pre-scaled by c
        float* tpart = tq + MAX_TAIL * HD;                                      // [4 warps][MAX_TAIL][66]
        const int row_first = n_qt * QT;
        int it = 0, kc = 0, vc = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const int ritem = n_items - 1 - item;
            const int r = ritem % rounds, bh = ritem / rounds;
            const int b = bh / H, h = bh - b * H;
            const bool work = (n_tail > 0) && (r == 0);
            float m_run[MAX_TAIL], l_run[MAX_TAIL], acc0[MAX_TAIL], acc1[MAX_TAIL];
#pragma unroll
            for (int q = 0; q < MAX_TAIL; ++q) m_run[q] = -INFINITY, l_run[q] = 0.f, acc0[q] = 0.f, acc1[q] = 0.f;
            if (work) {
                // q rows (pre-multiplied by scale * log2 e) into shared memory: thread i < n_tail*64 loads one channel
                for (int i = threadIdx.x - 20 * 32; i < n_tail * HD; i += 128) {
                    const int q = i / HD, d = i - q * HD;
                    tq[i] = static_cast<float>(qkv[(static_cast<int64_t>(b) * N + row_first + q) * (3 * H * HD) + h * HD + d]) *
                            scale_log2e;
                }
                ptx::named_bar_sync(11, 128);
            }
            for (int s = 0; s < nsteps; ++s) {
                const int j = (s >= nkb) ? s - nkb : s;
                const int kb = kc & 1;
                if (s < nkb || !work) {
                    // nothing to compute in this step: one thread waits for the data phase and releases the buffers
                    if (tw == 0 && lane == 0) {
                        ptx::mbar_wait(&k_full[kb], (kc >> 1) & 1);
                        ptx::mbar_arrive(&k_empty[kb]);
                        if (s >= nkb) {
                            const int vb = vc & 1;
                            ptx::mbar_wait(&v_full[vb], (vc >> 1) & 1);
                            ptx::mbar_arrive(&v_empty[vb]);
                        }
                    }
                } else {
                    const int vb = vc & 1;
                    if (tw == 0) {
                        ptx::mbar_wait(&k_full[kb], (kc >> 1) & 1);
                        ptx::mbar_wait(&v_full[vb], (vc >> 1) & 1);
                    }
                    ptx::named_bar_sync(11, 128);
                    const int nv = min(KB, N - j * KB);
                    const int per_warp = (nv + 3) / 4;
                    const int k_begin = tw * per_warp, k_end = min(nv, k_begin + per_warp);
                    const uint32_t sk = sbase + lay.k + kb * lay.kbytes;
                    const uint32_t sv = sbase + lay.v + vb * lay.kbytes;
                    for (int q = 0; q < n_tail; ++q) {
                        const float4* q4 = reinterpret_cast<const float4*>(tq + q * HD);
                        for (int k0 = k_begin; k0 < k_end; k0 += 32) {
                            const int k = k0 + lane;
                            float sc = -INFINITY;
                            if (k < k_end) {
                                sc = 0.f;
#pragma unroll
                                for (int c = 0; c < 8; ++c) {
                                    uint32_t w0, w1, w2, w3;
                                    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                                                 : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3)
                                                 : "r"(sk + k * 128 + ((c ^ (k & 7)) << 4)));
                                    const float4 qa = q4[2 * c], qb = q4[2 * c + 1];
                                    const uint32_t w[4] = {w0, w1, w2, w3};
                                    const float qq[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
#pragma unroll
                                    for (int i = 0; i < 4; ++i) {
                                        float lo, hi;
                                        if constexpr (std::is_same<T, __nv_bfloat16>::value) {
                                            lo = __uint_as_float(w[i] << 16);
                                            hi = __uint_as_float(w[i] & 0xffff0000u);
                                        } else {
                                            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
                                            lo = f.x;
                                            hi = f.y;
                                        }
                                        sc = fmaf(qq[2 * i], lo, sc);
                                        sc = fmaf(qq[2 * i + 1], hi, sc);
                                    }
                                }
                            }
                            // block-of-32 maximum, rescale the running state, then the weighted values
                            float bm = sc;
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, o));
                            const float m_new = fmaxf(m_run[q], bm);
                            const float corr = ptx::ex2_approx(m_run[q] - m_new);  // first block: exp2(-inf) = 0
                            const float p = (k < k_end) ? ptx::ex2_approx(sc - m_new) : 0.f;
                            float ps = p;
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, o);
                            l_run[q] = l_run[q] * corr + ps;
                            acc0[q] *= corr;
                            acc1[q] *= corr;
                            m_run[q] = m_new;
                            const int cnt = min(32, k_end - k0);
                            for (int kk = 0; kk < cnt; ++kk) {
                                const float pk = __shfl_sync(0xffffffffu, p, kk);
                                const int key = k0 + kk;
                                uint32_t vw;
                                asm volatile("ld.shared.b32 %0, [%1];"
                                             : "=r"(vw)
                                             : "r"(sv + key * 128 + (((lane >> 2) ^ (key & 7)) << 4) + (lane & 3) * 4));
                                float lo, hi;
                                if constexpr (std::is_same<T, __nv_bfloat16>::value) {
                                    lo = __uint_as_float(vw << 16);
                                    hi = __uint_as_float(vw & 0xffff0000u);
                                } else {
                                    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&vw));
                                    lo = f.x;
                                    hi = f.y;
                                }
                                acc0[q] = fmaf(pk, lo, acc0[q]);
                                acc1[q] = fmaf(pk, hi, acc1[q]);
                            }
                        }
                    }
                    ptx::named_bar_sync(11, 128);  // all four warps are done with this step's K and V tiles
                    if (tw == 0 && lane == 0) {
                        ptx::mbar_arrive(&k_empty[kb]);
                        ptx::mbar_arrive(&v_empty[vb]);
                    }
                }
                ++kc;
                if (s >= nkb) ++vc;
            }
            if (work) {
                // merge the four warps' partial softmaxes and write the rows
                for (int q = 0; q < n_tail; ++q) {
                    float* pp = tpart + (tw * MAX_TAIL + q) * 66;
                    pp[2 * lane] = acc0[q];
                    pp[2 * lane + 1] = acc1[q];
                    if (lane == 0) {
                        pp[64] = m_run[q];
                        pp[65] = l_run[q];
                    }
                }
                ptx::named_bar_sync(11, 128);
                if (tw == 0) {
                    for (int q = 0; q < n_tail; ++q) {
                        float M = -INFINITY;
#pragma unroll
                        for (int w = 0; w < 4; ++w) M = fmaxf(M, tpart[(w * MAX_TAIL + q) * 66 + 64]);
                        float den = 0.f, o0 = 0.f, o1 = 0.f;
#pragma unroll
                        for (int w = 0; w < 4; ++w) {
                            const float* pp = tpart + (w * MAX_TAIL + q) * 66;
                            const float f = ptx::ex2_approx(pp[64] - M);  // a warp without keys has m = -inf, f = 0
                            den = fmaf(f, pp[65], den);
                            o0 = fmaf(f, pp[2 * lane], o0);
                            o1 = fmaf(f, pp[2 * lane + 1], o1);
                        }
                        const float inv = 1.0f / den;
                        const uint32_t packed = pack2<T>(o0 * inv, o1 * inv);
                        *reinterpret_cast<uint32_t*>(out + (static_cast<int64_t>(b) * N + row_first + q) * (H * HD) + h * HD + 2 * lane) =
                            packed;
                    }
                }
                ptx::named_bar_sync(11, 128);  // tpart / tq are reused by the next item
            }
        }
    }

    ptx::tcgen05_fence_before();
    __syncthreads();
    if (warp == 3) ptx::tmem_dealloc<1>(tmem_base, 512);
}

// Query rows [row_first, N) of every (frame, head): plain fp32 SIMT softmax attention over all N keys, one 4-warp CTA per
// row (each warp a contiguous quarter of the keys, each lane every 32nd key of it, 16-byte loads of the K / V rows).
// Only used for the few rows past the last full 128-row tile (CLIP: the 257th token).
template <typename T>
__device__ __forceinline__ void load_row64(const T* p, float (&x)[HD]) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(p) + c);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if constexpr (std::is_same<T, __nv_bfloat16>::value) {
                x[8 * c + 2 * i] = __uint_as_float(w[i] << 16);
                x[8 * c + 2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
            } else {
                const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
                x[8 * c + 2 * i] = f.x;
                x[8 * c + 2 * i + 1] = f.y;
            }
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(128)
    attention_rows_kernel(const T* __restrict__ qkv, T* __restrict__ out, int N, int H, int row_first, float scale) {
    __shared__ float s_red[4];
    __shared__ float s_acc[4][HD];
    const int nrows = N - row_first;
    const int row = row_first + blockIdx.x % nrows;
    const int bh = blockIdx.x / nrows;
    const int b = bh / H, h = bh - b * H;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t tok_stride = static_cast<int64_t>(3) * H * HD;
    const T* base = qkv + static_cast<int64_t>(b) * N * tok_stride + h * HD;
    float q[HD];
    load_row64<T>(base + static_cast<int64_t>(row) * tok_stride, q);
#pragma unroll
    for (int d = 0; d < HD; ++d) q[d] *= scale;
    const int per_warp = (N + 3) / 4;
    const int k_begin = warp * per_warp, k_end = min(N, k_begin + per_warp);
    constexpr int MAXK = 6;  // keys per lane: N <= 768 -> 192 per warp -> 6 per lane
    float sc[MAXK];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < MAXK; ++i) {
        const int j = k_begin + lane + 32 * i;
        float s = -INFINITY;
        if (j < k_end) {
            float kk[HD];
            load_row64<T>(base + static_cast<int64_t>(H) * HD + static_cast<int64_t>(j) * tok_stride, kk);
            s = 0.f;
#pragma unroll
            for (int d = 0; d < HD; ++d) s = fmaf(q[d], kk[d], s);
        }
        sc[i] = s;
        mx = fmaxf(mx, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) s_red[warp] = mx;
    __syncthreads();
    mx = fmaxf(fmaxf(s_red[0], s_red[1]), fmaxf(s_red[2], s_red[3]));
    __syncthreads();
    float acc[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) acc[d] = 0.f;
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < MAXK; ++i) {
        const int j = k_begin + lane + 32 * i;
        if (j < k_end) {
            const float p = __expf(sc[i] - mx);
            sum += p;
            float vv[HD];
            load_row64<T>(base + static_cast<int64_t>(2) * H * HD + static_cast<int64_t>(j) * tok_stride, vv);
#pragma unroll
            for (int d = 0; d < HD; ++d) acc[d] = fmaf(p, vv[d], acc[d]);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) s_red[warp] = sum;
#pragma unroll
    for (int d = 0; d < HD; ++d) {
        float v = acc[d];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == (d & 31)) s_acc[warp][d] = v;
    }
    __syncthreads();
    if (threadIdx.x < HD) {
        const int d = threadIdx.x;
        const float tot = s_red[0] + s_red[1] + s_red[2] + s_red[3];
        const float v = s_acc[0][d] + s_acc[1][d] + s_acc[2][d] + s_acc[3][d];
        out[(static_cast<int64_t>(b) * N + row) * (H * HD) + h * HD + d] = static_cast<T>(v / tot);
    }
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

template <typename T>
int launch_long(const AttentionMaps& m, float scale, cudaStream_t stream) {
    auto kern = attention_tcl_kernel<T>;
    const Layout lay = make_layout(m.KB);
    // per launch (cheap): the attribute belongs to the current device's copy of the function
    VIDIL_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, make_layout(MAX_KB).total));
    const int rounds = (m.n_qt + 1) / 2;
    const int n_items = m.B * m.H * rounds;
    int grid = gemm_num_sms();
    if (grid > n_items) grid = n_items;
    if (grid < 1) return 1;
    VIDIL_CUDA_OK(launch_pdl(kern, dim3(grid), dim3(ATL_THREADS), lay.total, stream, m.q, m.kv, m.out, n_items, m.N, m.H, m.KB, m.nkb,
                             m.n_qt, m.n_tail, scale * 1.4426950408889634f, reinterpret_cast<const T*>(m.qkv),
                             reinterpret_cast<T*>(m.out_ptr)));
    count_launches(1);
    const int row_first = m.n_qt * QT;
    if (m.n_tail == 0 && row_first < m.N) {
        const int rows = m.B * m.H * (m.N - row_first);
        attention_rows_kernel<T><<<rows, 128, 0, stream>>>(reinterpret_cast<const T*>(m.qkv), reinterpret_cast<T*>(m.out_ptr),
                                                           m.N, m.H, row_first, scale);
        VIDIL_CUDA_OK(cudaGetLastError());
        count_launches(1);
    }
    return 0;
}

}  // namespace

bool attention_tcl_supported(int N) { return N > 128 && N <= 768; }

int attention_tcl_prepare(AttentionMaps& m, const void* qkv, void* out, DType dt, int B, int N, int H) {
    EncodeTiledFn fn = encode_fn();
    if (fn == nullptr) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return 1;
    }
    if ((reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) {
        set_error("attention: qkv and out must be 16-byte aligned");
        return 1;
    }
    // key blocks of at most 144 keys, all but the last the same multiple of 16
    const int nkb = (N + MAX_KB - 1) / MAX_KB;
    const int KB = (((N + nkb - 1) / nkb) + 15) & ~15;
    // query tiles for the tcgen05 kernel; a short tail of rows goes to the per-row kernel instead of a whole extra tile
    // (<= 4 rows: tail warps inside the kernel; <= 16: the per-row kernel)
    const int rem = N % QT;
    const int n_qt = (rem != 0 && rem <= 16) ? N / QT : (N + QT - 1) / QT;
    m.n_tail = (rem != 0 && rem <= MAX_TAIL) ? rem : 0;
    const CUtensorMapDataType cdt = (dt == DT_BF16) ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    const cuuint32_t estr3[3] = {1, 1, 1};
    {
        const cuuint64_t dims[2] = {static_cast<cuuint64_t>(3) * H * HD, static_cast<cuuint64_t>(B) * N};
        const cuuint64_t strides[1] = {static_cast<cuuint64_t>(3) * H * HD * 2};
        const cuuint32_t box_q[2] = {HD, QT};
        const cuuint32_t box_kv[2] = {HD, static_cast<cuuint32_t>(KB)};
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
    m.KB = KB;
    m.nkb = nkb;
    m.n_qt = n_qt;
    m.is_long = true;
    return 0;
}

int attention_tcl_run(const AttentionMaps& m, float scale, cudaStream_t stream) {
    if (gemm_num_sms() == 0) return 1;
    if (m.dt == DT_BF16) return launch_long<__nv_bfloat16>(m, scale, stream);
    return launch_long<__half>(m, scale, stream);
}

}  // namespace vidil
