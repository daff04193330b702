// D[M,N] = A[M,K] * W[N,K]^T with a fused epilogue — the dense contraction behind every Linear on
// the frame-encoding path (reference: models/vit.py:51,53,72,84 qkv/proj; :29,31,36,39 fc1/fc2;
// the patch-embed Conv2d at :144-145 as a GEMM over im2col rows; CLIP's visual_projection and the
// image x phrase-bank similarity at run_visual_tokenization.py:276).
//
// sm_100a design (one persistent kernel, warp-specialised):
//   warp 0 / lane 0 : TMA producer   — cp.async.bulk.tensor 2-D tiles of A and W into a ring of
//                                      SWIZZLE_128B shared-memory stages, completion on mbarriers
//   warp 1 / lane 0 : MMA issuer     — tcgen05.mma kind::f16 (fp16 or bf16 in, fp32 accumulate in TMEM);
//                                      tcgen05.commit releases smem stages and publishes accumulators
//   warp 2          : TMEM allocator — 512 columns = two 128x256 fp32 accumulators (double buffered)
//   warps 4..11     : epilogue       — tcgen05.ld TMEM->registers, bias / GELU; the 32-row x 128-byte piece is
//                                      staged in SWIZZLE_128B shared memory and leaves through TMA: a bulk tensor
//                                      store for 16-bit / fp32 outputs, a bulk tensor REDUCE-ADD for the fp32
//                                      residual stream (the += happens at L2; the old value is never read by the
//                                      SM).  Overlaps the next tile's MMAs.
// With CTA_GROUP == 2 two CTAs of a cluster (one TPC) run cta_group::2 MMAs on a 256x256 tile: each
// CTA stages its own 128 rows of A and half (128 rows) of the W tile, halving shared-memory and L2
// operand traffic per FLOP; accumulator rows [128r, 128r+128) live in CTA r's TMEM.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>

#include <string>
#include <type_traits>

#include "kernels.h"
#include "ptx.cuh"

namespace vidil {

namespace {

constexpr int BLOCK_M = 128;  // rows per CTA (UMMA M = 128 * CTA_GROUP)
constexpr int BLOCK_N = 256;  // UMMA N
constexpr int BLOCK_K = 64;   // one 128-byte swizzle atom of 16-bit elements
constexpr int UMMA_K = 16;
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = (4 + NUM_EPI_WARPS) * 32;
constexpr int NUM_ACC = 2;  // TMEM accumulator buffers
constexpr int TMEM_COLS = NUM_ACC * BLOCK_N;

template <int CG>
struct Cfg {
    static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;         // 16 KB
    static constexpr int W_ROWS = BLOCK_N / CG;                   // rows of W this CTA stages
    static constexpr int W_BYTES = W_ROWS * BLOCK_K * 2;          // 32 KB or 16 KB
    static constexpr int STAGE_BYTES = A_BYTES + W_BYTES;         // 48 KB or 32 KB
    static constexpr int STAGES = (CG == 1) ? 4 : 6;              // 192 KB of operand ring either way
    static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
    static constexpr int OUT_BUF_BYTES = 32 * 128;                // one epilogue piece: 32 rows x 128 bytes
    static constexpr int OUT_BYTES = NUM_EPI_WARPS * OUT_BUF_BYTES;      // one staging tile per epilogue warp: 32 KB
    static constexpr int BAR_BYTES = 256;
    static constexpr int SMEM_BYTES = RING_BYTES + OUT_BYTES + BAR_BYTES + 1024;  // +1024 alignment slack
};

struct EpiArgs {
    const float* bias;
    void* out;
    int64_t ldo;
    const float* pos;
    int patches_per_frame;
};

// Exact-erf GELU 0.5 x (1 + erf(x / sqrt 2)) = max(x, 0) - |x| * 0.5 erfc(|x| / sqrt 2), with
//   0.5 erfc(u / sqrt 2) = 2^Q(u),  Q = degree-7 near-minimax polynomial of log2(erfc(u / sqrt 2)) - 1 on u in [0, 6]
// (that logarithm is smooth and almost quadratic, which is why so low a degree is enough; |u| is clamped to 6, where
// the erfc term is 1e-9).  |abs error| <= 6.1e-7 over all x in fp32 arithmetic — the size of erff's own error and three
// orders below the rounding of the 16-bit output.  Two elements at a time with packed fp32x2 FMAs: 9 FFMA2 + 4 FMNMX +
// 2 MUFU per PAIR, against ~30 instructions per element for erff and 12 + 2 MUFU for the classic Abramowitz-Stegun
// 7.1.26 form — the fc1 epilogue's cost is its instruction count (measured: 122.6 M warp instructions and 330 us with
// A-S, 93.3 M and 307 us with this form, 44.9 M and 277 us with no activation at all).
// The x < 0 branch has no 1 - erf cancellation, so the negative tail keeps full relative accuracy.
__device__ __forceinline__ void gelu_erf2(float& x0, float& x1) {
    float s0, s1, q0, q1;
    ptx::fma2(s0, s1, fminf(fabsf(x0), 6.0f), fminf(fabsf(x1), 6.0f), 2.0f / 6.0f, -1.0f);  // [0, 6] -> [-1, 1]
    ptx::fma2(q0, q1, s0, s1, -4.132612573e-03f, 1.676645333e-02f);
    ptx::fma2v(q0, q1, s0, s1, q0, q1, -4.076132380e-02f, -4.076132380e-02f);
    ptx::fma2v(q0, q1, s0, s1, q0, q1, 9.175150894e-02f, 9.175150894e-02f);
    ptx::fma2v(q0, q1, s0, s1, q0, q1, -2.039687718e-01f, -2.039687718e-01f);
    ptx::fma2v(q0, q1, s0, s1, q0, q1, -6.033998262e+00f, -6.033998262e+00f);
    ptx::fma2v(q0, q1, s0, s1, q0, q1, -1.420955799e+01f, -1.420955799e+01f);
    ptx::fma2v(q0, q1, s0, s1, q0, q1, -9.532934736e+00f, -9.532934736e+00f);
    ptx::fma2v(x0, x1, -fabsf(x0), -fabsf(x1), ptx::ex2_approx(q0), ptx::ex2_approx(q1), fmaxf(x0, 0.0f), fmaxf(x1, 0.0f));
}
__device__ __forceinline__ float gelu_erf(float x) {
    float y = x, dummy = 0.f;
    gelu_erf2(y, dummy);
    return y;
}
__device__ __forceinline__ float quick_gelu(float x) {
    return x * ptx::rcp_approx(1.0f + ptx::ex2_approx(x * (-1.702f * 1.4426950408889634f)));
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

// One 32-column slice of one accumulator row: v[] holds acc, (row, col0) are global coordinates.
template <typename T, int EPI>
__device__ __forceinline__ void epilogue_chunk(float (&v)[32], const EpiArgs& e, int row, int col0, int N) {
    const bool full = (col0 + 32 <= N);
    if (e.bias != nullptr) {
        if (full) {
            const float4* b4 = reinterpret_cast<const float4*>(e.bias + col0);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float4 b = __ldg(b4 + i);
                v[4 * i + 0] += b.x;
                v[4 * i + 1] += b.y;
                v[4 * i + 2] += b.z;
                v[4 * i + 3] += b.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (col0 + i < N) v[i] += __ldg(e.bias + col0 + i);
        }
    }
    if constexpr (EPI == EPI_GELU) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
    } else if constexpr (EPI == EPI_QUICKGELU) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = quick_gelu(v[i]);
    }

    if constexpr (EPI == EPI_TOP4) {
        // Four largest of the 32 scores, each tagged with its position in the group in the 5 low mantissa bits (a perturbation
        // of <= 2^-18 relative, far inside the margin the exact re-ranking allows for): seven FMNMX + one LOP3 per score, no
        // index registers, no divergence.  Columns beyond N (zero-filled W rows of the last tile) never qualify.  Four rather
        // than two: the selection has to re-score a WHOLE group when the last entry it knows of is still above its bar, and
        // with two entries that happened for 4 % of the frames (two of a frame's ~6 candidates in one group of 313), which made
        // the one-wave selection kernel as slow as its unluckiest CTA; with four it needs five candidates in one group.
        float m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY, m4 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            float t = __uint_as_float((__float_as_uint(v[i]) & 0xFFFFFFE0u) | static_cast<uint32_t>(i));
            if (!full && col0 + i >= N) t = -INFINITY;
            m4 = fmaxf(m4, fminf(m3, t));
            m3 = fmaxf(m3, fminf(m2, t));
            m2 = fmaxf(m2, fminf(m1, t));
            m1 = fmaxf(m1, t);
        }
        float4* o = reinterpret_cast<float4*>(e.out) + static_cast<int64_t>(row) * (e.ldo >> 2) + (col0 >> 5);
        *o = make_float4(m1, m2, m3, m4);
    } else if constexpr (EPI == EPI_STORE || EPI == EPI_GELU || EPI == EPI_QUICKGELU) {
        T* o = reinterpret_cast<T*>(e.out) + static_cast<int64_t>(row) * e.ldo + col0;
        if (full) {
            uint4* o4 = reinterpret_cast<uint4*>(o);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                uint4 u;
                u.x = pack2<T>(v[8 * i + 0], v[8 * i + 1]);
                u.y = pack2<T>(v[8 * i + 2], v[8 * i + 3]);
                u.z = pack2<T>(v[8 * i + 4], v[8 * i + 5]);
                u.w = pack2<T>(v[8 * i + 6], v[8 * i + 7]);
                o4[i] = u;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (col0 + i < N) o[i] = static_cast<T>(v[i]);
        }
    } else {
        int64_t orow = row;
        if constexpr (EPI == EPI_PATCH) {
            const int f = row / e.patches_per_frame;
            const int p = row - f * e.patches_per_frame;
            orow = static_cast<int64_t>(f) * (e.patches_per_frame + 1) + 1 + p;
            const float* ps = e.pos + static_cast<int64_t>(1 + p) * N + col0;
            if (full) {
                const float4* p4 = reinterpret_cast<const float4*>(ps);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float4 b = __ldg(p4 + i);
                    v[4 * i + 0] += b.x;
                    v[4 * i + 1] += b.y;
                    v[4 * i + 2] += b.z;
                    v[4 * i + 3] += b.w;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (col0 + i < N) v[i] += __ldg(ps + i);
            }
        }
        float* o = reinterpret_cast<float*>(e.out) + orow * e.ldo + col0;
        if (full) {
            float4* o4 = reinterpret_cast<float4*>(o);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float4 r = make_float4(v[4 * i + 0], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                if constexpr (EPI == EPI_RESID) {
                    const float4 old = o4[i];
                    r.x += old.x;
                    r.y += old.y;
                    r.z += old.z;
                    r.w += old.w;
                }
                o4[i] = r;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (col0 + i < N) {
                    if constexpr (EPI == EPI_RESID)
                        o[i] += v[i];
                    else
                        o[i] = v[i];
                }
        }
    }
}

template <typename T, int EPI, int CG>
__global__ void __launch_bounds__(NUM_THREADS, 1)
    gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                        const __grid_constant__ CUtensorMap map_out, int M, int N, int K, EpiArgs epi) {
    using C = Cfg<CG>;
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B tiles must sit on 1024-byte boundaries of the shared address space.
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

    uint8_t* out_stage = smem + C::RING_BYTES;  // per-epilogue-warp staging for TMA stores (1024-byte aligned)
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::RING_BYTES + C::OUT_BYTES);
    uint64_t* empty_bar = full_bar + C::STAGES;
    uint64_t* tmem_full_bar = empty_bar + C::STAGES;
    uint64_t* tmem_empty_bar = tmem_full_bar + NUM_ACC;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + NUM_ACC);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = (CG == 2) ? ptx::cluster_ctarank() : 0u;
    const bool is_leader = (cta_rank == 0);

    if constexpr (CG == 2) ptx::cluster_sync();  // peer CTA is resident before any cross-CTA traffic

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&map_a);
        ptx::prefetch_tensormap(&map_w);
        if constexpr (EPI != EPI_PATCH && EPI != EPI_TOP4) ptx::prefetch_tensormap(&map_out);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            ptx::mbar_init(&full_bar[s], CG);  // one arrive per producing CTA (+ the TMA bytes)
            ptx::mbar_init(&empty_bar[s], 1);  // one tcgen05.commit
        }
        for (int a = 0; a < NUM_ACC; ++a) {
            ptx::mbar_init(&tmem_full_bar[a], 1);                    // one tcgen05.commit
            ptx::mbar_init(&tmem_empty_bar[a], CG * NUM_EPI_WARPS);  // one arrive per epilogue warp
        }
        ptx::fence_mbar_init();
    }
    if (warp == 2) ptx::tmem_alloc<CG>(tmem_ptr_smem, TMEM_COLS);
    ptx::tcgen05_fence_before();
    if constexpr (CG == 2)
        ptx::cluster_sync();
    else
        __syncthreads();
    ptx::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    // PDL: barriers, TMEM and descriptors are set up; the next kernel may start its own set-up on SMs this grid has left, and this
    // grid may now wait for its predecessors' results (every global read and write of the kernel lies below this line).
    ptx::griddep_launch();
    ptx::griddep_wait();

    // Static persistent schedule, identical in every role: cluster c takes tiles c, c+G, c+2G, ...
    // N is the fast index so concurrently running clusters share a few A row-bands and all of W in L2.
    const int tiles_m = (M + BLOCK_M * CG - 1) / (BLOCK_M * CG);
    const int tiles_n = (N + BLOCK_N - 1) / BLOCK_N;
    const int num_tiles = tiles_m * tiles_n;
    const int cluster_id = blockIdx.x / CG;
    const int num_clusters = gridDim.x / CG;
    const int num_kb = K / BLOCK_K;

    if (warp == 0) {
        if (lane == 0) {
            // ===================== TMA producer =====================
            int stage = 0;
            uint32_t phase = 0;
            for (int t = cluster_id; t < num_tiles; t += num_clusters) {
                const int m0 = ((t / tiles_n) * CG + static_cast<int>(cta_rank)) * BLOCK_M;
                const int n0 = (t % tiles_n) * BLOCK_N + static_cast<int>(cta_rank) * C::W_ROWS;
                for (int kb = 0; kb < num_kb; ++kb) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * C::STAGE_BYTES;
                    uint8_t* sw = sa + C::A_BYTES;
                    if (CG == 1) {
                        ptx::mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
                        ptx::tma_load_2d(&map_a, &full_bar[stage], sa, kb * BLOCK_K, m0);
                        ptx::tma_load_2d(&map_w, &full_bar[stage], sw, kb * BLOCK_K, n0);
                    } else {
                        // Both CTAs' bytes are credited to the leader's barrier, which the MMA issuer waits on.
                        if (is_leader)
                            ptx::mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES * CG);
                        else
                            ptx::mbar_arrive_cluster(&full_bar[stage], 0);
                        ptx::tma_load_2d_pair(&map_a, &full_bar[stage], sa, kb * BLOCK_K, m0);
                        ptx::tma_load_2d_pair(&map_w, &full_bar[stage], sw, kb * BLOCK_K, n0);
                    }
                    if (++stage == C::STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0 && is_leader) {
            // ===================== MMA issuer (leader CTA only) =====================
            constexpr bool kIsBf16 = std::is_same<T, __nv_bfloat16>::value;
            constexpr uint32_t idesc = ptx::make_idesc_f16(kIsBf16, BLOCK_M * CG, BLOCK_N);
            int stage = 0;
            uint32_t phase = 0;
            int iter = 0;
            for (int t = cluster_id; t < num_tiles; t += num_clusters, ++iter) {
                const int acc = iter & 1;
                const uint32_t acc_phase = (iter >> 1) & 1;
                ptx::mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);  // epilogue drained this accumulator
                ptx::tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
                for (int kb = 0; kb < num_kb; ++kb) {
                    ptx::mbar_wait(&full_bar[stage], phase);  // TMA bytes of this stage have landed
                    ptx::tcgen05_fence_after();
                    const uint32_t sa = ptx::smem_u32(smem + stage * C::STAGE_BYTES);
                    const uint64_t da = ptx::make_kmajor_sw128_desc(sa);
                    const uint64_t dw = ptx::make_kmajor_sw128_desc(sa + C::A_BYTES);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        // +32 bytes per UMMA_K step inside the swizzle atom: start-address field is >>4
                        ptx::umma_f16<CG>(tmem_d, da + 2 * k, dw + 2 * k, idesc, (kb | k) != 0);
                    }
                    ptx::umma_commit<CG>(&empty_bar[stage]);  // stage reusable once these MMAs retire
                    if (kb == num_kb - 1) ptx::umma_commit<CG>(&tmem_full_bar[acc]);
                    if (++stage == C::STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ===================== Epilogue =====================
        const int ew = warp - 4;
        const int lane_quarter = ew & 3;  // TMEM lanes a warp may read: 32 * (warp % 4) ...
        const int col_half = ew >> 2;     // this warp's 128-column half of the 256-column accumulator
        constexpr bool kTmaOut = (EPI != EPI_PATCH && EPI != EPI_TOP4);
        constexpr bool kOut16 = (EPI == EPI_STORE || EPI == EPI_GELU || EPI == EPI_QUICKGELU);
        constexpr int PIECE_COLS = kOut16 ? 64 : 32;  // one staged piece is 32 rows x 128 bytes
        constexpr int PIECES = 128 / PIECE_COLS;
        const uint32_t stage_base = ptx::smem_u32(out_stage) + ew * C::OUT_BUF_BYTES;
        const uint32_t swz = static_cast<uint32_t>(lane & 7);  // 128-byte swizzle: 16-byte chunk index ^= row & 7
        const uint32_t row_off = static_cast<uint32_t>(lane) * 128u;
        int iter = 0;
        for (int t = cluster_id; t < num_tiles; t += num_clusters, ++iter) {
            const int acc = iter & 1;
            const uint32_t acc_phase = (iter >> 1) & 1;
            const int m0 = ((t / tiles_n) * CG + static_cast<int>(cta_rank)) * BLOCK_M;
            const int n0 = (t % tiles_n) * BLOCK_N;
            const int row0 = m0 + lane_quarter * 32;  // first row of this warp's 32-row band
            // This warp's 128 bias values, lane l holding columns l, 32+l, 64+l, 96+l of the range: fetched before the
            // accumulator is waited for (a load issued at the point of use stalls every piece on an L2 round trip) and
            // broadcast with shuffles below.
            float bias_r[4] = {0.f, 0.f, 0.f, 0.f};
            if (kTmaOut && epi.bias != nullptr) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int col = n0 + col_half * 128 + 32 * j + lane;
                    if (col < N) bias_r[j] = __ldg(epi.bias + col);
                }
            }
            // One epilogue warp polls the accumulator barrier, the other seven sleep in a hardware named barrier: eight
            // warps spinning on an mbarrier take issue slots from the warps that still have epilogue math to do.
            if (ew == 0) ptx::mbar_wait(&tmem_full_bar[acc], acc_phase);
            ptx::named_bar_sync(1, NUM_EPI_WARPS * 32);
            ptx::tcgen05_fence_after();
            const uint32_t taddr =
                tmem_base + (static_cast<uint32_t>(lane_quarter * 32) << 16) + acc * BLOCK_N + col_half * 128;
            if constexpr (kTmaOut) {
#pragma unroll 1
                for (int pc = 0; pc < PIECES; ++pc) {
                    const int col0 = n0 + col_half * 128 + pc * PIECE_COLS;
                    uint32_t r[PIECE_COLS];
                    ptx::tmem_ld_32x32b_x32(taddr + pc * PIECE_COLS, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
                    if constexpr (PIECE_COLS == 64)
                        ptx::tmem_ld_32x32b_x32(taddr + pc * PIECE_COLS + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
                    // the staging tile about to be overwritten was handed to TMA one piece ago by lane 0; its read
                    // has had this piece's TMEM load to complete
                    if (lane == 0) ptx::bulk_wait_group_read<0>();
                    ptx::tmem_ld_wait();
                    if (pc == PIECES - 1) {
                        // All of this warp's TMEM reads for the tile are complete: hand the accumulator back.
                        ptx::tcgen05_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if (CG == 1)
                                ptx::mbar_arrive(&tmem_empty_bar[acc]);
                            else
                                ptx::mbar_arrive_cluster(&tmem_empty_bar[acc], 0);
                        }
                    } else {
                        __syncwarp();
                    }
                    if (row0 < M && col0 < N) {  // warp-uniform; TMA clips the partial tails
                        const uint32_t buf = stage_base + row_off;
                        // this piece's bias registers, picked with selects so the array is never indexed dynamically
                        float bsel[2];
                        if constexpr (PIECE_COLS == 64) {
                            bsel[0] = pc == 0 ? bias_r[0] : bias_r[2];
                            bsel[1] = pc == 0 ? bias_r[1] : bias_r[3];
                        } else {
                            bsel[0] = pc == 0 ? bias_r[0] : (pc == 1 ? bias_r[1] : (pc == 2 ? bias_r[2] : bias_r[3]));
                            bsel[1] = 0.f;
                        }
#pragma unroll
                        for (int c = 0; c < PIECE_COLS / 4; ++c) {  // 4 columns at a time
                            float v[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[4 * c + i]);
                            if (epi.bias != nullptr) {  // warp-uniform
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const int cc = 4 * c + i;  // column within the piece
                                    v[i] += __shfl_sync(0xffffffffu, bsel[cc >> 5], cc & 31);
                                }
                            }
                            if constexpr (EPI == EPI_GELU) {
#pragma unroll
                                for (int i = 0; i < 4; i += 2) gelu_erf2(v[i], v[i + 1]);
                            } else if constexpr (EPI == EPI_QUICKGELU) {
#pragma unroll
                                for (int i = 0; i < 4; ++i) v[i] = quick_gelu(v[i]);
                            }
                            if constexpr (kOut16) {
                                // two consecutive 4-column groups make one 16-byte chunk: keep the first half
                                r[4 * c + 0] = pack2<T>(v[0], v[1]);
                                r[4 * c + 1] = pack2<T>(v[2], v[3]);
                                if (c & 1) {
                                    const uint32_t chunk = static_cast<uint32_t>(c >> 1);
                                    ptx::st_shared_v4(buf + ((chunk ^ swz) << 4), r[4 * (c - 1)], r[4 * (c - 1) + 1],
                                                      r[4 * c], r[4 * c + 1]);
                                }
                            } else {
                                ptx::st_shared_v4(buf + ((static_cast<uint32_t>(c) ^ swz) << 4), __float_as_uint(v[0]),
                                                  __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
                            }
                        }
                        ptx::fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            const void* src = out_stage + ew * C::OUT_BUF_BYTES;
                            if constexpr (EPI == EPI_RESID)
                                ptx::tma_reduce_add_2d(&map_out, src, col0, row0);
                            else
                                ptx::tma_store_2d(&map_out, src, col0, row0);
                        }
                    }
                    if (lane == 0) ptx::bulk_commit_group();  // one group per piece, empty or not, keeps the count in step
                }
            } else {
                const int row = row0 + lane;
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    uint32_t r[32];
                    ptx::tmem_ld_32x32b_x32(taddr + c * 32, r);
                    ptx::tmem_ld_wait();
                    if (c == 3) {
                        ptx::tcgen05_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if (CG == 1)
                                ptx::mbar_arrive(&tmem_empty_bar[acc]);
                            else
                                ptx::mbar_arrive_cluster(&tmem_empty_bar[acc], 0);
                        }
                    }
                    const int col0 = n0 + col_half * 128 + c * 32;
                    if (row < M && col0 < N) {
                        float v[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
                        epilogue_chunk<T, EPI>(v, epi, row, col0, N);
                    }
                }
            }
        }
        if constexpr (kTmaOut) {
            if (lane == 0) ptx::bulk_wait_group<0>();  // every store / reduce of this warp has been performed
        }
    }

    // Teardown: every role is done (the epilogue's last wait implies all MMAs and TMA loads retired).
    ptx::tcgen05_fence_before();
    if constexpr (CG == 2)
        ptx::cluster_sync();
    else
        __syncthreads();
    if (warp == 2) ptx::tmem_dealloc<CG>(tmem_base, TMEM_COLS);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        // Resolved through the runtime so the library never links libcuda.so directly.
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// Output map for the epilogue's TMA stores: dims {N (inner), M}, box {128 bytes of columns, 32 rows}, 128-byte swizzle.
int encode_output_map(CUtensorMap* out, CUtensorMapDataType cdt, int elt_bytes, void* ptr, uint64_t N, uint64_t M,
                      uint64_t ld_elems) {
    EncodeTiledFn fn = get_encode_fn();
    if (fn == nullptr) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return 1;
    }
    const cuuint64_t dims[2] = {N, M};
    const cuuint64_t strides[1] = {ld_elems * elt_bytes};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / elt_bytes), 32};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, cdt, 2, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (output) failed with CUresult %d (ptr=%p N=%llu M=%llu ld=%llu)", static_cast<int>(r),
                  ptr, (unsigned long long)N, (unsigned long long)M, (unsigned long long)ld_elems);
        return 1;
    }
    return 0;
}

// 2-D K-major operand map: dims {K (inner), rows}, box {64, box_rows}, 128-byte swizzle, zero OOB fill.
int encode_operand_map(CUtensorMap* out, DType dt, const void* ptr, uint64_t K, uint64_t rows,
                       uint64_t ld_elems, uint32_t box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (fn == nullptr) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return 1;
    }
    const cuuint64_t dims[2] = {K, rows};
    const cuuint64_t strides[1] = {ld_elems * 2};
    const cuuint32_t box[2] = {BLOCK_K, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUtensorMapDataType cdt = (dt == DT_BF16) ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    CUresult r = fn(out, cdt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (ptr=%p K=%llu rows=%llu ld=%llu box_rows=%u)",
                  static_cast<int>(r), ptr, (unsigned long long)K, (unsigned long long)rows,
                  (unsigned long long)ld_elems, box_rows);
        return 1;
    }
    return 0;
}

template <typename T, int EPI, int CG>
int launch(const GemmProblem& p, cudaStream_t stream) {
    using C = Cfg<CG>;
    auto kern = gemm_tcgen05_kernel<T, EPI, CG>;
    // per launch (cheap): the attribute belongs to the current device's copy of the function
    VIDIL_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    const int tiles_m = (p.M + BLOCK_M * CG - 1) / (BLOCK_M * CG);
    const int tiles_n = (p.N + BLOCK_N - 1) / BLOCK_N;
    const int num_tiles = tiles_m * tiles_n;
    int clusters = gemm_num_sms() / CG;
    if (clusters > num_tiles) clusters = num_tiles;
    if (clusters < 1) return 0;

    EpiArgs e{p.bias, p.out, p.ldo, p.pos, p.patches_per_frame};
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(clusters * CG);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    VIDIL_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, p.map_a, p.map_w, p.map_out, p.M, p.N, p.K, e));
    count_launches(1);
    return 0;
}

template <typename T, int CG>
int dispatch_epi(const GemmProblem& p, cudaStream_t s) {
    switch (p.epi) {
        case EPI_STORE: return launch<T, EPI_STORE, CG>(p, s);
        case EPI_GELU: return launch<T, EPI_GELU, CG>(p, s);
        case EPI_QUICKGELU: return launch<T, EPI_QUICKGELU, CG>(p, s);
        case EPI_RESID: return launch<T, EPI_RESID, CG>(p, s);
        case EPI_PATCH: return launch<T, EPI_PATCH, CG>(p, s);
        case EPI_STORE_F32: return launch<T, EPI_STORE_F32, CG>(p, s);
        case EPI_TOP4: return launch<T, EPI_TOP4, CG>(p, s);
        default: set_error("gemm: unknown epilogue mode %d", p.epi); return 1;
    }
}

}  // namespace

int gemm_num_sms() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 0;
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 0;
        if (prop.major != 10) {
            set_error("vidil_b200 kernels are built for sm_100a only; device is sm_%d%d", prop.major, prop.minor);
            return 0;
        }
        sms = prop.multiProcessorCount;
    }
    return sms;
}

int gemm_prepare(GemmProblem& p) {
    if (p.M <= 0 || p.N <= 0 || p.K <= 0) {
        set_error("gemm: empty problem M=%d N=%d K=%d", p.M, p.N, p.K);
        return 1;
    }
    if (p.K % BLOCK_K != 0) {
        set_error("gemm: K=%d must be a multiple of %d", p.K, BLOCK_K);
        return 1;
    }
    if (p.lda % 8 != 0 || p.ldw % 8 != 0 || (reinterpret_cast<uintptr_t>(p.A) & 15) ||
        (reinterpret_cast<uintptr_t>(p.W) & 15)) {
        set_error("gemm: operands must be 16-byte aligned with leading dimensions that are multiples of 8");
        return 1;
    }
    const bool out_f32 = (p.epi == EPI_RESID || p.epi == EPI_PATCH || p.epi == EPI_STORE_F32 || p.epi == EPI_TOP4);
    if (p.ldo % (out_f32 ? 4 : 8) != 0 || (reinterpret_cast<uintptr_t>(p.out) & 15)) {
        set_error("gemm: output must be 16-byte aligned with a 16-byte multiple leading dimension");
        return 1;
    }
    if (p.bias != nullptr && (reinterpret_cast<uintptr_t>(p.bias) & 15)) {
        set_error("gemm: bias must be 16-byte aligned");
        return 1;
    }
    if (p.epi == EPI_PATCH && (p.pos == nullptr || p.patches_per_frame <= 0)) {
        set_error("gemm: EPI_PATCH needs a position table and patches_per_frame");
        return 1;
    }
    if (p.cta_group != 1 && p.cta_group != 2) {
        set_error("gemm: cta_group must be 1 or 2");
        return 1;
    }
    if (encode_operand_map(&p.map_a, p.dt, p.A, p.K, p.M, p.lda, BLOCK_M)) return 1;
    if (encode_operand_map(&p.map_w, p.dt, p.W, p.K, p.N, p.ldw, BLOCK_N / p.cta_group)) return 1;
    if (p.epi == EPI_TOP4 && p.ldo < 4 * ((p.N + 31) / 32)) {
        set_error("gemm: EPI_TOP4 needs ldo >= 4 * ceil(N / 32) floats per row");
        return 1;
    }
    if (p.epi != EPI_PATCH && p.epi != EPI_TOP4) {
        const CUtensorMapDataType odt = out_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                                : (p.dt == DT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
        if (encode_output_map(&p.map_out, odt, out_f32 ? 4 : 2, p.out, p.N, p.M, p.ldo)) return 1;
    } else {
        p.map_out = p.map_a;  // unused by the scatter epilogue; any valid descriptor
    }
    p.prepared = true;
    return 0;
}

int gemm_run(const GemmProblem& p, cudaStream_t stream) {
    if (!p.prepared) {
        set_error("gemm_run called on an unprepared problem");
        return 1;
    }
    if (gemm_num_sms() == 0) return 1;
    if (p.dt == DT_BF16)
        return p.cta_group == 2 ? dispatch_epi<__nv_bfloat16, 2>(p, stream) : dispatch_epi<__nv_bfloat16, 1>(p, stream);
    return p.cta_group == 2 ? dispatch_epi<__half, 2>(p, stream) : dispatch_epi<__half, 1>(p, stream);
}

}  // namespace vidil
