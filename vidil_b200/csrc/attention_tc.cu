// Fused softmax attention on tcgen05 for sequences that fit one key tile (N <= 208 tokens, head_dim 64) — the
// ViT-L/16 @224 shape (N = 197).  Reference: models/vit.py:72-83 (Attention.forward); the reference materialises
// [B,H,N,N] scores in HBM, here S and P never leave the SM.
//
// One persistent CTA per SM walks (frame, head) items.  Per item the queries form up to two 128-row tiles t:
//   S_t = Q_t K^T        tcgen05.mma M=128, N=ceil16(keys), K=64;   fp32 S_t in TMEM columns [256t, 256t+N)
//   P_t = exp2(c S_t - c max)   one thread per query row reads its S row from TMEM twice (max, then exp/sum),
//                               writes bf16/fp16 P_t to shared memory in the no-swizzle K-major UMMA layout
//   O_t = P_t V          tcgen05.mma M=128, N=64, K=keys; V is consumed straight from its TMA tile as an MN-major
//                        B operand (no transpose anywhere); fp32 O_t in TMEM over S_t's first 64 columns
//   out = O_t / rowsum   -> 16-bit -> swizzled staging -> 3-D TMA store (rows past the frame's last token clipped)
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer (+ TMEM allocation), warps 2-5 and 6-9 two softmax/epilogue
// groups.  Group g owns tile (g + item) & 1, so the 128-row and the 69-row tile alternate between the groups and while
// one group runs its exponentials the tensor core works for the other.
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
constexpr int MAX_KEYS = 208;    // 13 x 16: S_t (fp32) fits 256 TMEM columns, K/V tiles fit one TMA box
constexpr int ATC_THREADS = 320;

constexpr int Q_BYTES = QT * HD * 2;               // 16 KB
constexpr int KV_BYTES = MAX_KEYS * HD * 2;        // 26 KB
constexpr int P_BYTES = (MAX_KEYS / 8) * QT * 16;  // 26 chunks of [128 rows][8 keys]: 52 KB
constexpr int OFF_Q = 0;
constexpr int OFF_K = 2 * Q_BYTES;
constexpr int OFF_V = OFF_K + KV_BYTES;
constexpr int OFF_P = OFF_V + KV_BYTES;
constexpr int OFF_OUT = OFF_P + 2 * P_BYTES;       // 8 warps x [32 rows][128 B]
constexpr int OFF_BAR = OFF_OUT + 8 * 4096;
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

template <typename T>
__global__ void __launch_bounds__(ATC_THREADS, 1)
    attention_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv,
                        const __grid_constant__ CUtensorMap map_out, int n_items, int N, int H, float scale_log2e) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    const uint32_t sbase = ptx::smem_u32(smem);

    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* qk_full = bars + 0;
    uint64_t* v_full = bars + 1;
    uint64_t* qk_empty = bars + 2;
    uint64_t* v_empty = bars + 3;
    uint64_t* s_full = bars + 4;   // [2]
    uint64_t* p_full = bars + 6;   // [2]
    uint64_t* o_full = bars + 8;   // [2]
    uint64_t* s_free = bars + 10;  // [2]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 12);

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
        ptx::mbar_init(qk_empty, 1);
        ptx::mbar_init(v_empty, 1);
        for (int t = 0; t < 2; ++t) {
            ptx::mbar_init(&s_full[t], 1);
            ptx::mbar_init(&p_full[t], 4);  // lane 0 of each of the owning group's 4 warps
            ptx::mbar_init(&o_full[t], 1);
            ptx::mbar_init(&s_free[t], 4);
        }
        ptx::fence_mbar_init();
    }
    if (warp == 1) ptx::tmem_alloc<1>(tmem_ptr_smem, 512);
    ptx::tcgen05_fence_before();
    __syncthreads();
    ptx::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        if (lane == 0) {
            // ===================== TMA producer =====================
            int it = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const uint32_t ph = it & 1;
                const int b = item / H, h = item - b * H;
                const int row0 = b * N;
                ptx::mbar_wait(qk_empty, ph ^ 1);  // previous item's S MMAs have consumed Q and K
                ptx::mbar_arrive_expect_tx(qk_full, n_tiles * Q_BYTES + nk16 * HD * 2);
                ptx::tma_load_2d(&map_q, qk_full, smem + OFF_Q, h * HD, row0);
                if (n_tiles == 2) ptx::tma_load_2d(&map_q, qk_full, smem + OFF_Q + Q_BYTES, h * HD, row0 + QT);
                ptx::tma_load_2d(&map_kv, qk_full, smem + OFF_K, D + h * HD, row0);
                ptx::mbar_wait(v_empty, ph ^ 1);  // previous item's PV MMAs have consumed V
                ptx::mbar_arrive_expect_tx(v_full, nk16 * HD * 2);
                ptx::tma_load_2d(&map_kv, v_full, smem + OFF_V, 2 * D + h * HD, row0);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            // ===================== MMA issuer =====================
            constexpr bool kIsBf16 = std::is_same<T, __nv_bfloat16>::value;
            const uint32_t idesc_s = ptx::make_idesc_f16(kIsBf16, QT, nk16);
            const uint32_t idesc_o = ptx::make_idesc_f16_bmn(kIsBf16, QT, HD);
            const int ksteps = nk16 / 16;
            int it = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const uint32_t ph = it & 1;
                ptx::mbar_wait(qk_full, ph);
                ptx::tcgen05_fence_after();
                const uint64_t dk = ptx::make_kmajor_sw128_desc(sbase + OFF_K);
                for (int t = 0; t < n_tiles; ++t) {
                    ptx::mbar_wait(&s_free[t], ph ^ 1);  // last item's O_t (aliasing S_t) has been read out
                    ptx::tcgen05_fence_after();
                    const uint64_t dq = ptx::make_kmajor_sw128_desc(sbase + OFF_Q + t * Q_BYTES);
#pragma unroll
                    for (int k = 0; k < HD / 16; ++k)
                        ptx::umma_f16<1>(tmem_base + t * 256, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
                    ptx::umma_commit<1>(&s_full[t]);
                }
                ptx::umma_commit<1>(qk_empty);
                ptx::mbar_wait(v_full, ph);
                for (int t = 0; t < n_tiles; ++t) {
                    ptx::mbar_wait(&p_full[t], ph);  // P_t is in shared memory, S_t has been consumed
                    ptx::tcgen05_fence_after();
                    const uint32_t sp = sbase + OFF_P + t * P_BYTES;
                    for (int j = 0; j < ksteps; ++j) {
                        // A: P_t k-step j = two [128 rows][8 keys] chunks 2048 B apart, 8-row groups 128 B apart
                        const uint64_t da = ptx::make_smem_desc(sp + j * 4096, 2048, 128, 0);
                        // B: V rows 16j..16j+15 (two 8-row 1024-byte swizzle groups), 64 contiguous channels per row
                        const uint64_t db = ptx::make_smem_desc(sbase + OFF_V + j * 2048, 0, 1024, 2);
                        ptx::umma_f16<1>(tmem_base + t * 256, da, db, idesc_o, j != 0);
                    }
                    ptx::umma_commit<1>(&o_full[t]);
                }
                ptx::umma_commit<1>(v_empty);
            }
        }
        __syncwarp();
    } else {
        // ===================== softmax + output groups =====================
        const int g = (warp - 2) >> 2;
        const int quarter = warp & 3;  // TMEM lanes this warp may touch: 32 * (warp % 4) ...
        const uint32_t stage = sbase + OFF_OUT + (warp - 2) * 4096;
        const void* stage_ptr = smem + OFF_OUT + (warp - 2) * 4096;
        int it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const int t = (g + it) & 1;
            if (t >= n_tiles) continue;
            const uint32_t ph = it & 1;
            const int b = item / H, h = item - b * H;
            const int row_in_tile = quarter * 32 + lane;
            const bool warp_valid = (t * QT + quarter * 32) < N;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + t * 256;
            float inv_sum = 0.f;

            ptx::mbar_wait(&s_full[t], ph);
            ptx::tcgen05_fence_after();
            if (warp_valid) {
                // pass 1: row maximum over the valid keys
                float mx = -INFINITY;
                for (int c0 = 0; c0 < nk16; c0 += 32) {
                    uint32_t r[32];
                    if (c0 + 32 <= nk16) {
                        ptx::tmem_ld_32x32b_x32(taddr + c0, r);
                    } else {
                        ptx::tmem_ld_32x32b_x16(taddr + c0, *reinterpret_cast<uint32_t(*)[16]>(&r[0]));
#pragma unroll
                        for (int i = 16; i < 32; ++i) r[i] = 0xff800000u;  // -inf
                    }
                    ptx::tmem_ld_wait();
                    if (c0 + 32 <= N) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, (c0 + i < N) ? __uint_as_float(r[i]) : -INFINITY);
                    }
                }
                const float mxs = mx * scale_log2e;
                // pass 2: p = exp2(c s - c max), row sum, 16-bit P into the UMMA A layout
                float sum = 0.f;
                const uint32_t prow = sbase + OFF_P + t * P_BYTES + row_in_tile * 16;
                for (int c0 = 0; c0 < nk16; c0 += 32) {
                    uint32_t r[32];
                    const bool wide = (c0 + 32 <= nk16);
                    if (wide)
                        ptx::tmem_ld_32x32b_x32(taddr + c0, r);
                    else
                        ptx::tmem_ld_32x32b_x16(taddr + c0, *reinterpret_cast<uint32_t(*)[16]>(&r[0]));
                    ptx::tmem_ld_wait();
                    const bool nomask = (c0 + 32 <= N);
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {  // 8 keys = one 16-byte chunk
                        if (q4 < 2 || wide) {
                            float p[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float e = ptx::ex2_approx(fmaf(__uint_as_float(r[8 * q4 + i]), scale_log2e, -mxs));
                                p[i] = (nomask || c0 + 8 * q4 + i < N) ? e : 0.f;
                                sum += p[i];
                            }
                            ptx::st_shared_v4(prow + ((c0 >> 3) + q4) * 2048, pack2<T>(p[0], p[1]), pack2<T>(p[2], p[3]),
                                              pack2<T>(p[4], p[5]), pack2<T>(p[6], p[7]));
                        }
                    }
                }
                inv_sum = 1.0f / sum;
            }
            ptx::fence_proxy_async_smem();  // P (generic-proxy stores) before the MMA's async-proxy reads
            ptx::tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&p_full[t]);

            ptx::mbar_wait(&o_full[t], ph);
            ptx::tcgen05_fence_after();
            if (warp_valid) {
                uint32_t r[64];
                ptx::tmem_ld_32x32b_x32(taddr, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
                ptx::tmem_ld_32x32b_x32(taddr + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
                if (lane == 0) ptx::bulk_wait_group_read<0>();  // this warp's previous output store has left the staging tile
                ptx::tmem_ld_wait();
                __syncwarp();
                const uint32_t srow = stage + lane * 128;
                const uint32_t swz = static_cast<uint32_t>(lane & 7);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    uint32_t u[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        u[i] = pack2<T>(__uint_as_float(r[8 * c + 2 * i]) * inv_sum, __uint_as_float(r[8 * c + 2 * i + 1]) * inv_sum);
                    ptx::st_shared_v4(srow + ((static_cast<uint32_t>(c) ^ swz) << 4), u[0], u[1], u[2], u[3]);
                }
                ptx::fence_proxy_async_smem();
            }
            ptx::tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
                ptx::mbar_arrive(&s_free[t]);  // S_t / O_t columns may be overwritten by the next item's S_t
                if (warp_valid) {
                    ptx::tma_store_3d(&map_out, stage_ptr, h * HD, t * QT + quarter * 32, b);
                    ptx::bulk_commit_group();
                }
            }
        }
        if (lane == 0) ptx::bulk_wait_group<0>();
    }

    ptx::tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc<1>(tmem_base, 512);
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
    kern<<<grid, ATC_THREADS, ATC_SMEM, stream>>>(m.q, m.kv, m.out, n_items, N, H, scale_log2e);
    VIDIL_CUDA_OK(cudaGetLastError());
    count_launches(1);
    return 0;
}

}  // namespace

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
