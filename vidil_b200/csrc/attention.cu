// Fused scaled-dot-product attention for short ViT sequences, head_dim 64:
//   out[b, n, h*64 + d] = softmax_j( q[b,h,n,:] . k[b,h,j,:] * scale ) @ v[b,h,j,d]
// Reference: models/vit.py:72-83 (Attention.forward) — the reference materialises the [B,H,N,N] score
// tensor (636 MB fp32 at B=256, L/16); here scores never leave the SM.
//
// q/k/v are read in place from the fused-QKV GEMM output, layout [B, N, 3, H, 64] (the reshape at
// vit.py:72 without the permute copy), and the result is written head-merged ([B, N, H*64], the
// transpose(1,2).reshape at vit.py:83) ready to be the proj GEMM's A operand.
//
// Round-1 kernel: one CTA = 64 queries of one (frame, head); 4 warps x 16 query rows; keys/values
// streamed in 64-row blocks through a cp.async double buffer; QK^T and PV on mma.sync m16n8k16
// (fp16/bf16 in, fp32 accumulate) fed by ldmatrix from XOR-swizzled shared memory; online softmax in
// fp32 with exp2 on pre-scaled logits.  Per layer the kernel is HBM-bound on reading qkv once and
// writing out once (98 FLOP/B at N=197), so the legacy tensor path is not the limiter yet; a tcgen05
// version (S and P resident in TMEM) is the planned replacement.
#include <math.h>

#include "kernels.h"
#include "ptx.cuh"

namespace vidil {
namespace {

constexpr int HD = 64;       // head dim
constexpr int BQ = 64;       // queries per CTA
constexpr int BKV = 64;      // keys per pipeline stage
constexpr int ATT_THREADS = 128;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;  // src-size 0 => the 16 destination bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}

template <typename T>
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1);
template <>
__device__ __forceinline__ void mma16816<__half>(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <>
__device__ __forceinline__ void mma16816<__nv_bfloat16>(float (&c)[4], const uint32_t (&a)[4], uint32_t b0,
                                                        uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
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

// Byte offset of 16-byte chunk `chunk` (0..7) of row `row` in a [rows][64 x 16-bit] tile whose chunks are
// XOR-swizzled by the low row bits so ldmatrix's 8 row addresses hit 8 different bank groups.
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) {
    return static_cast<uint32_t>(row * 128 + ((chunk ^ (row & 7)) << 4));
}

// Copy a [ROWS x 64] tile (rows row0..row0+ROWS-1 of a strided matrix, zero beyond `nrows`) into shared memory.
template <typename T, int THREADS = ATT_THREADS, int ROWS = BKV>
__device__ __forceinline__ void load_tile_async(uint32_t smem_tile, const T* g, int64_t row_stride, int row0,
                                                int nrows) {
#pragma unroll
    for (int i = 0; i < (ROWS * 8) / THREADS; ++i) {
        const int idx = threadIdx.x + i * THREADS;
        const int r = idx >> 3, c = idx & 7;
        const int gr = row0 + r;
        const bool ok = gr < nrows;
        const T* src = g + static_cast<int64_t>(ok ? gr : 0) * row_stride + c * 8;
        cp_async16(smem_tile + tile_off(r, c), src, ok);
    }
}

template <typename T>
__global__ void __launch_bounds__(ATT_THREADS)
    attention_kernel(const T* __restrict__ qkv, T* __restrict__ out, int N, int H, float scale_log2e) {
    __shared__ __align__(128) uint8_t smem[BQ * HD * 2 + 2 * 2 * BKV * HD * 2];  // Q | K0 V0 | K1 V1 : 40 KB
    const uint32_t sQ = ptx::smem_u32(smem);
    const uint32_t sKV = sQ + BQ * HD * 2;  // stage s: K at sKV + s*16384, V at +8192

    const int qblk = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t tok_stride = static_cast<int64_t>(3) * H * HD;
    const T* q_base = qkv + static_cast<int64_t>(b) * N * tok_stride + h * HD;
    const T* k_base = q_base + static_cast<int64_t>(H) * HD;
    const T* v_base = k_base + static_cast<int64_t>(H) * HD;
    const int q0 = qblk * BQ;
    const int num_kb = (N + BKV - 1) / BKV;

    load_tile_async<T>(sQ, q_base, tok_stride, q0, N);
    load_tile_async<T>(sKV, k_base, tok_stride, 0, N);
    load_tile_async<T>(sKV + 8192, v_base, tok_stride, 0, N);
    cp_async_commit();

    uint32_t qf[4][4];  // Q fragments for the 4 k-steps over head_dim
    float o[8][4];      // output accumulator: 16 rows x 64 dims per warp
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m_run[2] = {-INFINITY, -INFINITY};  // running max of scaled logits (rows g and g+8)
    float l_run[2] = {0.f, 0.f};              // running sum of exp2

    for (int kb = 0; kb < num_kb; ++kb) {
        const int stage = kb & 1;
        if (kb + 1 < num_kb) {
            const uint32_t nxt = sKV + (stage ^ 1) * 16384;
            load_tile_async<T>(nxt, k_base, tok_stride, (kb + 1) * BKV, N);
            load_tile_async<T>(nxt + 8192, v_base, tok_stride, (kb + 1) * BKV, N);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();

        if (kb == 0) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int c = kk * 2 + (lane >> 4);
                ldsm_x4(sQ + tile_off(r, c), qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3]);
            }
        }

        const uint32_t sK = sKV + stage * 16384;
        const uint32_t sV = sK + 8192;

        // S = Q K^T : 16 x 64 per warp
        float s[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
            for (int kk2 = 0; kk2 < 2; ++kk2) {
                uint32_t b0, b1, b2, b3;
                ldsm_x4(sK + tile_off(nt * 8 + (lane & 7), kk2 * 4 + (lane >> 3)), b0, b1, b2, b3);
                mma16816<T>(s[nt], qf[2 * kk2], b0, b1);
                mma16816<T>(s[nt], qf[2 * kk2 + 1], b2, b3);
            }
        }

        // scale into the log2 domain, mask keys beyond N in the last block
        const int key0 = kb * BKV + (lane & 3) * 2;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const int key = key0 + nt * 8;
            s[nt][0] = (key < N) ? s[nt][0] * scale_log2e : -INFINITY;
            s[nt][1] = (key + 1 < N) ? s[nt][1] * scale_log2e : -INFINITY;
            s[nt][2] = (key < N) ? s[nt][2] * scale_log2e : -INFINITY;
            s[nt][3] = (key + 1 < N) ? s[nt][3] * scale_log2e : -INFINITY;
        }

        // online softmax (rows g = lane/4 and g+8; a row lives in the 4 lanes of a quad)
        float mx[2] = {m_run[0], m_run[1]};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
            mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 1));
            mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 2));
        }
        // every key block holds at least one valid key (kb*64 < N), so mx is finite here
        const float corr0 = exp2f(m_run[0] - mx[0]);
        const float corr1 = exp2f(m_run[1] - mx[1]);
        m_run[0] = mx[0];
        m_run[1] = mx[1];
        float rs[2] = {0.f, 0.f};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            s[nt][0] = exp2f(s[nt][0] - mx[0]);
            s[nt][1] = exp2f(s[nt][1] - mx[0]);
            s[nt][2] = exp2f(s[nt][2] - mx[1]);
            s[nt][3] = exp2f(s[nt][3] - mx[1]);
            rs[0] += s[nt][0] + s[nt][1];
            rs[1] += s[nt][2] + s[nt][3];
        }
        l_run[0] = l_run[0] * corr0 + rs[0];
        l_run[1] = l_run[1] * corr1 + rs[1];
#pragma unroll
        for (int nd = 0; nd < 8; ++nd) {
            o[nd][0] *= corr0;
            o[nd][1] *= corr0;
            o[nd][2] *= corr1;
            o[nd][3] *= corr1;
        }

        // O += P V : P re-used straight from the S accumulator registers as the A operand
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t pa[4];
            pa[0] = pack2<T>(s[2 * j][0], s[2 * j][1]);
            pa[1] = pack2<T>(s[2 * j][2], s[2 * j][3]);
            pa[2] = pack2<T>(s[2 * j + 1][0], s[2 * j + 1][1]);
            pa[3] = pack2<T>(s[2 * j + 1][2], s[2 * j + 1][3]);
#pragma unroll
            for (int nd2 = 0; nd2 < 4; ++nd2) {
                uint32_t b0, b1, b2, b3;
                const int r = j * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int c = nd2 * 2 + (lane >> 4);
                ldsm_x4_trans(sV + tile_off(r, c), b0, b1, b2, b3);
                mma16816<T>(o[2 * nd2], pa, b0, b1);
                mma16816<T>(o[2 * nd2 + 1], pa, b2, b3);
            }
        }
        __syncthreads();  // all warps are done with this stage before it is refilled
    }

    // finalise: full row sums across the quad, normalise, stage through this warp's rows of the Q tile
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        l_run[i] += __shfl_xor_sync(0xffffffffu, l_run[i], 1);
        l_run[i] += __shfl_xor_sync(0xffffffffu, l_run[i], 2);
    }
    const float inv0 = 1.0f / l_run[0], inv1 = 1.0f / l_run[1];
    const int g = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int nd = 0; nd < 8; ++nd) {
        // element columns nd*8 + tq*2 + {0,1}: chunk nd, byte offset tq*4 inside the chunk
        const int r0 = warp * 16 + g, r1 = r0 + 8;
        const uint32_t v0 = pack2<T>(o[nd][0] * inv0, o[nd][1] * inv0);
        const uint32_t v1 = pack2<T>(o[nd][2] * inv1, o[nd][3] * inv1);
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(sQ + tile_off(r0, nd) + tq * 4), "r"(v0) : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(sQ + tile_off(r1, nd) + tq * 4), "r"(v1) : "memory");
    }
    __syncwarp();
    T* o_base = out + static_cast<int64_t>(b) * N * (H * HD) + h * HD;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int idx = lane + i * 32;  // 16 rows x 8 chunks
        const int r = warp * 16 + (idx >> 3), c = idx & 7;
        const int n = q0 + r;
        if (n < N) {
            uint4 v;
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                         : "r"(sQ + tile_off(r, c)));
            *reinterpret_cast<uint4*>(o_base + static_cast<int64_t>(n) * (H * HD) + c * 8) = v;
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// The same mma.sync pipeline with separate query and key/value sequences — the attentions of the med.py text stack
// (models/med.py:146-232): self-attention over short text sequences read in place from the fused projection output
// (padding mask, blip_itm.py:49-51, or causal mask, med.py:636), and cross-attention of a group's nq text rows onto the nk
// image tokens of its frame, whose K/V were projected once per frame (med.py:160-163).  grid = (ceil(nq/64), H, groups).
// Masked keys get -10000 added to the scaled score like the reference (med.py:667); warps whose 16 query rows lie beyond
// nq (decode: 3 beams per frame) only help with the loads.
// ---------------------------------------------------------------------------------------------------------------------
template <typename T, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32)
    attention_x_kernel(const T* __restrict__ q, int64_t q_stride, const T* __restrict__ k, const T* __restrict__ v,
                       int64_t kv_stride, const int32_t* __restrict__ frame_of_group, const int32_t* __restrict__ key_mask,
                       T* __restrict__ out, int64_t out_stride, int nq, int nk, int causal, float scale_log2e) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    // Q tile (16 rows per warp) | K/V stage 0 | K/V stage 1 (only when there is more than one key block)
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int XBQ = NWARPS * 16, XT = NWARPS * 32;
    const uint32_t sQ = ptx::smem_u32(smem);
    const uint32_t sKV = sQ + XBQ * HD * 2;
    constexpr float MASKV = -10000.0f * 1.4426950408889634f;

    const int qblk = blockIdx.x, h = blockIdx.y, g = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = frame_of_group ? frame_of_group[g] : g;
    const T* q_base = q + static_cast<int64_t>(g) * nq * q_stride + h * HD;
    const T* k_base = k + static_cast<int64_t>(f) * nk * kv_stride + h * HD;
    const T* v_base = v + static_cast<int64_t>(f) * nk * kv_stride + h * HD;
    const int32_t* mrow = key_mask ? key_mask + static_cast<int64_t>(g) * nk : nullptr;
    const int q0 = qblk * XBQ;
    const int num_kb = (nk + BKV - 1) / BKV;
    const bool active = q0 + warp * 16 < nq;

    load_tile_async<T, XT, XBQ>(sQ, q_base, q_stride, q0, nq);
    load_tile_async<T, XT>(sKV, k_base, kv_stride, 0, nk);
    load_tile_async<T, XT>(sKV + 8192, v_base, kv_stride, 0, nk);
    cp_async_commit();

    uint32_t qf[4][4];
    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m_run[2] = {-INFINITY, -INFINITY};
    float l_run[2] = {0.f, 0.f};
    const int qrow0 = q0 + warp * 16 + (lane >> 2), qrow1 = qrow0 + 8;

    for (int kb = 0; kb < num_kb; ++kb) {
        const int stage = kb & 1;
        if (kb + 1 < num_kb) {
            const uint32_t nxt = sKV + (stage ^ 1) * 16384;
            load_tile_async<T, XT>(nxt, k_base, kv_stride, (kb + 1) * BKV, nk);
            load_tile_async<T, XT>(nxt + 8192, v_base, kv_stride, (kb + 1) * BKV, nk);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();

        if (active) {
            if (kb == 0) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                    const int c = kk * 2 + (lane >> 4);
                    ldsm_x4(sQ + tile_off(r, c), qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3]);
                }
            }
            const uint32_t sK = sKV + stage * 16384;
            const uint32_t sV = sK + 8192;
            float s[8][4];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
                for (int kk2 = 0; kk2 < 2; ++kk2) {
                    uint32_t b0, b1, b2, b3;
                    ldsm_x4(sK + tile_off(nt * 8 + (lane & 7), kk2 * 4 + (lane >> 3)), b0, b1, b2, b3);
                    mma16816<T>(s[nt], qf[2 * kk2], b0, b1);
                    mma16816<T>(s[nt], qf[2 * kk2 + 1], b2, b3);
                }
            }
            const int key0 = kb * BKV + (lane & 3) * 2;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int key = key0 + nt * 8 + (e & 1);
                    const int row = (e < 2) ? qrow0 : qrow1;
                    float x = s[nt][e] * scale_log2e;
                    if (key >= nk) {
                        x = -INFINITY;
                    } else {
                        if (mrow != nullptr && mrow[key] == 0) x += MASKV;
                        if (causal && key > row) x += MASKV;
                    }
                    s[nt][e] = x;
                }
            }
            float mx[2] = {m_run[0], m_run[1]};
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
                mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 1));
                mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 2));
            }
            const float corr0 = exp2f(m_run[0] - mx[0]);
            const float corr1 = exp2f(m_run[1] - mx[1]);
            m_run[0] = mx[0];
            m_run[1] = mx[1];
            float rs[2] = {0.f, 0.f};
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                s[nt][0] = exp2f(s[nt][0] - mx[0]);
                s[nt][1] = exp2f(s[nt][1] - mx[0]);
                s[nt][2] = exp2f(s[nt][2] - mx[1]);
                s[nt][3] = exp2f(s[nt][3] - mx[1]);
                rs[0] += s[nt][0] + s[nt][1];
                rs[1] += s[nt][2] + s[nt][3];
            }
            l_run[0] = l_run[0] * corr0 + rs[0];
            l_run[1] = l_run[1] * corr1 + rs[1];
#pragma unroll
            for (int nd = 0; nd < 8; ++nd) {
                o[nd][0] *= corr0;
                o[nd][1] *= corr0;
                o[nd][2] *= corr1;
                o[nd][3] *= corr1;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t pa[4];
                pa[0] = pack2<T>(s[2 * j][0], s[2 * j][1]);
                pa[1] = pack2<T>(s[2 * j][2], s[2 * j][3]);
                pa[2] = pack2<T>(s[2 * j + 1][0], s[2 * j + 1][1]);
                pa[3] = pack2<T>(s[2 * j + 1][2], s[2 * j + 1][3]);
#pragma unroll
                for (int nd2 = 0; nd2 < 4; ++nd2) {
                    uint32_t b0, b1, b2, b3;
                    const int r = j * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                    const int c = nd2 * 2 + (lane >> 4);
                    ldsm_x4_trans(sV + tile_off(r, c), b0, b1, b2, b3);
                    mma16816<T>(o[2 * nd2], pa, b0, b1);
                    mma16816<T>(o[2 * nd2 + 1], pa, b2, b3);
                }
            }
        }
        __syncthreads();
    }
    if (!active) return;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        l_run[i] += __shfl_xor_sync(0xffffffffu, l_run[i], 1);
        l_run[i] += __shfl_xor_sync(0xffffffffu, l_run[i], 2);
    }
    const float inv0 = 1.0f / l_run[0], inv1 = 1.0f / l_run[1];
    const int gq = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int nd = 0; nd < 8; ++nd) {
        const int r0 = warp * 16 + gq, r1 = r0 + 8;
        const uint32_t v0 = pack2<T>(o[nd][0] * inv0, o[nd][1] * inv0);
        const uint32_t v1 = pack2<T>(o[nd][2] * inv1, o[nd][3] * inv1);
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(sQ + tile_off(r0, nd) + tq * 4), "r"(v0) : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(sQ + tile_off(r1, nd) + tq * 4), "r"(v1) : "memory");
    }
    __syncwarp();
    T* o_base = out + static_cast<int64_t>(g) * nq * out_stride + h * HD;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int idx = lane + i * 32;
        const int r = warp * 16 + (idx >> 3), c = idx & 7;
        const int n = q0 + r;
        if (n < nq) {
            uint4 val;
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w)
                         : "r"(sQ + tile_off(r, c)));
            *reinterpret_cast<uint4*>(o_base + static_cast<int64_t>(n) * out_stride + c * 8) = val;
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// Decode-step cross-attention of the caption decoder on the same mma.sync pipeline, fed by TMA: one warp per (head, frame).
// The frame's K and V tiles ([Nv, 64] each, Nv <= 256) land in 128B-swizzled shared memory through ONE bulk tensor load each,
// issued before anything else; the NQ <= 4 beam queries are rows 0..NQ-1 of the 16-row A fragment (rows 8..15 are zero and
// their accumulators are never read).  HBM-bound on the K/V stream: 3 CTAs per SM keep ~150 KB in flight and the ~200 MMAs per
// CTA hide under the loads.
// ---------------------------------------------------------------------------------------------------------------------
// One 64-key block of the decode cross-attention: S = Q K^T for the warp's (<= 8 valid) query rows, online softmax in the log2
// domain, O += P V.  sKb / sVb: 128B-swizzled [64, 64] tiles; key_base: index of the block's first key within the frame.
template <typename T>
__device__ __forceinline__ void xdec_block(uint32_t sKb, uint32_t sVb, int key_base, int Nv, float scale_log2e, const uint32_t (&qf)[4][4],
                                           float (&o)[8][4], float& m_run, float& l_run, int lane) {
    const int tq = lane & 3;
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
        for (int kk2 = 0; kk2 < 2; ++kk2) {
            uint32_t b0, b1, b2, b3;
            ldsm_x4(sKb + tile_off(nt * 8 + (lane & 7), kk2 * 4 + (lane >> 3)), b0, b1, b2, b3);
            mma16816<T>(s[nt], qf[2 * kk2], b0, b1);
            mma16816<T>(s[nt], qf[2 * kk2 + 1], b2, b3);
        }
    }
    const int key0 = key_base + tq * 2;
    float mx = m_run;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        const int key = key0 + nt * 8;
        s[nt][0] = (key < Nv) ? s[nt][0] * scale_log2e : -INFINITY;
        s[nt][1] = (key + 1 < Nv) ? s[nt][1] * scale_log2e : -INFINITY;
        mx = fmaxf(mx, fmaxf(s[nt][0], s[nt][1]));
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    if (mx == -INFINITY) return;  // a block entirely past the frame's last key (chunk padding)
    const float corr = exp2f(m_run - mx);
    m_run = mx;
    float rs = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        s[nt][0] = exp2f(s[nt][0] - mx);
        s[nt][1] = exp2f(s[nt][1] - mx);
        rs += s[nt][0] + s[nt][1];
    }
    l_run = l_run * corr + rs;
#pragma unroll
    for (int nd = 0; nd < 8; ++nd) {
        o[nd][0] *= corr;
        o[nd][1] *= corr;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t pa[4];
        pa[0] = pack2<T>(s[2 * j][0], s[2 * j][1]);
        pa[1] = 0u;
        pa[2] = pack2<T>(s[2 * j + 1][0], s[2 * j + 1][1]);
        pa[3] = 0u;
#pragma unroll
        for (int nd2 = 0; nd2 < 4; ++nd2) {
            uint32_t b0, b1, b2, b3;
            const int r = j * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
            const int c = nd2 * 2 + (lane >> 4);
            ldsm_x4_trans(sVb + tile_off(r, c), b0, b1, b2, b3);
            mma16816<T>(o[2 * nd2], pa, b0, b1);
            mma16816<T>(o[2 * nd2 + 1], pa, b2, b3);
        }
    }
}

// Q fragments of the decode cross-attention straight from global memory: row g = lane / 4 of the m16 tile.
template <typename T>
__device__ __forceinline__ void xdec_load_q(const T* q, int f, int nq, int D, int h, int lane, uint32_t (&qf)[4][4]) {
    const int g = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        qf[kk][0] = qf[kk][1] = qf[kk][2] = qf[kk][3] = 0u;
        if (g < nq) {
            const T* qp = q + (static_cast<int64_t>(f) * nq + g) * D + h * HD + kk * 16 + tq * 2;
            qf[kk][0] = *reinterpret_cast<const uint32_t*>(qp);
            qf[kk][2] = *reinterpret_cast<const uint32_t*>(qp + 8);
        }
    }
}

template <typename T>
__device__ __forceinline__ void xdec_store(T* out, int f, int nq, int D, int h, int lane, const float (&o)[8][4], float l_run) {
    const int g = lane >> 2, tq = lane & 3;
    l_run += __shfl_xor_sync(0xffffffffu, l_run, 1);
    l_run += __shfl_xor_sync(0xffffffffu, l_run, 2);
    if (g < nq) {
        const float inv = 1.0f / l_run;
        T* op = out + (static_cast<int64_t>(f) * nq + g) * D + h * HD + tq * 2;
#pragma unroll
        for (int nd = 0; nd < 8; ++nd) *reinterpret_cast<uint32_t*>(op + nd * 8) = pack2<T>(o[nd][0] * inv, o[nd][1] * inv);
    }
}

// More than 256 image tokens per frame (ViT-B/16 @384: 577): the K/V stream goes through a two-stage ring of 128-row boxes, the
// next chunk's two TMA loads in flight while the warp works on the current one.  Chunk padding past the frame's last key reads
// the neighbouring frame's rows (or the tensor map's zero fill at the very end): masked in S, finite in V.
constexpr int XCH = 128;

template <typename T>
__global__ void __launch_bounds__(32)
    cross_decode_mma_chunked_kernel(const __grid_constant__ CUtensorMap kv_map, const T* __restrict__ q, T* __restrict__ out, int row0,
                                    int Nv, int nq, int H, float scale_log2e) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    extern __shared__ uint8_t xsmem_raw[];
    uint8_t* xsmem = xsmem_raw + ((1024u - (ptx::smem_u32(xsmem_raw) & 1023u)) & 1023u);
    const int h = blockIdx.x, f = blockIdx.y;
    const int D = H * HD;
    const int lane = threadIdx.x;
    constexpr uint32_t TILE = XCH * 128;                       // one K or V chunk
    uint64_t* bars = reinterpret_cast<uint64_t*>(xsmem + 4 * TILE);
    if (lane == 0) {
        ptx::mbar_init(&bars[0], 1);
        ptx::mbar_init(&bars[1], 1);
        ptx::fence_mbar_init();
    }
    __syncwarp();
    const int n_chunks = (Nv + XCH - 1) / XCH;
    auto issue = [&](int c) {
        const int st = c & 1;
        uint8_t* dst = xsmem + st * 2 * TILE;
        ptx::mbar_arrive_expect_tx(&bars[st], 2 * TILE);
        ptx::tma_load_2d(&kv_map, &bars[st], dst, h * HD, row0 + f * Nv + c * XCH);
        ptx::tma_load_2d(&kv_map, &bars[st], dst + TILE, D + h * HD, row0 + f * Nv + c * XCH);
    };
    if (lane == 0) {
        issue(0);
        if (n_chunks > 1) issue(1);
    }
    uint32_t qf[4][4];
    xdec_load_q<T>(q, f, nq, D, h, lane, qf);
    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    for (int c = 0; c < n_chunks; ++c) {
        const int st = c & 1;
        ptx::mbar_wait(&bars[st], (c >> 1) & 1);
        const uint32_t sK = ptx::smem_u32(xsmem + st * 2 * TILE), sV = sK + TILE;
#pragma unroll
        for (int kb = 0; kb < XCH / BKV; ++kb)
            xdec_block<T>(sK + kb * (BKV * 128), sV + kb * (BKV * 128), c * XCH + kb * BKV, Nv, scale_log2e, qf, o, m_run, l_run, lane);
        __syncwarp();
        if (c + 2 < n_chunks) {
            ptx::fence_proxy_async_smem();  // this warp's ldmatrix reads of the stage precede the async-proxy refill
            if (lane == 0) issue(c + 2);
        }
    }
    xdec_store<T>(out, f, nq, D, h, lane, o, l_run);
}

template <typename T>
__global__ void __launch_bounds__(32)
    cross_decode_mma_kernel(const __grid_constant__ CUtensorMap kv_map, const T* __restrict__ q, T* __restrict__ out, int row0, int Nv,
                            int nq, int H, float scale_log2e) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    extern __shared__ uint8_t xsmem_raw[];
    // the 128B swizzle pattern is a function of the address bits: tiles must start on a 1024-byte boundary
    uint8_t* xsmem = xsmem_raw + ((1024u - (ptx::smem_u32(xsmem_raw) & 1023u)) & 1023u);
    const int h = blockIdx.x, f = blockIdx.y;
    const int D = H * HD;
    const int lane = threadIdx.x;
    const int rows_pad = (Nv + BKV - 1) / BKV * BKV;
    uint8_t* sKp = xsmem;
    uint8_t* sVp = xsmem + static_cast<size_t>(rows_pad) * 128;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sVp + static_cast<size_t>(rows_pad) * 128);
    const uint32_t sK = ptx::smem_u32(sKp), sV = ptx::smem_u32(sVp);
    if (lane == 0) {
        ptx::mbar_init(&bars[0], 1);
        ptx::mbar_init(&bars[1], 1);
        ptx::fence_mbar_init();
    }
    __syncwarp();
    if (lane == 0) {
        const uint32_t bytes = static_cast<uint32_t>(Nv) * 128;
        ptx::mbar_arrive_expect_tx(&bars[0], bytes);
        ptx::tma_load_2d(&kv_map, &bars[0], sKp, h * HD, row0 + f * Nv);
        ptx::mbar_arrive_expect_tx(&bars[1], bytes);
        ptx::tma_load_2d(&kv_map, &bars[1], sVp, D + h * HD, row0 + f * Nv);
    }
    // rows Nv..rows_pad-1 are not written by the TMA boxes: zero them so that P = 0 times V stays 0
    for (int i = Nv * 8 + lane; i < rows_pad * 8; i += 32) {
        *reinterpret_cast<uint4*>(sKp + static_cast<size_t>(i) * 16) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(sVp + static_cast<size_t>(i) * 16) = make_uint4(0, 0, 0, 0);
    }
    uint32_t qf[4][4];
    xdec_load_q<T>(q, f, nq, D, h, lane, qf);
    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    __syncwarp();
    ptx::mbar_wait(&bars[0], 0);
    ptx::mbar_wait(&bars[1], 0);
    const int num_kb = rows_pad / BKV;
    for (int kb = 0; kb < num_kb; ++kb)
        xdec_block<T>(sK + kb * (BKV * 128), sV + kb * (BKV * 128), kb * BKV, Nv, scale_log2e, qf, o, m_run, l_run, lane);
    xdec_store<T>(out, f, nq, D, h, lane, o, l_run);
}

}  // namespace

int cross_decode_mma_chunk_rows() { return XCH; }

int cross_decode_mma_run(const CUtensorMap& kv_map_sw128, int row0, const void* q, void* out, DType dt, int F, int nq, int Nv, int H,
                         float scale, cudaStream_t stream) {
    if (F <= 0) return 0;
    if (nq < 1 || nq > 8 || Nv < 1 || F > 65535) {
        set_error("decode cross-attention (mma): %d query rows (1..8), %d image tokens, %d frames (<= 65535)", nq, Nv, F);
        return 1;
    }
    const bool chunked = Nv > 256;  // the map's box is [Nv, 64] up to 256 tokens, [XCH, 64] beyond
    const int rows_pad = (Nv + BKV - 1) / BKV * BKV;
    const size_t smem = (chunked ? 4 * static_cast<size_t>(XCH) * 128 : 2 * static_cast<size_t>(rows_pad) * 128) + 16 + 1024;
    const float sl2 = scale * 1.4426950408889634f;
#define VIDIL_XDEC(T)                                                                                                          \
    do {                                                                                                                       \
        auto k = chunked ? cross_decode_mma_chunked_kernel<T> : cross_decode_mma_kernel<T>;                                    \
        /* per launch (cheap): the attribute belongs to the current device's copy of the function */                         \
        VIDIL_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));           \
        VIDIL_CUDA_OK(launch_pdl(k, dim3(H, F), dim3(32), smem, stream, kv_map_sw128, reinterpret_cast<const T*>(q),         \
                                 reinterpret_cast<T*>(out), row0, Nv, nq, H, sl2));                                           \
    } while (0)
    if (dt == DT_BF16)
        VIDIL_XDEC(__nv_bfloat16);
    else
        VIDIL_XDEC(__half);
#undef VIDIL_XDEC
    VIDIL_CUDA_OK(cudaGetLastError());
    count_launches(1);
    return 0;
}

namespace {
}  // namespace

int attention_x_run(const void* q, int64_t q_stride, const void* k, const void* v, int64_t kv_stride, const int32_t* frame_of_group,
                    const int32_t* key_mask, void* out, int64_t out_stride, DType dt, int groups, int nq, int nk, int H, bool causal,
                    float scale, cudaStream_t stream) {
    if (groups <= 0 || nq <= 0 || nk <= 0 || H <= 0) return 0;
    if (H > 65535 || groups > 65535) {
        set_error("attention: H and the number of query groups must be <= 65535 (grid y/z limits); got H=%d groups=%d", H, groups);
        return 1;
    }
    if ((q_stride | kv_stride | out_stride) % 8 != 0 || ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) |
                                                            reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(out)) & 15)) {
        set_error("attention: operands must be 16-byte aligned with row strides that are multiples of 8 elements");
        return 1;
    }
    const float sl2 = scale * 1.4426950408889634f;
    const int stages = nk > BKV ? 2 : 1;
    // short query groups (text self-attention, <= 32 tokens) run with 2 warps per CTA: less shared memory and fewer registers per
    // CTA, so twice as many of these latency-bound CTAs are resident
#define VIDIL_XATT(T, NW)                                                                                                        \
    do {                                                                                                                         \
        const size_t smem = static_cast<size_t>(NW) * 16 * HD * 2 + static_cast<size_t>(stages) * 2 * BKV * HD * 2;               \
        auto kern = attention_x_kernel<T, NW>;                                                                                   \
        const dim3 grid((nq + NW * 16 - 1) / (NW * 16), H, groups);                                                               \
        VIDIL_CUDA_OK(launch_pdl(kern, grid, dim3(NW * 32), smem, stream, reinterpret_cast<const T*>(q), q_stride,                \
                                 reinterpret_cast<const T*>(k), reinterpret_cast<const T*>(v), kv_stride, frame_of_group, key_mask, \
                                 reinterpret_cast<T*>(out), out_stride, nq, nk, causal ? 1 : 0, sl2));                            \
    } while (0)
    if (dt == DT_BF16) {
        if (nq <= 32) VIDIL_XATT(__nv_bfloat16, 2); else VIDIL_XATT(__nv_bfloat16, 4);
    } else {
        if (nq <= 32) VIDIL_XATT(__half, 2); else VIDIL_XATT(__half, 4);
    }
#undef VIDIL_XATT
    VIDIL_CUDA_OK(cudaGetLastError());
    count_launches(1);
    return 0;
}

int attention_run(const void* qkv, void* out, DType dt, int B, int N, int H, float scale, cudaStream_t stream) {
    if (B <= 0 || N <= 0 || H <= 0) return 0;
    if (H > 65535 || B > 65535) {
        set_error("attention: H and B must be <= 65535 (grid y/z limits); got H=%d B=%d", H, B);
        return 1;
    }
    const dim3 grid((N + BQ - 1) / BQ, H, B);
    const float sl2 = scale * 1.4426950408889634f;
    if (dt == DT_BF16)
        attention_kernel<__nv_bfloat16><<<grid, ATT_THREADS, 0, stream>>>(
            reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<__nv_bfloat16*>(out), N, H, sl2);
    else
        attention_kernel<__half><<<grid, ATT_THREADS, 0, stream>>>(reinterpret_cast<const __half*>(qkv),
                                                                   reinterpret_cast<__half*>(out), N, H, sl2);
    VIDIL_CUDA_OK(cudaGetLastError());
    count_launches(1);
    return 0;
}

}  // namespace vidil
