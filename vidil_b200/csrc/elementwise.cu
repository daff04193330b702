// Layout and elementwise kernels around the GEMMs.  All are HBM-bound streaming kernels: coalesced
// 128-bit accesses, grid sized from the element count.
#include "kernels.h"
#include "ptx.cuh"

namespace vidil {
namespace {

template <typename T>
__device__ __forceinline__ T from_float(float x);
template <>
__device__ __forceinline__ __nv_bfloat16 from_float<__nv_bfloat16>(float x) {
    return __float2bfloat16_rn(x);
}
template <>
__device__ __forceinline__ __half from_float<__half>(float x) {
    return __float2half_rn(x);
}

// ---------------------------------------------------------------------------------------------
// fp32 [rows, cols] -> T [rows, ld_out], zero padding columns cols..ld_out (weight packing at load).
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void cast_kernel(const float* __restrict__ in, T* __restrict__ out, int64_t rows, int64_t cols,
                            int64_t ld_out) {
    const int64_t total = rows * ld_out;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t r = i / ld_out, c = i - r * ld_out;
        out[i] = (c < cols) ? from_float<T>(in[r * cols + c]) : from_float<T>(0.f);
    }
}

// ---------------------------------------------------------------------------------------------
// Patch gather (the data movement half of the patch-embed Conv2d, vit.py:144-145,182; CLIP conv 14x14).
// frames: fp32 NCHW [B, C, S, S]; patches: T [B*G*G, Kpad] with column c*ps*ps + i*ps + j, i.e. the
// flattening of Conv2d.weight[D, C, ps, ps], so the conv becomes patches @ weight.view(D,-1)^T.
// One thread moves VEC horizontally adjacent pixels: reads are contiguous along image rows (a warp
// covers a 32*VEC*4-byte run), writes are VEC*2-byte pieces of ps*2-byte runs.
// ---------------------------------------------------------------------------------------------
template <typename T, int VEC>
__global__ void im2col_kernel(const float* __restrict__ frames, T* __restrict__ patches, int B, int C, int S, int ps,
                              int Kpad) {
    const int G = S / ps;
    const int xv = S / VEC;  // vectors per image row
    const int64_t total = static_cast<int64_t>(B) * C * S * xv;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int xq = static_cast<int>(i % xv);
        int64_t t = i / xv;
        const int y = static_cast<int>(t % S);
        t /= S;
        const int c = static_cast<int>(t % C);
        const int b = static_cast<int>(t / C);
        const int x = xq * VEC;
        const float* src = frames + ((static_cast<int64_t>(b) * C + c) * S + y) * S + x;
        const int py = y / ps, iy = y - py * ps;
        const int px = x / ps, ix = x - px * ps;  // VEC divides ps, so the vector stays inside one patch
        T* dst = patches + (static_cast<int64_t>(b) * G * G + py * G + px) * Kpad + (c * ps + iy) * ps + ix;
        // element offsets are multiples of VEC (ps, Kpad and ix all are), so the packed stores are aligned
        if constexpr (VEC == 4) {
            const float4 v = *reinterpret_cast<const float4*>(src);
            const T t0 = from_float<T>(v.x), t1 = from_float<T>(v.y), t2 = from_float<T>(v.z), t3 = from_float<T>(v.w);
            uint2 u;
            u.x = static_cast<uint32_t>(*reinterpret_cast<const uint16_t*>(&t0)) |
                  (static_cast<uint32_t>(*reinterpret_cast<const uint16_t*>(&t1)) << 16);
            u.y = static_cast<uint32_t>(*reinterpret_cast<const uint16_t*>(&t2)) |
                  (static_cast<uint32_t>(*reinterpret_cast<const uint16_t*>(&t3)) << 16);
            *reinterpret_cast<uint2*>(dst) = u;
        } else {
            const float2 v = *reinterpret_cast<const float2*>(src);
            const T t0 = from_float<T>(v.x), t1 = from_float<T>(v.y);
            *reinterpret_cast<uint32_t*>(dst) = static_cast<uint32_t>(*reinterpret_cast<const uint16_t*>(&t0)) |
                                                (static_cast<uint32_t>(*reinterpret_cast<const uint16_t*>(&t1)) << 16);
        }
    }
}

// resid[b * tokens + 0, :] = cls + pos[0, :]   (torch.cat((cls, x)) + pos_embed, vit.py:184-187)
__global__ void cls_pos_kernel(const float* __restrict__ cls, const float* __restrict__ pos, float* __restrict__ resid,
                               int B, int tokens, int D) {
    const int64_t total = static_cast<int64_t>(B) * D;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int b = static_cast<int>(i / D), d = static_cast<int>(i - static_cast<int64_t>(b) * D);
        resid[static_cast<int64_t>(b) * tokens * D + d] = cls[d] + pos[d];
    }
}

// out[r, :] = in[r, :] / ||in[r, :]||_2 — one warp per row (CLIP image_embeds normalisation).
__global__ void l2norm_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int D) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* src = in + static_cast<int64_t>(row) * D;
    float s = 0.f;
    for (int i = lane; i < D; i += 32) {
        const float v = src[i];
        s += v * v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float nrm = sqrtf(s);
    float* dst = out + static_cast<int64_t>(row) * D;
    for (int i = lane; i < D; i += 32) dst[i] = src[i] / nrm;
}

template <typename T>
__global__ void uncast_kernel(const T* __restrict__ in, float* __restrict__ out, int64_t n) {
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x)
        out[i] = static_cast<float>(in[i]);
}

// resid[b*L + t, :] = tok[ids[b*L + t], :] + pos[t, :]   (CLIPTextEmbeddings.forward: token + position embedding)
__global__ void embed_tokens_kernel(const int32_t* __restrict__ ids, const float* __restrict__ tok,
                                    const float* __restrict__ pos, float* __restrict__ resid, int64_t rows, int L, int D,
                                    int vocab) {
    const int d4 = D / 4;
    const int64_t total = rows * d4;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t r = i / d4;
        const int c = static_cast<int>(i - r * d4);
        int id = ids[r];
        id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);  // out-of-range ids are clamped, never read out of bounds
        const int t = static_cast<int>(r % L);
        const float4 a = reinterpret_cast<const float4*>(tok + static_cast<int64_t>(id) * D)[c];
        const float4 b = reinterpret_cast<const float4*>(pos + static_cast<int64_t>(t) * D)[c];
        reinterpret_cast<float4*>(resid + r * D)[c] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    }
}

// out[b, :] = in[b*L + idx[b], :]   (the EOS row of every sequence, CLIPTextTransformer pooled_output)
__global__ void gather_rows_kernel(const float* __restrict__ in, const int32_t* __restrict__ idx, float* __restrict__ out,
                                   int B, int L, int D) {
    const int d4 = D / 4;
    const int64_t total = static_cast<int64_t>(B) * d4;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int b = static_cast<int>(i / d4), c = static_cast<int>(i - static_cast<int64_t>(b) * d4);
        int t = idx[b];
        t = t < 0 ? 0 : (t >= L ? L - 1 : t);
        reinterpret_cast<float4*>(out + static_cast<int64_t>(b) * D)[c] =
            reinterpret_cast<const float4*>(in + (static_cast<int64_t>(b) * L + t) * D)[c];
    }
}

int grid_for(int64_t total, int block) {
    int64_t g = (total + block - 1) / block;
    const int64_t cap = 148 * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return static_cast<int>(g);
}

}  // namespace

int cast_run(const float* in, void* out, DType dt, int64_t rows, int64_t cols, int64_t ld_out, cudaStream_t stream) {
    if (rows <= 0 || cols <= 0) return 0;
    if (ld_out < cols) {
        set_error("cast: ld_out %lld < cols %lld", (long long)ld_out, (long long)cols);
        return 1;
    }
    const int grid = grid_for(rows * ld_out, 256);
    if (dt == DT_BF16)
        cast_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(in, reinterpret_cast<__nv_bfloat16*>(out), rows, cols, ld_out);
    else
        cast_kernel<__half><<<grid, 256, 0, stream>>>(in, reinterpret_cast<__half*>(out), rows, cols, ld_out);
    VIDIL_CUDA_OK(cudaGetLastError());
    count_launches(1);
    return 0;
}

int im2col_run(const float* frames, void* patches, DType dt, int B, int C, int img, int ps, int Kpad,
               cudaStream_t stream) {
    if (B <= 0) return 0;
    if (img % ps != 0 || C * ps * ps > Kpad) {
        set_error("im2col: image %d not divisible by patch %d, or Kpad %d too small", img, ps, Kpad);
        return 1;
    }
    const int vec = (ps % 4 == 0 && img % 4 == 0) ? 4 : ((ps % 2 == 0 && img % 2 == 0) ? 2 : 0);
    if (vec == 0) {
        set_error("im2col: patch size %d must be even", ps);
        return 1;
    }
    const int64_t total = static_cast<int64_t>(B) * C * img * (img / vec);
    const int grid = grid_for(total, 256);
#define VIDIL_IM2COL(T, V)                                                                                   \
    im2col_kernel<T, V><<<grid, 256, 0, stream>>>(frames, reinterpret_cast<T*>(patches), B, C, img, ps, Kpad)
    if (dt == DT_BF16) {
        if (vec == 4) VIDIL_IM2COL(__nv_bfloat16, 4); else VIDIL_IM2COL(__nv_bfloat16, 2);
    } else {
        if (vec == 4) VIDIL_IM2COL(__half, 4); else VIDIL_IM2COL(__half, 2);
    }
#undef VIDIL_IM2COL
    VIDIL_CUDA_OK(cudaGetLastError());
    count_launches(1);
    return 0;
}

int cls_pos_run(const float* cls, const float* pos, float* resid, int B, int tokens, int D, cudaStream_t stream) {
    if (B <= 0) return 0;
    cls_pos_kernel<<<grid_for(static_cast<int64_t>(B) * D, 256), 256, 0, stream>>>(cls, pos, resid, B, tokens, D);
    VIDIL_CUDA_OK(cudaGetLastError());
    count_launches(1);
    return 0;
}

int embed_tokens_run(const int32_t* ids, const float* tok, const float* pos, float* resid, int64_t rows, int L, int D,
                     int vocab, cudaStream_t stream) {
    if (rows <= 0) return 0;
    embed_tokens_kernel<<<grid_for(rows * (D / 4), 256), 256, 0, stream>>>(ids, tok, pos, resid, rows, L, D, vocab);
    VIDIL_CUDA_OK(cudaGetLastError());
    count_launches(1);
    return 0;
}

int gather_rows_run(const float* in, const int32_t* idx, float* out, int B, int L, int D, cudaStream_t stream) {
    if (B <= 0) return 0;
    gather_rows_kernel<<<grid_for(static_cast<int64_t>(B) * (D / 4), 256), 256, 0, stream>>>(in, idx, out, B, L, D);
    VIDIL_CUDA_OK(cudaGetLastError());
    count_launches(1);
    return 0;
}

int uncast_run(const void* in, float* out, DType dt, int64_t n, cudaStream_t stream) {
    if (n <= 0) return 0;
    const int grid = grid_for(n, 256);
    if (dt == DT_BF16)
        uncast_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(in), out, n);
    else
        uncast_kernel<__half><<<grid, 256, 0, stream>>>(reinterpret_cast<const __half*>(in), out, n);
    VIDIL_CUDA_OK(cudaGetLastError());
    count_launches(1);
    return 0;
}

int l2norm_run(const float* in, float* out, int rows, int D, cudaStream_t stream) {
    if (rows <= 0) return 0;
    l2norm_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(in, out, rows, D);
    VIDIL_CUDA_OK(cudaGetLastError());
    count_launches(1);
    return 0;
}

}  // namespace vidil
