// LayerNorm over the channel dim of the fp32 residual stream, emitting the 16-bit operand the next
// GEMM's TMA load consumes (or fp32 for the module output).  Reference: nn.LayerNorm(eps=1e-6) at
// models/vit.py:94,99 (Block.norm1/norm2) and :161,192 (final norm); CLIP uses eps=1e-5.
//
// HBM-bound: one warp owns one row, reads it once with 128-bit coalesced loads, keeps it in registers,
// reduces mean and centred variance with warp shuffles (two-pass in registers, fp32), and writes once.
#include <stdlib.h>

#include "kernels.h"
#include "ptx.cuh"

namespace vidil {
namespace {

constexpr int LN_WARPS = 8;

template <typename T>
__device__ __forceinline__ void store4(T* p, float a, float b, float c, float d);
template <>
__device__ __forceinline__ void store4<float>(float* p, float a, float b, float c, float d) {
    *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}
template <>
__device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, float a, float b, float c, float d) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&lo);
    u.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(p) = u;
}
template <>
__device__ __forceinline__ void store4<__half>(__half* p, float a, float b, float c, float d) {
    __half2 lo = __floats2half2_rn(a, b), hi = __floats2half2_rn(c, d);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&lo);
    u.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(p) = u;
}

// VPL = float4 vectors per lane; D = VPL * 128.
template <typename OutT, int VPL>
__global__ void __launch_bounds__(LN_WARPS * 32)
    layernorm_kernel(const float* __restrict__ in, int64_t in_row_stride, const float* __restrict__ gamma,
                     const float* __restrict__ beta, OutT* __restrict__ out, int rows, float eps, int ascending) {
    constexpr int D = VPL * 128;
    ptx::griddep_launch();
    ptx::griddep_wait();
    // Rows are walked from the LAST to the first: the producer of `in` (the residual-add GEMM, tiles in ascending row order) wrote
    // the last rows most recently, so that is where the 126 MB L2 still holds part of a 206 MB activation; and the GEMM that
    // consumes `out` starts at row 0, which this kernel writes last.
    const int walk = blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
    const int row = ascending ? walk : rows - 1 - walk;   // ascending: developer A/B switch VIDIL_ROWS_ASC=1
    const int lane = threadIdx.x & 31;
    if (walk >= rows) return;
    const float4* src = reinterpret_cast<const float4*>(in + static_cast<int64_t>(row) * in_row_stride);
    float4 x[VPL];
#pragma unroll
    for (int j = 0; j < VPL; ++j) x[j] = src[lane + 32 * j];

    float s = 0.f;
#pragma unroll
    for (int j = 0; j < VPL; ++j) s += (x[j].x + x[j].y) + (x[j].z + x[j].w);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / D);

    float q = 0.f;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
        const float a = x[j].x - mean, b = x[j].y - mean, c = x[j].z - mean, d = x[j].w - mean;
        q += (a * a + b * b) + (c * c + d * d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.0f / D) + eps);

    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    const float4* b4 = reinterpret_cast<const float4*>(beta);
    OutT* dst = out + static_cast<int64_t>(row) * D;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
        const float4 g = __ldg(g4 + lane + 32 * j);
        const float4 b = __ldg(b4 + lane + 32 * j);
        store4<OutT>(dst + 4 * (lane + 32 * j), (x[j].x - mean) * rstd * g.x + b.x, (x[j].y - mean) * rstd * g.y + b.y,
                     (x[j].z - mean) * rstd * g.z + b.z, (x[j].w - mean) * rstd * g.w + b.w);
    }
}

// Post-LN of the text stack: the normalised row replaces the fp32 row in place and is also emitted as the next GEMM's
// 16-bit operand; or (IN16) a 16-bit row in, 16-bit row out.  Same one-warp-per-row, registers-only scheme.
template <typename T, int VPL, bool IN16>
__global__ void __launch_bounds__(LN_WARPS * 32)
    layernorm_post_kernel(float* __restrict__ resid, const T* __restrict__ in16, const float* __restrict__ gamma,
                          const float* __restrict__ beta, T* __restrict__ out16, int rows, float eps) {
    constexpr int D = VPL * 128;
    ptx::griddep_launch();
    ptx::griddep_wait();
    const int row = blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    float4 x[VPL];
    if (IN16) {
        const uint2* src = reinterpret_cast<const uint2*>(in16 + static_cast<int64_t>(row) * D);
#pragma unroll
        for (int j = 0; j < VPL; ++j) {
            const uint2 u = src[lane + 32 * j];
            const T* h = reinterpret_cast<const T*>(&u);
            x[j] = make_float4(static_cast<float>(h[0]), static_cast<float>(h[1]), static_cast<float>(h[2]), static_cast<float>(h[3]));
        }
    } else {
        const float4* src = reinterpret_cast<const float4*>(resid + static_cast<int64_t>(row) * D);
#pragma unroll
        for (int j = 0; j < VPL; ++j) x[j] = src[lane + 32 * j];
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < VPL; ++j) s += (x[j].x + x[j].y) + (x[j].z + x[j].w);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
        const float a = x[j].x - mean, b = x[j].y - mean, c = x[j].z - mean, d = x[j].w - mean;
        q += (a * a + b * b) + (c * c + d * d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.0f / D) + eps);
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
        const float4 g = __ldg(g4 + lane + 32 * j);
        const float4 b = __ldg(b4 + lane + 32 * j);
        const float y0 = (x[j].x - mean) * rstd * g.x + b.x, y1 = (x[j].y - mean) * rstd * g.y + b.y;
        const float y2 = (x[j].z - mean) * rstd * g.z + b.z, y3 = (x[j].w - mean) * rstd * g.w + b.w;
        if (!IN16) store4<float>(resid + static_cast<int64_t>(row) * D + 4 * (lane + 32 * j), y0, y1, y2, y3);
        if (out16 != nullptr) store4<T>(out16 + static_cast<int64_t>(row) * D + 4 * (lane + 32 * j), y0, y1, y2, y3);
    }
}

template <typename T, bool IN16>
int launch_ln_post(float* resid, const void* in16, const float* g, const float* b, void* out16, int rows, int D, float eps,
                   cudaStream_t s) {
    const int grid = (rows + LN_WARPS - 1) / LN_WARPS;
    const T* i16 = reinterpret_cast<const T*>(in16);
    T* o = reinterpret_cast<T*>(out16);
    switch (D) {
        case 768: VIDIL_CUDA_OK(launch_pdl(layernorm_post_kernel<T, 6, IN16>, dim3(grid), dim3(LN_WARPS * 32), 0, s, resid, i16, g, b, o, rows, eps)); break;
        case 1024: VIDIL_CUDA_OK(launch_pdl(layernorm_post_kernel<T, 8, IN16>, dim3(grid), dim3(LN_WARPS * 32), 0, s, resid, i16, g, b, o, rows, eps)); break;
        case 512: VIDIL_CUDA_OK(launch_pdl(layernorm_post_kernel<T, 4, IN16>, dim3(grid), dim3(LN_WARPS * 32), 0, s, resid, i16, g, b, o, rows, eps)); break;
        case 256: VIDIL_CUDA_OK(launch_pdl(layernorm_post_kernel<T, 2, IN16>, dim3(grid), dim3(LN_WARPS * 32), 0, s, resid, i16, g, b, o, rows, eps)); break;
        case 128: VIDIL_CUDA_OK(launch_pdl(layernorm_post_kernel<T, 1, IN16>, dim3(grid), dim3(LN_WARPS * 32), 0, s, resid, i16, g, b, o, rows, eps)); break;
        default: set_error("layernorm: unsupported width %d (supported: 128, 256, 512, 768, 1024)", D); return 1;
    }
    VIDIL_CUDA_OK(cudaGetLastError());
    count_launches(1);
    return 0;
}

template <typename OutT>
int launch_ln(const float* in, int64_t stride, const float* g, const float* b, void* out, int rows, int D, float eps,
              cudaStream_t s) {
    const int grid = (rows + LN_WARPS - 1) / LN_WARPS;
    static const int asc = [] { const char* e = getenv("VIDIL_ROWS_ASC"); return e ? atoi(e) : 0; }();
    OutT* o = reinterpret_cast<OutT*>(out);
    switch (D) {
        case 768: VIDIL_CUDA_OK(launch_pdl(layernorm_kernel<OutT, 6>, dim3(grid), dim3(LN_WARPS * 32), 0, s, in, stride, g, b, o, rows, eps, asc)); break;
        case 1024: VIDIL_CUDA_OK(launch_pdl(layernorm_kernel<OutT, 8>, dim3(grid), dim3(LN_WARPS * 32), 0, s, in, stride, g, b, o, rows, eps, asc)); break;
        case 1280: VIDIL_CUDA_OK(launch_pdl(layernorm_kernel<OutT, 10>, dim3(grid), dim3(LN_WARPS * 32), 0, s, in, stride, g, b, o, rows, eps, asc)); break;
        case 512: VIDIL_CUDA_OK(launch_pdl(layernorm_kernel<OutT, 4>, dim3(grid), dim3(LN_WARPS * 32), 0, s, in, stride, g, b, o, rows, eps, asc)); break;
        case 256: VIDIL_CUDA_OK(launch_pdl(layernorm_kernel<OutT, 2>, dim3(grid), dim3(LN_WARPS * 32), 0, s, in, stride, g, b, o, rows, eps, asc)); break;
        case 128: VIDIL_CUDA_OK(launch_pdl(layernorm_kernel<OutT, 1>, dim3(grid), dim3(LN_WARPS * 32), 0, s, in, stride, g, b, o, rows, eps, asc)); break;
        default: set_error("layernorm: unsupported width %d (supported: 128, 256, 512, 768, 1024, 1280)", D); return 1;
    }
    VIDIL_CUDA_OK(cudaGetLastError());
    count_launches(1);
    return 0;
}

}  // namespace

int layernorm_run(const float* in, int64_t in_row_stride, const float* gamma, const float* beta, void* out,
                  bool out_f32, DType dt, int rows, int D, float eps, cudaStream_t stream) {
    if (rows <= 0) return 0;
    if (in_row_stride % 4 != 0) {
        set_error("layernorm: row stride must be a multiple of 4 floats");
        return 1;
    }
    if (out_f32) return launch_ln<float>(in, in_row_stride, gamma, beta, out, rows, D, eps, stream);
    if (dt == DT_BF16) return launch_ln<__nv_bfloat16>(in, in_row_stride, gamma, beta, out, rows, D, eps, stream);
    return launch_ln<__half>(in, in_row_stride, gamma, beta, out, rows, D, eps, stream);
}

int layernorm_post_run(float* resid, const float* gamma, const float* beta, void* out16, DType dt, int rows, int D, float eps,
                       cudaStream_t stream) {
    if (rows <= 0) return 0;
    if (dt == DT_BF16) return launch_ln_post<__nv_bfloat16, false>(resid, nullptr, gamma, beta, out16, rows, D, eps, stream);
    return launch_ln_post<__half, false>(resid, nullptr, gamma, beta, out16, rows, D, eps, stream);
}

int layernorm16_run(const void* in16, const float* gamma, const float* beta, void* out16, DType dt, int rows, int D, float eps,
                    cudaStream_t stream) {
    if (rows <= 0) return 0;
    if (dt == DT_BF16) return launch_ln_post<__nv_bfloat16, true>(nullptr, in16, gamma, beta, out16, rows, D, eps, stream);
    return launch_ln_post<__half, true>(nullptr, in16, gamma, beta, out16, rows, D, eps, stream);
}

}  // namespace vidil
