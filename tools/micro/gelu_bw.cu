// Micro-benchmark: cost per element of the GEMM epilogue math (bias shuffle + A&S erf GELU + pack) per SM sub-partition.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../../vidil_b200/csrc/ptx.cuh"
using namespace vidil;

__device__ __forceinline__ float gelu_erf(float x) {
    const float t = ptx::rcp_approx(fmaf(fabsf(x), 0.3275911f * 0.70710678118654752f, 1.0f));
    float p = fmaf(t, 1.061405429f, -1.453152027f);
    p = fmaf(t, p, 1.421413741f);
    p = fmaf(t, p, -0.284496736f);
    p = fmaf(t, p, 0.254829592f);
    p *= t;
    const float e = ptx::ex2_approx(x * x * (-0.5f * 1.4426950408889634f));
    return fmaf(-fabsf(0.5f * x), p * e, fmaxf(x, 0.0f));
}
__device__ __forceinline__ float gelu_tanh_like(float x) {  // 1 MUFU variant for comparison: x * sigmoid(1.702 x)
    return x * ptx::rcp_approx(1.0f + ptx::ex2_approx(x * (-1.702f * 1.4426950408889634f)));
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(int iters, long long* cycles, float* sink) {
    float x[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) x[i] = (threadIdx.x % 7) * 0.3f - 1.0f + i * 0.01f;
    float b = threadIdx.x * 0.001f;
    uint32_t acc = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
            float v0 = x[i], v1 = x[i + 1];
            if (MODE >= 1) { v0 += __shfl_sync(0xffffffffu, b, i); v1 += __shfl_sync(0xffffffffu, b, i + 1); }
            if (MODE == 0 || MODE == 1) { v0 = gelu_erf(v0); v1 = gelu_erf(v1); }
            if (MODE == 3) { v0 = gelu_tanh_like(v0); v1 = gelu_tanh_like(v1); }
            __nv_bfloat162 pk = __floats2bfloat162_rn(v0, v1);
            acc ^= *reinterpret_cast<uint32_t*>(&pk);
            x[i] = v0 * 0.5f + 0.1f; x[i + 1] = v1 * 0.5f - 0.1f;
        }
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(acc);
}

int main() {
    long long* cyc; float* sink;
    cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 148 * 512 * 4);
    const int iters = 2000;
    const char* names[4] = {"gelu_erf", "shfl+gelu_erf", "shfl only", "shfl+quick_gelu"};
    for (int mode = 0; mode < 4; ++mode) for (int threads : {256, 512}) {
        switch (mode) {
            case 0: k<0><<<148, threads>>>(iters, cyc, sink); break;
            case 1: k<1><<<148, threads>>>(iters, cyc, sink); break;
            case 2: k<2><<<148, threads>>>(iters, cyc, sink); break;
            default: k<3><<<148, threads>>>(iters, cyc, sink); break;
        }
        cudaDeviceSynchronize();
        long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        const double warps_per_smsp = threads / 32 / 4.0;
        printf("%-16s %3d threads: %.2f cycles per warp-element per SMSP (%.0f warps/SMSP)\n", names[mode], threads,
               h / ((double)iters * 32 * warps_per_smsp), warps_per_smsp);
    }
    return 0;
}
