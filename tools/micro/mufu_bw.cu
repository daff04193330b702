// Micro-benchmark: MUFU.EX2 and FFMA2 issue throughput per SM sub-partition.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../vidil_b200/csrc/ptx.cuh"
using namespace vidil;

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(int iters, long long* cycles, float* sink) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-3f + i;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = ptx::ex2_approx(x[i]);
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) ptx::fma2(x[i], x[i + 1], x[i], x[i + 1], 0.999f, 0.001f);
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], 0.999f, 0.001f);
        } else if (MODE == 4) {  // packed half exponentials: two results per MUFU instruction?
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                uint32_t u = __float_as_uint(x[i]);
                asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u));
                x[i] = __uint_as_float(u);
            }
        } else if (MODE == 5) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                uint32_t u = __float_as_uint(x[i]);
                asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u));
                x[i] = __uint_as_float(u);
            }
        } else {  // mixed: per pair 1 FFMA2 + 2 MUFU + 1 FADD2 (the softmax inner loop)
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                float a0, a1;
                ptx::fma2(a0, a1, x[i], x[i + 1], 0.999f, 0.001f);
                a0 = ptx::ex2_approx(a0);
                a1 = ptx::ex2_approx(a1);
                ptx::add2(x[i], x[i + 1], a0, a1);
            }
        }
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    float acc = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc += x[i];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
    long long* cyc; float* sink;
    cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 148 * 512 * 4);
    const int iters = 4000;
    const char* names[6] = {"MUFU.EX2", "FFMA2", "FFMA", "softmax-mix", "EX2.f16x2", "EX2.bf16x2"};
    for (int mode = 0; mode < 6; ++mode) for (int threads : {128, 256, 512}) {
        switch (mode) {
            case 0: k<0><<<148, threads>>>(iters, cyc, sink); break;
            case 1: k<1><<<148, threads>>>(iters, cyc, sink); break;
            case 2: k<2><<<148, threads>>>(iters, cyc, sink); break;
            case 3: k<3><<<148, threads>>>(iters, cyc, sink); break;
            case 4: k<4><<<148, threads>>>(iters, cyc, sink); break;
            default: k<5><<<148, threads>>>(iters, cyc, sink); break;
        }
        cudaDeviceSynchronize();
        long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        const double warps_per_smsp = threads / 32 / 4.0;
        const double elem_per_warp = (double)iters * 16;
        printf("%-12s %3d threads: %8lld cycles -> %.2f cycles per warp-element-op per SMSP (%.1f warps/SMSP)\n", names[mode], threads, h,
               h / (elem_per_warp * warps_per_smsp), warps_per_smsp);
    }
    return 0;
}
