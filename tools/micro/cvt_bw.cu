// Micro-benchmark: throughput of fp32 -> 16-bit pack conversions and a few ALU ops per SM sub-partition.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(int iters, long long* cycles, float* sink) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-3f + i;
    uint32_t acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { __nv_bfloat162 v = __floats2bfloat162_rn(x[2 * i] , x[2 * i + 1]); acc[i] += *reinterpret_cast<uint32_t*>(&v); }
            if (MODE == 1) { __half2 v = __floats2half2_rn(x[2 * i], x[2 * i + 1]); acc[i] += *reinterpret_cast<uint32_t*>(&v); }
            if (MODE == 2) { acc[i] += __byte_perm(__float_as_uint(x[2 * i]), __float_as_uint(x[2 * i + 1]), 0x7632); }
            if (MODE == 3) { acc[i] += (__float_as_uint(x[2 * i]) >> 16) | (__float_as_uint(x[2 * i + 1]) & 0xffff0000u); }
            x[2 * i] += 1.0f;
        }
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    uint32_t a = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) a ^= acc[i];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(a);
}
int main() {
    long long* cyc; float* sink;
    cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 148 * 512 * 4);
    const int iters = 4000;
    const char* names[4] = {"F2FP.BF16 pack (+FADD,IADD)", "F2FP.F16 pack (+FADD,IADD)", "PRMT truncate (+FADD,IADD)", "SHF/LOP truncate (+FADD,IADD)"};
    for (int mode = 0; mode < 4; ++mode) for (int threads : {128, 512}) {
        switch (mode) {
            case 0: k<0><<<148, threads>>>(iters, cyc, sink); break;
            case 1: k<1><<<148, threads>>>(iters, cyc, sink); break;
            case 2: k<2><<<148, threads>>>(iters, cyc, sink); break;
            default: k<3><<<148, threads>>>(iters, cyc, sink); break;
        }
        cudaDeviceSynchronize();
        long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        const double warps_per_smsp = threads / 32 / 4.0;
        printf("%-32s %3d threads: %.2f cycles per warp-pack-group per SMSP\n", names[mode], threads, h / ((double)iters * 8 * warps_per_smsp));
    }
    return 0;
}
