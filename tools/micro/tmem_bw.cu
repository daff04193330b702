// Micro-benchmark: tcgen05.ld throughput per SM as a function of the number of reading warps and the load width.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../vidil_b200/csrc/ptx.cuh"
using namespace vidil;

template <int WIDTH>
__global__ void __launch_bounds__(512, 1) tmem_read(int iters, long long* cycles, float* sink, int nwarps) {
    __shared__ uint32_t tptr;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) ptx::tmem_alloc<1>(&tptr, 512);
    ptx::tcgen05_fence_before();
    __syncthreads();
    ptx::tcgen05_fence_after();
    const uint32_t base = tptr + ((uint32_t)((warp & 3) * 32) << 16);
    float acc = 0.f;
    __syncthreads();
    long long t0 = clock64();
    if (warp < nwarps) {
        for (int i = 0; i < iters; ++i) {
            if (WIDTH == 32) {
                uint32_t r[32];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    ptx::tmem_ld_32x32b_x32(base + ((i + c) & 7) * 32, r);
                    ptx::tmem_ld_wait();
                    acc += __uint_as_float(r[c]);
                }
            } else {
                uint32_t r0[32], r1[32], r2[32], r3[32];
                ptx::tmem_ld_32x32b_x32(base + 0, r0);
                ptx::tmem_ld_32x32b_x32(base + 32, r1);
                ptx::tmem_ld_32x32b_x32(base + 64, r2);
                ptx::tmem_ld_32x32b_x32(base + 96, r3);
                ptx::tmem_ld_wait();
                acc += __uint_as_float(r0[i & 31]) + __uint_as_float(r1[1]) + __uint_as_float(r2[2]) + __uint_as_float(r3[3]);
            }
        }
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    ptx::tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) ptx::tmem_dealloc<1>(tptr, 512);
}

int main() {
    long long* cyc; float* sink;
    cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 148 * 512 * 4);
    const int iters = 2000;
    for (int width : {32, 128}) for (int nw : {1, 4, 8, 16}) {
        if (width == 32) tmem_read<32><<<148, 512>>>(iters, cyc, sink, nw); else tmem_read<128><<<148, 512>>>(iters, cyc, sink, nw);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double bytes = (double)iters * 4 * 32 * 32 * 4 * nw;  // per SM
        printf("width %3d warps %2d: %lld cycles, %.1f B/clk/SM  (%s)\n", width, nw, h[0], bytes / h[0], cudaGetErrorString(e));
    }
    return 0;
}
