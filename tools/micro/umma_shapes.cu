// Micro-benchmark: clocks per tcgen05.mma (kind::f16, cta_group::1, M=128, K=16) as a function of N, of where the A operand
// lives (shared memory descriptor vs tensor memory) and of the B operand's major-ness.  One thread of one CTA per SM issues R
// accumulating MMAs back to back, commits, and waits; the clock64 delta / R is the sustained per-instruction time.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma_shapes umma_shapes.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../vidil_b200/csrc/ptx.cuh"
using namespace vidil;

// mode bit0: A from TMEM; bit1: B MN-major
__global__ void __launch_bounds__(128, 1) k(int mode, int N, int R, long long* cycles, int naccs, int two_issuers = 0) {
    uint32_t tmem_off = 0;
    __shared__ uint64_t bar2;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const uint32_t raw = ptx::smem_u32(smem_raw);
    const uint32_t sbase = raw + ((1024u - (raw & 1023u)) & 1023u);
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x3c003c00u;  // fp16 1.0
    if (threadIdx.x == 0) {
        ptx::mbar_init(&bar, 1);
        ptx::mbar_init(&bar2, 1);
        ptx::fence_mbar_init();
    }
    if (threadIdx.x < 32) ptx::tmem_alloc<1>(&tmem_slot, 512);
    ptx::tcgen05_fence_before();
    __syncthreads();
    ptx::tcgen05_fence_after();
    uint32_t tmem = tmem_slot;
    ptx::fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0 || (two_issuers && threadIdx.x == 32)) {
        const bool a_tmem = mode & 1, b_mn = mode & 2;
        if (threadIdx.x == 32) tmem_off = 128;
        const uint32_t idesc = b_mn ? ptx::make_idesc_f16_bmn(false, 128, N) : ptx::make_idesc_f16(false, 128, N);
        const uint64_t da = ptx::make_kmajor_sw128_desc(sbase);
        const uint64_t db = b_mn ? ptx::make_smem_desc(sbase + 16384, 8192, 1024, 2) : ptx::make_kmajor_sw128_desc(sbase + 16384);
        tmem += tmem_off;
        uint64_t* mybar = threadIdx.x == 0 ? &bar : &bar2;
        const long long t0 = clock64();
        if (naccs < 0) {  // unrolled, rotating over -naccs independent accumulators (64 columns apart)
            const int na = -naccs;
            for (int r = 0; r < R; r += 8) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t d = tmem + 256 + (j % (na == 2 ? 2 : 4)) * 64;
                    if (a_tmem)
                        ptx::umma_f16_tmem_a(d, tmem + (j & 3) * 8, db + (b_mn ? (j & 3) * 128 : 2 * (j & 3)), idesc, 1);
                    else
                        ptx::umma_f16<1>(d, da + 2 * (j & 3), db + (b_mn ? (j & 3) * 128 : 2 * (j & 3)), idesc, 1);
                }
            }
        } else if (naccs == 0) {  // unrolled issue loop, every operand address a compile-time offset: the issue cost of the thread itself
            for (int r = 0; r < R; r += 8) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (a_tmem)
                        ptx::umma_f16_tmem_a(tmem + 256, tmem + (j & 3) * 8, db + (b_mn ? (j & 3) * 128 : 2 * (j & 3)), idesc, 1);
                    else
                        ptx::umma_f16<1>(tmem + 256, da + 2 * (j & 3), db + (b_mn ? (j & 3) * 128 : 2 * (j & 3)), idesc, 1);
                }
            }
        } else
        for (int r = 0; r < R; ++r) {
            const int j = r & 3;
            if (a_tmem)
                ptx::umma_f16_tmem_a(tmem + 256 + (r % naccs) * 64, tmem + j * 8, db + (b_mn ? j * 128 : 2 * j), idesc, 1);
            else
                ptx::umma_f16<1>(tmem + 256 + (r % naccs) * 64, da + 2 * j, db + (b_mn ? j * 128 : 2 * j), idesc, 1);
        }
        ptx::umma_commit<1>(mybar);
        ptx::mbar_wait(mybar, 0);
        const long long t1 = clock64();
        if (blockIdx.x == 0) cycles[0] = t1 - t0;
    }
    ptx::tcgen05_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) ptx::tmem_dealloc<1>(tmem_slot, 512);
}

int main() {
    long long* cyc;
    cudaMalloc(&cyc, 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int R = 4000;
    const char* names[4] = {"A smem, B K-major ", "A TMEM, B K-major ", "A smem, B MN-major", "A TMEM, B MN-major"};
    for (int N : {64, 80, 128, 208, 256})
        for (int mode = 0; mode < 4; ++mode) {
            k<<<148, 128, 64 * 1024>>>(mode, N, R, cyc, 1);
            cudaError_t e = cudaDeviceSynchronize();
            long long h = 0;
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            printf("M=128 N=%3d K=16 %s: %7.1f clocks per MMA (floor %d)%s\n", N, names[mode], (double)h / R, 128 * N / 256,
                   e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    // same-accumulator dependency or a per-instruction floor?  rotate N=64 MMAs over 1, 2, 4 independent accumulators
    for (int naccs : {0, -2, -4})
        for (int mode : {0, 3}) {
            k<<<148, 128, 64 * 1024>>>(mode, 64, R, cyc, naccs);
            cudaDeviceSynchronize();
            long long h = 0;
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            printf("M=128 N= 64 K=16 %s, %s issue loop: %7.1f clocks per MMA\n", names[mode], naccs == 0 ? "unrolled x8, 1 accumulator" : (naccs == -2 ? "unrolled x8, 2 accumulators" : "unrolled x8, 4 accumulators"), (double)h / R);
        }
    for (int mode : {0, 3}) {
        k<<<148, 128, 64 * 1024>>>(mode, 64, R, cyc, 0, 1);
        cudaDeviceSynchronize();
        long long h = 0;
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("M=128 N= 64 K=16 %s, TWO issuing threads (own accumulators): %7.1f clocks per MMA of one thread\n", names[mode], (double)h / R);
    }
    return 0;
}
