// Micro-benchmark: the pass-2 softmax chunk of attention_tc.cu (TMEM load, FFMA2, MUFU.EX2, FADD2, F2FP pack, TMEM store) in
// isolation, one or two warps per SM sub-partition, with pieces removed one at a time to see which one sets the pace.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o softmax_chunk softmax_chunk.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../vidil_b200/csrc/attention_tc.cu"
using namespace vidil;
using namespace vidil::attn;

// mode: 0 full, 1 no STTM, 2 no LDTM (registers reused), 3 pack by truncation (PRMT) instead of F2FP, 4 no pack and no STTM,
//       5 MUFU only
template <int MODE, bool MASKED = false, int LAG = 8, int POLY = 0>
__global__ void __launch_bounds__(512, 1) k(int iters, long long* cycles, float* sink, int nwork, int sleep_ns, int nv = 1000) {
    __shared__ uint32_t tptr;
    __shared__ uint64_t spin_bar;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) ptx::tmem_alloc<1>(&tptr, 512);
    if (threadIdx.x == 0) { ptx::mbar_init(&spin_bar, nwork * 32); ptx::fence_mbar_init(); }
    ptx::tcgen05_fence_before();
    __syncthreads();
    ptx::tcgen05_fence_after();
    const uint32_t taddr = tptr + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 256;
    float sum[4] = {0.f, 0.f, 0.f, 0.f};
    uint32_t ra[32], rb[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) ra[i] = rb[i] = __float_as_uint(-1.0f - 0.01f * i - threadIdx.x * 1e-3f);
    __syncthreads();
    const long long t0 = clock64();
    if (warp >= nwork) {  // spinner warps: what the TMA producer / MMA issuers / barrier pollers of the real kernel do
        if ((threadIdx.x & 31) == 0) {
            if (sleep_ns == 0) ptx::mbar_wait(&spin_bar, 0);
            else { uint32_t done = 0; while (!done) { asm volatile("{.reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p;}" : "=r"(done) : "r"(ptx::smem_u32(&spin_bar))); if (!done) __nanosleep(sleep_ns); } }
        }
        __syncwarp();
    } else {
    if (MODE != 2) ptx::tmem_ld_32x32b_x32(taddr, ra);
    for (int it = 0; it < iters; ++it) {
#pragma unroll 1
        for (int c = 0; c < 6; c += 2) {
            auto body = [&](uint32_t (&r)[32], uint32_t (&nxt)[32], int c0) {
                if (MODE != 2) {
                    ptx::tmem_ld_wait();
                    ptx::tmem_ld_32x32b_x32(taddr + ((c0 + 32) % 192), nxt);
                }
                if (MODE == 0 || MODE == 2) {
                    chunk_exp_impl<__nv_bfloat16, 32, MASKED, POLY, LAG>(r, c0, nv, 0.18f, -0.5f, taddr, sum);
                } else {
                    float a[32];
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 32; i += 2) ptx::fma2(a[i], a[i + 1], __uint_as_float(r[i]), __uint_as_float(r[i + 1]), 0.18f, -0.5f);
#pragma unroll
                    for (int i = 0; i < 40; i += 2) {
                        if (i < 32) {
                            a[i] = ptx::ex2_approx(a[i]);
                            a[i + 1] = ptx::ex2_approx(a[i + 1]);
                        }
                        if (i >= 8) {
                            const int j = i - 8;
                            if (MODE != 5) ptx::add2(sum[j & 2], sum[(j & 2) + 1], a[j], a[j + 1]);
                            if (MODE == 1) pk[j >> 1] = pack2<__nv_bfloat16>(a[j], a[j + 1]);
                            if (MODE == 3) pk[j >> 1] = __byte_perm(__float_as_uint(a[j]), __float_as_uint(a[j + 1]), 0x7632);
                            if (MODE == 4 || MODE == 5) sum[0] += (MODE == 5) ? a[j] * 1e-9f + a[j + 1] * 1e-9f : 0.f;
                        }
                    }
                    if (MODE == 3) ptx::tmem_st_32x32b_x16(taddr + (c0 >> 1), pk);
                    if (MODE == 1) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) sum[1] += __uint_as_float(pk[i]) * 1e-30f;
                    }
                }
            };
            body(ra, rb, 32 * c);
            body(rb, ra, 32 * c + 32);
        }
    }
    ptx::tmem_st_wait();
    ptx::tmem_ld_wait();
    ptx::mbar_arrive(&spin_bar);
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = sum[0] + sum[1] + sum[2] + sum[3] + __uint_as_float(ra[3]) + __uint_as_float(rb[5]);
    ptx::tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) ptx::tmem_dealloc<1>(tptr, 512);
}

int main() {
    long long* cyc;
    float* sink;
    cudaMalloc(&cyc, 148 * 8);
    cudaMalloc(&sink, 148 * 256 * 4);
    const int iters = 500;
    const char* names[6] = {"full chunk (LDTM FFMA2 MUFU FADD2 F2FP STTM)", "no STTM", "no LDTM", "PRMT truncation instead of F2FP", "no pack, no STTM", "MUFU + FFMA2 only"};
    auto report = [&](const char* what, int lag, cudaError_t e) {
        long long h = 0;
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("full chunk, 1 working warp/SMSP, %s, LAG %2d: %6.1f clocks per 32-column chunk %s\n", what, lag, (double)h / (iters * 6.0),
               e == cudaSuccess ? "" : cudaGetErrorString(e));
    };
#define RUN(POLY, LAG)                                                        \
    k<0, false, LAG, POLY><<<148, 128>>>(iters, cyc, sink, 4, 0, 1000); \
    printf("poly pairs per 16: %d  ", POLY); report("unmasked path", LAG, cudaDeviceSynchronize());
    RUN(0, 8) RUN(4, 8) RUN(6, 8) RUN(8, 8) RUN(10, 8) RUN(6, 2) RUN(6, 16) RUN(8, 16)
    return 0;
}
// link stubs for the host side of attention_tc.cu (unused here)
namespace vidil {
void set_error(const char*, ...) {}
int gemm_num_sms() { return 148; }
bool pdl_enabled() { return false; }
void count_launches(int) {}
void attention_tc257_set_trace(long long*) {}
}  // namespace vidil
