// Per-frame top-k phrase selection with exact fp32 re-ranking.
// Reference: run_visual_tokenization.py:276 (sims = image_embeds @ text_embeds.t(), fp32) and :306
// (inds = np.argsort(frm_score)[::-1][:topk]) — the reference ships the whole [F,T] fp32 matrix to the
// host and argsorts every row there.
//
// Here the tensor-core GEMM produces approximate scores (fp16 operands, fp32 accumulate).  For each row
// this kernel (one CTA per frame) pulls the row into shared memory once, extracts the NCAND best
// approximate candidates by repeated block-wide argmax, recomputes those NCAND dot products from the
// original fp32 embeddings (one warp per candidate, fixed summation order), and emits the k best by
// (fp32 score descending, index descending on exact ties — what a stable ascending argsort reversed gives).
// NCAND - k spare candidates absorb the ~1e-4 error of the fp16 scores, so the indices equal the fp32 ranking.
#include <float.h>

#include "kernels.h"
#include "ptx.cuh"

namespace vidil {
namespace {

constexpr int TK_THREADS = 256;
constexpr int NCAND = 16;

struct Best {
    float v;
    int i;
};
__device__ __forceinline__ bool better(float v, int i, float bv, int bi) { return v > bv || (v == bv && i > bi); }

__device__ __forceinline__ Best warp_argmax(Best b) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, b.v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, b.i, o);
        if (better(ov, oi, b.v, b.i)) {
            b.v = ov;
            b.i = oi;
        }
    }
    return b;
}

__global__ void __launch_bounds__(TK_THREADS)
    topk_rerank_kernel(const float* __restrict__ scores, int64_t ld, const float* __restrict__ img,
                       const float* __restrict__ bank, int T, int D, int k, float* __restrict__ out_scores,
                       int32_t* __restrict__ out_idx) {
    extern __shared__ float srow[];  // T floats
    __shared__ float red_v[TK_THREADS / 32];
    __shared__ int red_i[TK_THREADS / 32];
    __shared__ int cand_i[NCAND];
    __shared__ float cand_v[NCAND];

    const int f = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float* src = scores + static_cast<int64_t>(f) * ld;
    for (int i = tid; i < T; i += TK_THREADS) srow[i] = src[i];
    __syncthreads();

    const int ncand = (T < NCAND) ? T : NCAND;
    for (int c = 0; c < ncand; ++c) {
        Best b{-FLT_MAX, -1};
        for (int i = tid; i < T; i += TK_THREADS) {
            const float v = srow[i];
            if (better(v, i, b.v, b.i)) {
                b.v = v;
                b.i = i;
            }
        }
        b = warp_argmax(b);
        if (lane == 0) {
            red_v[warp] = b.v;
            red_i[warp] = b.i;
        }
        __syncthreads();
        if (warp == 0) {
            Best w{lane < TK_THREADS / 32 ? red_v[lane] : -FLT_MAX, lane < TK_THREADS / 32 ? red_i[lane] : -1};
            w = warp_argmax(w);
            if (lane == 0) {
                cand_i[c] = w.i;
                if (w.i >= 0) srow[w.i] = -FLT_MAX;  // remove from further rounds
            }
        }
        __syncthreads();
    }

    // exact fp32 scores of the candidates: warp w handles candidates w, w+8
    const float* q = img + static_cast<int64_t>(f) * D;
    for (int c = warp; c < ncand; c += TK_THREADS / 32) {
        const int idx = cand_i[c];
        float acc = 0.f;
        if (idx >= 0) {
            const float* p = bank + static_cast<int64_t>(idx) * D;
            for (int d = lane; d < D; d += 32) acc = fmaf(q[d], p[d], acc);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) cand_v[c] = (idx >= 0) ? acc : -FLT_MAX;
    }
    __syncthreads();

    // final ordering of <= 16 candidates by one warp: rank = number of candidates that beat this one
    if (warp == 0 && lane < ncand) {
        const float v = cand_v[lane];
        const int i = cand_i[lane];
        int rank = 0;
        for (int j = 0; j < ncand; ++j)
            if (j != lane && better(cand_v[j], cand_i[j], v, i)) ++rank;
        if (rank < k) {
            out_scores[static_cast<int64_t>(f) * k + rank] = v;
            out_idx[static_cast<int64_t>(f) * k + rank] = i;
        }
    }
}

}  // namespace

int topk_rerank_run(const float* scores, int64_t ld_scores, const float* img, const float* bank, int F, int T, int D,
                    int k, float* out_scores, int32_t* out_idx, cudaStream_t stream) {
    if (F <= 0) return 0;
    if (k < 1 || k > NCAND - 4 || k > T) {
        set_error("sim_topk: k=%d must be in [1, min(%d, T=%d)]", k, NCAND - 4, T);
        return 1;
    }
    const size_t smem = static_cast<size_t>(T) * sizeof(float);
    if (smem > 200 * 1024) {
        set_error("sim_topk: phrase bank of %d rows exceeds the %d-row shared-memory limit of this kernel", T,
                  200 * 1024 / 4);
        return 1;
    }
    // per launch (cheap): the attribute belongs to the current device's copy of the function
    VIDIL_CUDA_OK(cudaFuncSetAttribute(topk_rerank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    topk_rerank_kernel<<<F, TK_THREADS, smem, stream>>>(scores, ld_scores, img, bank, T, D, k, out_scores, out_idx);
    VIDIL_CUDA_OK(cudaGetLastError());
    count_launches(1);
    return 0;
}

}  // namespace vidil
