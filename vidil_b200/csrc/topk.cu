// Per-frame top-k phrase selection with exact fp32 re-ranking.
// Reference: run_visual_tokenization.py:276 (sims = image_embeds @ text_embeds.t(), fp32) and :306
// (inds = np.argsort(frm_score)[::-1][:topk]) — the reference ships the whole [F,T] fp32 matrix to the
// host and argsorts every row there.  Here the [F,T] matrix does not exist: see topk_select_kernel below.
#include <float.h>

#include "kernels.h"
#include "ptx.cuh"

namespace vidil {
namespace {


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

// ---------------------------------------------------------------------------------------------------------------------
// Exact top-k from the similarity GEMM's EPI_TOP2 output (two index-tagged approximate scores per 32-phrase group): one warp
// per frame.  With eps = 2^-10 |img| max|bank row| bounding |exact - approximate| (fp16 rounding of both operands,
// Cauchy-Schwarz; the fp32 accumulation error and the 2^-18 index tag are far below) and s_k the k-th largest approximate
// score, every phrase of the exact top k has an approximate score >= s_k - 2 eps.  So: find s_k (k extractions), re-score in
// fp32 from the original embeddings every emitted entry above that bar (fixed summation order: lane l sums elements l,
// l + 32, ..., then a butterfly), and — because a phrase that was never emitted is only known to lie below its group's
// second entry — the whole group wherever the SECOND entry is above the bar.  The k best exact scores are reported (score
// descending, index descending on exact ties — what np.argsort(...)[::-1] gives for the stable sort's tie order).  The
// indices are therefore those of the fp32 ranking whatever the data; clusters of near-synonyms only cost more re-scoring.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int SEL_MAX_K = 12;
constexpr int SEL_BATCH = 8;   // candidates re-scored per round (their bank rows are in flight together)
constexpr int SEL_SCRATCH = 32;  // floats of per-warp scratch behind the embedding (>= 2 * SEL_MAX_K, >= 2 * SEL_BATCH)

// Exact fp32 dot products of the frame's embedding q (shared memory) with up to SEL_BATCH bank rows (col[c] < 0: skip):
// lane l sums elements 4l..4l+3, 4l+128.., ... in order, then a butterfly — ONE arithmetic for every exact score, so equal
// rows give equal scores.  16-byte loads, d outer / candidate inner, the d loop unrolled by 3: 24 row segments of 16 bytes
// in flight per lane (the kernel is bound by the latency of these loads: 47 % of its stall samples).
__device__ __forceinline__ void score_batch(const float* __restrict__ q, const float* __restrict__ bank, int D, int lane,
                                            const int (&col)[SEL_BATCH], float (&acc)[SEL_BATCH]) {
    const float4* p[SEL_BATCH];
#pragma unroll
    for (int c = 0; c < SEL_BATCH; ++c) {
        p[c] = reinterpret_cast<const float4*>(bank + static_cast<int64_t>(col[c] < 0 ? 0 : col[c]) * D);
        acc[c] = 0.f;
    }
    const float4* q4 = reinterpret_cast<const float4*>(q);
#pragma unroll 3
    for (int d4 = lane; d4 < (D >> 2); d4 += 32) {
        const float4 qd = q4[d4];
#pragma unroll
        for (int c = 0; c < SEL_BATCH; ++c) {
            if (col[c] >= 0) {
                const float4 b = __ldg(p[c] + d4);
                acc[c] = fmaf(qd.x, b.x, acc[c]);
                acc[c] = fmaf(qd.y, b.y, acc[c]);
                acc[c] = fmaf(qd.z, b.z, acc[c]);
                acc[c] = fmaf(qd.w, b.w, acc[c]);
            }
        }
    }
#pragma unroll
    for (int c = 0; c < SEL_BATCH; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
    }
}

// Warp-replicated list of the SEL_MAX_K best (score, index) so far, best first; every index below is a compile-time constant
// after unrolling, so the list lives in registers.
struct TopList {
    float v[SEL_MAX_K];
    int i[SEL_MAX_K];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int s = 0; s < SEL_MAX_K; ++s) {
            v[s] = -FLT_MAX;
            i[s] = -1;
        }
    }
    __device__ __forceinline__ bool contains(int idx) const {
        bool hit = false;
#pragma unroll
        for (int s = 0; s < SEL_MAX_K; ++s) hit |= (i[s] == idx);
        return hit;
    }
    __device__ __forceinline__ float kth(int k) const {  // score of the k-th best (k runtime, 1-based)
        float r = -FLT_MAX;
#pragma unroll
        for (int s = 0; s < SEL_MAX_K; ++s) r = (s == k - 1) ? v[s] : r;
        return r;
    }
    __device__ __forceinline__ void insert(float nv, int ni) {
        if (!better(nv, ni, v[SEL_MAX_K - 1], i[SEL_MAX_K - 1])) return;
        v[SEL_MAX_K - 1] = nv;
        i[SEL_MAX_K - 1] = ni;
#pragma unroll
        for (int s = SEL_MAX_K - 1; s > 0; --s) {  // bubble up
            if (better(v[s], i[s], v[s - 1], i[s - 1])) {
                const float tv = v[s];
                const int ti = i[s];
                v[s] = v[s - 1];
                i[s] = i[s - 1];
                v[s - 1] = tv;
                i[s - 1] = ti;
            }
        }
    }
};

__global__ void __launch_bounds__(128, 4)
    topk_select_kernel(const float* __restrict__ top2, int ld_top2, int G, const float* __restrict__ img,
                       const float* __restrict__ bank, float eps_scale, const float* __restrict__ bank_max_norm, int F, int T, int D,
                       int k, float* __restrict__ out_scores, int32_t* __restrict__ out_idx) {
    extern __shared__ float sel_smem[];
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = blockIdx.x * warps + warp;
    if (f >= F) return;
    const int pool_len = (2 * G + 3) & ~3;   // keeps the embedding behind it 16-byte aligned
    float* pool = sel_smem + static_cast<size_t>(warp) * (pool_len + D + SEL_SCRATCH);   // [2 G] tagged scores, the frame's embedding [D], batch scratch
    float* q = pool + pool_len;
    const float* src = top2 + static_cast<int64_t>(f) * ld_top2;
    for (int i = lane; i < 2 * G; i += 32) pool[i] = src[i];
    float nrm = 0.f;
    for (int d = lane; d < D; d += 32) {
        const float x = img[static_cast<int64_t>(f) * D + d];
        q[d] = x;
        nrm = fmaf(x, x, nrm);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
    __syncwarp();
    // |exact - approx| <= eps: operand rounding (2 * 2^-11 relative, Cauchy-Schwarz) + the index tag (32 ulp of the score)
    const float bound = sqrtf(nrm) * bank_max_norm[0];
    const float eps = eps_scale * bound + 8e-6f * bound;

    // ---- 1. s_k = the k-th largest tagged score of the pool (k cheap extractions: every lane keeps the best of its own
    //         strided share, the warp takes the best of those, only the winning lane re-scans its share) ----
    auto local_best = [&]() {
        Best b{-INFINITY, -1};
        for (int i = lane; i < 2 * G; i += 32) {
            const float v = pool[i];
            if (better(v, i, b.v, b.i)) {
                b.v = v;
                b.i = i;
            }
        }
        return b;
    };
    int* taken = reinterpret_cast<int*>(q + D);   // [SEL_MAX_K] entries taken out while looking for s_k (restored below)
    float* taken_v = q + D + SEL_MAX_K;
    Best mine = local_best();
    float s_k = -INFINITY;
    int n_taken = 0;
#pragma unroll 1
    for (int r = 0; r < k; ++r) {
        const Best w = warp_argmax(mine);
        if (w.i < 0 || w.v == -INFINITY) break;
        s_k = w.v;
        if (w.i == mine.i) {   // exactly one lane owns entry w.i
            taken[r] = w.i;
            taken_v[r] = w.v;
            pool[w.i] = -INFINITY;
            mine = local_best();
        }
        ++n_taken;
    }
    __syncwarp();
    if (lane < n_taken) pool[taken[lane]] = taken_v[lane];
    __syncwarp();
    // ---- 2. candidates: a phrase of the exact top k has an exact score >= the k-th exact score >= s_k - eps, hence a tagged
    //         approximate score >= s_k - 2 eps.  Every pool entry above that bar is re-scored; if a group's SECOND entry is
    //         above it, the group's other members (only known to lie below that entry) are re-scored too. ----
    const float bar = s_k - 2.0f * eps;
    TopList top;
    top.clear();
    int* ccol = reinterpret_cast<int*>(q + D);     // [SEL_BATCH] columns of the current batch
    float* cval = q + D + SEL_BATCH;               // [SEL_BATCH] their exact scores
    int n_batch = 0;
    auto flush = [&]() {
        // exact scores of up to SEL_BATCH columns (rows in flight together), then sorted insertion
        __syncwarp();
        int col[SEL_BATCH];
        float acc[SEL_BATCH];
#pragma unroll
        for (int c = 0; c < SEL_BATCH; ++c) col[c] = (c < n_batch) ? ccol[c] : -1;
        score_batch(q, bank, D, lane, col, acc);
        __syncwarp();
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < SEL_BATCH; ++c) cval[c] = acc[c];
        }
        __syncwarp();
#pragma unroll 1
        for (int c = 0; c < n_batch; ++c) {
            const int cc = ccol[c];
            if (!top.contains(cc)) top.insert(cval[c], cc);
        }
        __syncwarp();
        n_batch = 0;
    };
    auto push = [&](bool want, int colv) {   // warp-wide: lanes with `want` append their column, flushing full batches
        unsigned m = __ballot_sync(0xffffffffu, want);
        while (m) {
            const int room = SEL_BATCH - n_batch;
            const int rank = __popc(m & ((1u << lane) - 1u));
            const bool now = want && rank < room;
            if (now) ccol[n_batch + rank] = colv;
            const int added = min(room, __popc(m));
            n_batch += added;
            // drop the lanes that were served
            if (now) want = false;
            m = __ballot_sync(0xffffffffu, want);
            if (n_batch == SEL_BATCH) flush();
        }
    };
#pragma unroll 1
    for (int i0 = 0; i0 < 2 * G; i0 += 32) {
        const int i = i0 + lane;
        const float v = (i < 2 * G) ? pool[i] : -INFINITY;
        const bool hit = v >= bar && v != -INFINITY;
        if (!__any_sync(0xffffffffu, hit)) continue;   // the usual case: nothing of these 32 entries is near the top
        push(hit && !(i & 1), (i >> 1) * 32 + static_cast<int>(__float_as_uint(v) & 31u));
        // second entries above the bar: the whole group (rare)
        unsigned gm = __ballot_sync(0xffffffffu, hit && (i & 1));
        while (gm) {
            const int src = __ffs(gm) - 1;
            gm &= gm - 1;
            const int g = (i0 + src) >> 1;
            push(g * 32 + lane < T, g * 32 + lane);
        }
    }
    if (n_batch > 0) flush();
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < SEL_MAX_K; ++s) {
            if (s < k) {
                out_scores[static_cast<int64_t>(f) * k + s] = top.v[s];
                out_idx[static_cast<int64_t>(f) * k + s] = top.i[s];
            }
        }
    }
}

__global__ void __launch_bounds__(256) max_row_norm_kernel(const float* __restrict__ m, int rows, int D, float* __restrict__ out) {
    // one warp per row, grid-stride; a float atomicMax on non-negative values is an integer atomicMax on the bit pattern
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    float best = 0.f;
    for (int r = warp; r < rows; r += nwarps) {
        float s = 0.f;
        for (int d = lane; d < D; d += 32) {
            const float x = m[static_cast<int64_t>(r) * D + d];
            s = fmaf(x, x, s);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        best = fmaxf(best, sqrtf(s));
    }
    if (lane == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(best));
}

}  // namespace

int topk_select_run(const float* top2, int ld_top2, int G, const float* img, const float* bank, const float* bank_max_norm_dev, int F, int T,
                    int D, int k, float* out_scores, int32_t* out_idx, cudaStream_t stream) {
    if (F <= 0) return 0;
    if (k < 1 || k > SEL_MAX_K || k > T) {
        set_error("sim_topk: k=%d must be in [1, min(%d, T=%d)]", k, SEL_MAX_K, T);
        return 1;
    }
    const size_t per_warp = (((2 * static_cast<size_t>(G) + 3) & ~static_cast<size_t>(3)) + D + SEL_SCRATCH) * sizeof(float);
    int warps = 4;
    while (warps > 1 && per_warp * warps > 200 * 1024) warps >>= 1;
    if (per_warp * warps > 200 * 1024) {
        set_error("sim_topk: a phrase bank of %d rows x %d exceeds the shared-memory pool of the selection kernel", T, D);
        return 1;
    }
    // per launch (cheap): the attribute belongs to the current device's copy of the function
    VIDIL_CUDA_OK(cudaFuncSetAttribute(topk_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const int grid = (F + warps - 1) / warps;
    topk_select_kernel<<<grid, warps * 32, per_warp * warps, stream>>>(top2, ld_top2, G, img, bank, 1.0f / 1024.0f, bank_max_norm_dev, F, T, D,
                                                                      k, out_scores, out_idx);
    VIDIL_CUDA_OK(cudaGetLastError());
    count_launches(1);
    return 0;
}

int max_row_norm_run(const float* m, int rows, int D, float* out, cudaStream_t stream) {
    VIDIL_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(float), stream));
    if (rows <= 0) return 0;
    int grid = (rows + 7) / 8;
    if (grid > 1184) grid = 1184;
    max_row_norm_kernel<<<grid, 256, 0, stream>>>(m, rows, D, out);
    VIDIL_CUDA_OK(cudaGetLastError());
    count_launches(1);
    return 0;
}

}  // namespace vidil
