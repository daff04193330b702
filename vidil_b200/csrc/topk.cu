// Per-frame top-k phrase selection with exact fp32 re-ranking.
// Reference: run_visual_tokenization.py:276 (sims = image_embeds @ text_embeds.t(), fp32) and :306
// (inds = np.argsort(frm_score)[::-1][:topk]) — the reference ships the whole [F,T] fp32 matrix to the
// host and argsorts every row there.  Here the [F,T] matrix does not exist: see topk_select_kernel below.
#include <float.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

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
// Exact top-k from the similarity GEMM's EPI_TOP4 output (the four best index-tagged approximate scores of every 32-phrase
// group): one warp per frame.  With eps = 2^-10 |img| max|bank row| bounding |exact - approximate| (fp16 rounding of both
// operands, Cauchy-Schwarz; the fp32 accumulation error and the 2^-18 index tag are far below) and s_k the k-th largest
// approximate score, every phrase of the exact top k has an approximate score >= s_k - 2 eps.  So: find s_k, re-score in fp32
// from the original embeddings every emitted entry above that bar (ONE arithmetic for every exact score: lane l sums elements
// 4l..4l+3, 4l+128.., ... in order, then a butterfly — equal rows give equal scores), and — because a phrase that was never
// emitted is only known to lie below its group's fourth entry — the whole group wherever the FOURTH entry is above the bar.
// The k best exact scores are reported (score descending, index descending on exact ties — what np.argsort(...)[::-1] gives
// for the stable sort's tie order).  The indices are therefore those of the fp32 ranking whatever the data; clusters of
// near-synonyms only cost more re-scoring.
//
// What the kernel costs is latency, not bandwidth: per frame 5 KB of tagged scores, then a handful of 3 KB bank rows at
// data-dependent addresses, and one warp's dependent instruction chain.  Measured on the way here (2 048 frames x 10 000
// phrases x 768, ncu and a per-warp %globaltimer trace, VIDIL_SEL_TRACE / tools/sel_trace.py):
//   * rows fetched through registers serialise into several round trips per batch (47 % long-scoreboard stalls, 87 us): rows
//     and the tagged scores now go global -> shared by cp.async, 16 bytes per lane and instruction, nothing held in registers,
//     a whole batch one round trip.  (cp.async.bulk, one instruction per row, measured no better.)
//   * a lone warp runs this code at CPI ~7 and a one-wave kernel is as slow as its slowest warp: s_k comes from per-lane
//     sorted heads (one pass over the pool) and k rounds of one redux.max + one ballot instead of k destructive
//     argmax-and-rescan rounds (kept as the fallback for a lane that owns more than SEL_LOCAL of the k best); the candidate
//     scan is one branch-free pass per lane into a bit mask and the warp only visits slices that hit; scores that cannot
//     enter the top k skip the sorted insertion; the embedding lives in registers; the slab is 15 KB, so 14 warps per SM hold
//     2 048 frames in one wave.
//   * with TWO entries per group 4 % of the frames had to re-score a whole group (32 rows, 7 us of one warp's time; a shared
//     slab under a lock made the unluckiest CTA serialise four of them: 45 us).  The epilogue now keeps FOUR.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int SEL_MAX_K = 12;
constexpr int SEL_PER_GROUP = 4; // tagged entries per 32-phrase group (EPI_TOP4)
constexpr int SEL_BATCH = 3;     // candidates re-scored per round: their bank rows are fetched together
constexpr int SEL_CAND = 64;     // listed candidate columns per collect / drain round
constexpr int SEL_GROUPS = 32;   // listed whole groups per round
constexpr int SEL_SCRATCH = SEL_CAND + SEL_GROUPS;  // ints of per-warp scratch (>= 2 * SEL_MAX_K for the s_k fallback)
constexpr int SEL_QV_MAX = 10;   // 16-byte pieces of the frame's embedding a lane keeps in registers (template QV): D <= 1280
constexpr int SEL_WARPS_MAX = 14;
constexpr int SEL_LOCAL = 4;     // sorted per-lane heads kept while looking for s_k

__device__ __forceinline__ uint32_t ordered_bits(float v) {   // monotonic float -> uint (any non-NaN, -inf included)
    const uint32_t u = __float_as_uint(v);
    return u ^ ((u >> 31) ? 0xffffffffu : 0x80000000u);
}

// Warp-replicated list of the k best (score, index) so far, best first, in registers: every index below is a compile-time
// constant after unrolling.  Capacity is made k at run time without a dynamic index: the first SEL_MAX_K - k slots hold
// sentinels that beat everything, so the k real entries sit in the LAST k slots and the list's last slot is always the
// current k-th best — the one a new score has to beat to matter at all.
struct TopList {
    float v[SEL_MAX_K];
    int i[SEL_MAX_K];
    __device__ __forceinline__ void clear(int k) {
#pragma unroll
        for (int s = 0; s < SEL_MAX_K; ++s) {
            const bool sentinel = s < SEL_MAX_K - k;
            v[s] = sentinel ? INFINITY : -FLT_MAX;
            i[s] = sentinel ? 0x7fffffff : -1;
        }
    }
    __device__ __forceinline__ bool contains(int idx) const {
        bool hit = false;
#pragma unroll
        for (int s = 0; s < SEL_MAX_K; ++s) hit |= (i[s] == idx);
        return hit;
    }
    // can (nv, ni) still be among the k best?
    __device__ __forceinline__ bool may_enter(float nv, int ni) const { return better(nv, ni, v[SEL_MAX_K - 1], i[SEL_MAX_K - 1]); }
    __device__ __forceinline__ void insert(float nv, int ni) {   // caller checked may_enter and !contains
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

// Per-warp shared-memory slab (floats): [NE] tagged scores | [SEL_SCRATCH] | [SEL_BATCH][D] rows, NE = 4 G
__host__ __device__ inline size_t sel_slab_floats(int G, int D) {
    return static_cast<size_t>(SEL_PER_GROUP) * G + SEL_SCRATCH + static_cast<size_t>(SEL_BATCH) * D;
}

template <int SEL_QV>
__global__ void __launch_bounds__(SEL_WARPS_MAX * 32, 1)
    topk_select_kernel(const float* __restrict__ top4, int ld_top4, int G, const float* __restrict__ img,
                       const float* __restrict__ bank, float eps_scale, const float* __restrict__ bank_max_norm, int F, int T, int D,
                       int k, float* __restrict__ out_scores, int32_t* __restrict__ out_idx, unsigned long long* __restrict__ trace) {
    extern __shared__ __align__(16) float sel_smem[];
    unsigned long long t0 = 0, t1 = 0, t2 = 0;
    int n_flush = 0, n_group = 0;
    if (trace) t0 = ptx::globaltimer();
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = blockIdx.x * warps + warp;
    if (f >= F) return;
    const int NE = SEL_PER_GROUP * G;        // pool entries; a multiple of 4, so everything behind it stays 16-byte aligned
    const int D4 = D >> 2;
    float* pool = sel_smem + static_cast<size_t>(warp) * sel_slab_floats(G, D);   // [NE] tagged scores of the frame
    float* scratch = pool + NE;              // [SEL_SCRATCH]
    float* rows = scratch + SEL_SCRATCH;     // [SEL_BATCH][D] bank rows of the current batch
    {   // the frame's tagged scores, 16 bytes per lane and copy
        const float* src = top4 + static_cast<int64_t>(f) * ld_top4;
        for (int i4 = lane; i4 < G; i4 += 32) ptx::cp_async_16(pool + 4 * i4, src + 4 * i4);
        ptx::cp_async_commit();
    }
    // the embedding: lane l keeps elements 4 (l + 32 j) .. + 3, the pieces the scoring below pairs with the same pieces of a row
    float4 qv[SEL_QV];
    {
        const float4* q4 = reinterpret_cast<const float4*>(img + static_cast<int64_t>(f) * D);
#pragma unroll
        for (int j = 0; j < SEL_QV; ++j) qv[j] = (lane + 32 * j < D4) ? __ldg(q4 + lane + 32 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float nrm = 0.f;
#pragma unroll
    for (int j = 0; j < SEL_QV; ++j) nrm = fmaf(qv[j].x, qv[j].x, fmaf(qv[j].y, qv[j].y, fmaf(qv[j].z, qv[j].z, fmaf(qv[j].w, qv[j].w, nrm))));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
    // |exact - approx| <= eps: operand rounding (2 * 2^-11 relative, Cauchy-Schwarz) + the index tag (32 ulp of the score)
    const float bound = sqrtf(nrm) * bank_max_norm[0];
    const float eps = eps_scale * bound + 8e-6f * bound;
    ptx::cp_async_wait<0>();
    __syncwarp();
    if (trace) t1 = ptx::globaltimer();

    // ---- 1. s_k = the k-th largest tagged score of the pool (with multiplicity; -inf entries are padding).  Every lane keeps
    //         the SEL_LOCAL best of its own strided share, sorted; a round is one redux.max over the heads and a pop by the
    //         lowest winning lane.  A lane asked for more than SEL_LOCAL entries cannot answer from its heads: then (rare —
    //         one lane owning five of the k best) the destructive extraction below recomputes s_k from the pool. ----
    //         (A lane's share is its slot of every 32-entry slice ROTATED by the slice number: the group maxima, where the k
    //         best live, sit at every fourth pool entry and would otherwise all belong to the same eight lanes.)
    float h0 = -INFINITY, h1 = -INFINITY, h2 = -INFINITY, h3 = -INFINITY;
    int mine_cnt = 0;
#pragma unroll 4
    for (int j = 0; (j << 5) < NE; ++j) {
        const int i = (j << 5) + ((lane + j) & 31);
        if (i >= NE) continue;
        const float v = pool[i];
        ++mine_cnt;
        if (v > h3) {
            h3 = v;
            if (h3 > h2) { const float t = h2; h2 = h3; h3 = t; }
            if (h2 > h1) { const float t = h1; h1 = h2; h2 = t; }
            if (h1 > h0) { const float t = h0; h0 = h1; h1 = t; }
        }
    }
    float s_k = -INFINITY;
    bool exhausted = false;
    {
        int pops = 0;
        const uint32_t neg_inf = ordered_bits(-INFINITY);
#pragma unroll 1
        for (int r = 0; r < k; ++r) {
            const uint32_t key = ordered_bits(h0);
            const uint32_t m = __reduce_max_sync(0xffffffffu, key);
            if (m <= neg_inf) break;   // nothing but padding left
            const int winner = __ffs(__ballot_sync(0xffffffffu, key == m)) - 1;
            s_k = __shfl_sync(0xffffffffu, h0, winner);
            if (lane == winner) {
                h0 = h1; h1 = h2; h2 = h3; h3 = -INFINITY;
                if (++pops == SEL_LOCAL && mine_cnt > SEL_LOCAL && r + 1 < k) exhausted = true;
            }
        }
    }
    if (__any_sync(0xffffffffu, exhausted)) {
        auto local_best = [&]() {
            Best b{-INFINITY, -1};
            for (int i = lane; i < NE; i += 32) {
                const float v = pool[i];
                if (better(v, i, b.v, b.i)) {
                    b.v = v;
                    b.i = i;
                }
            }
            return b;
        };
        int* taken = reinterpret_cast<int*>(scratch);   // [SEL_MAX_K] entries taken out while looking for s_k (restored below)
        float* taken_v = scratch + SEL_MAX_K;
        Best mine = local_best();
        s_k = -INFINITY;
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
    }
    // ---- 2. candidates.  COLLECT: one branch-free pass per lane marks, per 32-entry slice of the pool, whether its entry is
    //         above the bar; the warp visits only the slices somebody marked and appends entries 0..2 of a group to a column
    //         list, a group whose entry 3 qualifies to a group list.  DRAIN: ONE piece of code fetches the next SEL_BATCH
    //         columns — listed ones first, then the members of a listed group — scores them and inserts what can still enter
    //         the top k. ----
    const float bar = s_k - 2.0f * eps;
    if (trace) t2 = ptx::globaltimer();
    TopList top;
    top.clear(k);
    int* cand = reinterpret_cast<int*>(scratch);          // [SEL_CAND] columns waiting to be re-scored
    int* groups = cand + SEL_CAND;                        // [SEL_GROUPS] groups waiting to be re-scored whole
#pragma unroll 1
    for (int base = 0; base < NE; base += 1024) {
        const int nit = min(32, (NE - base + 31) >> 5);
        uint32_t hm = 0;
#pragma unroll 4
        for (int it = 0; it < nit; ++it) {
            const int i = base + (it << 5) + lane;
            const float v = (i < NE) ? pool[i] : -INFINITY;
            hm |= (v >= bar && v != -INFINITY) ? (1u << it) : 0u;
        }
        uint32_t any = __reduce_or_sync(0xffffffffu, hm);
#pragma unroll 1
        while (any) {
            int n_cand = 0, n_groups = 0;
#pragma unroll 1
            while (any && n_cand + 24 <= SEL_CAND && n_groups + 8 <= SEL_GROUPS) {   // a slice adds at most 24 columns, 8 groups
                const int it = __ffs(any) - 1;
                any &= any - 1;
                const int i = base + (it << 5) + lane;
                const bool hit = (hm >> it) & 1u;
                const bool single = hit && (i & 3) != 3, whole = hit && (i & 3) == 3;
                const unsigned fm = __ballot_sync(0xffffffffu, single), wm = __ballot_sync(0xffffffffu, whole);
                const unsigned below = (1u << lane) - 1u;
                if (single) cand[n_cand + __popc(fm & below)] = (i >> 2) * 32 + static_cast<int>(__float_as_uint(pool[i]) & 31u);
                if (whole) groups[n_groups + __popc(wm & below)] = i >> 2;
                n_cand += __popc(fm);
                n_groups += __popc(wm);
            }
            __syncwarp();
            int pos = 0, gi = 0, goff = 0;
#pragma unroll 1
            while (pos < n_cand || gi < n_groups) {
                int n, my_col;
                if (pos < n_cand) {
                    n = min(SEL_BATCH, n_cand - pos);
                    my_col = cand[pos + min(lane, n - 1)];
                    pos += n;
                } else {   // members goff.. of listed group gi (its entries 0..3 are re-scored a second time: contains() drops them)
                    const int g = groups[gi];
                    const int members = min(32, T - g * 32);
                    n = min(SEL_BATCH, members - goff);
                    my_col = g * 32 + goff + min(lane, n - 1);
                    goff += n;
                    if (goff == members) {
                        goff = 0;
                        ++gi;
                        ++n_group;
                    }
                }
                ++n_flush;
                // the rows: every lane copies its 16-byte pieces of every row, all in flight at once
#pragma unroll
                for (int r = 0; r < SEL_BATCH; ++r) {
                    if (r < n) {   // warp-uniform
                        const float* src = bank + static_cast<int64_t>(__shfl_sync(0xffffffffu, my_col, r)) * D;
                        float* d = rows + static_cast<size_t>(r) * D;
#pragma unroll
                        for (int j = 0; j < SEL_QV; ++j)
                            if (lane + 32 * j < D4) ptx::cp_async_16(d + 4 * (lane + 32 * j), src + 4 * (lane + 32 * j));
                    }
                }
                ptx::cp_async_commit();
                ptx::cp_async_wait<0>();
                __syncwarp();
                float acc[SEL_BATCH];
#pragma unroll
                for (int c = 0; c < SEL_BATCH; ++c) acc[c] = 0.f;
                const float4* r4 = reinterpret_cast<const float4*>(rows);
#pragma unroll
                for (int j = 0; j < SEL_QV; ++j) {
                    if (lane + 32 * j < D4) {
#pragma unroll
                        for (int c = 0; c < SEL_BATCH; ++c) {
                            if (c < n) {   // warp-uniform; slots past the batch hold stale bytes
                                const float4 b = r4[c * D4 + lane + 32 * j];
                                acc[c] = fmaf(qv[j].x, b.x, acc[c]);
                                acc[c] = fmaf(qv[j].y, b.y, acc[c]);
                                acc[c] = fmaf(qv[j].z, b.z, acc[c]);
                                acc[c] = fmaf(qv[j].w, b.w, acc[c]);
                            }
                        }
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                    for (int c = 0; c < SEL_BATCH; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
                }
#pragma unroll 1
                for (int c = 0; c < n; ++c) {   // rolled: one copy of the sorted insertion
                    float sc = acc[0];
#pragma unroll
                    for (int u = 1; u < SEL_BATCH; ++u) sc = (c == u) ? acc[u] : sc;
                    const int cc = __shfl_sync(0xffffffffu, my_col, c);
                    if (top.may_enter(sc, cc) && !top.contains(cc)) top.insert(sc, cc);
                }
                __syncwarp();   // every lane is done with the rows before the next fetch overwrites them
            }
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < SEL_MAX_K; ++s) {
            const int o = s - (SEL_MAX_K - k);   // the k real entries are the last k slots, best first
            if (o >= 0) {
                out_scores[static_cast<int64_t>(f) * k + o] = top.v[s];
                out_idx[static_cast<int64_t>(f) * k + o] = top.i[s];
            }
        }
        if (trace) {   // developer trace (VIDIL_SEL_TRACE): per frame start / pool landed / s_k known / done (ns), SM, batches, groups
            unsigned long long* tr = trace + static_cast<size_t>(f) * 8;
            tr[0] = t0; tr[1] = t1; tr[2] = t2; tr[3] = ptx::globaltimer();
            tr[4] = ptx::smid(); tr[5] = n_flush; tr[6] = n_group; tr[7] = blockIdx.x;
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

int topk_select_run(const float* top4, int ld_top4, int G, const float* img, const float* bank, const float* bank_max_norm_dev, int F, int T,
                    int D, int k, float* out_scores, int32_t* out_idx, cudaStream_t stream) {
    if (F <= 0) return 0;
    if (k < 1 || k > SEL_MAX_K || k > T) {
        set_error("sim_topk: k=%d must be in [1, min(%d, T=%d)]", k, SEL_MAX_K, T);
        return 1;
    }
    if ((reinterpret_cast<uintptr_t>(img) & 15) || (reinterpret_cast<uintptr_t>(bank) & 15) || (reinterpret_cast<uintptr_t>(top4) & 15) ||
        D % 4 != 0 || ld_top4 % 4 != 0 || ld_top4 < SEL_PER_GROUP * G) {
        set_error("sim_topk: embeddings and phrase bank must be 16-byte aligned fp32 rows (16-byte asynchronous copies)");
        return 1;
    }
    if (D > 128 * SEL_QV_MAX) {
        set_error("sim_topk: embedding width %d exceeds %d (the selection kernel keeps the frame's embedding in registers)", D, 128 * SEL_QV_MAX);
        return 1;
    }
    const size_t per_warp = sel_slab_floats(G, D) * sizeof(float);
    // one CTA per SM with as many warps as fit (227 KB per CTA): 14 for 10 000 phrases x 768, which holds 2 048 frames on 148
    // SMs in a single wave
    constexpr size_t SEL_SMEM_MAX = 227 * 1024;
    int warps = static_cast<int>(SEL_SMEM_MAX / per_warp);
    if (warps > SEL_WARPS_MAX) warps = SEL_WARPS_MAX;
    if (warps < 1) {
        set_error("sim_topk: a phrase bank of %d rows x %d exceeds the shared-memory pool of the selection kernel", T, D);
        return 1;
    }
    // small calls: fewer warps per CTA spread the frames over more SMs
    while (warps > 1 && (F + warps - 2) / (warps - 1) <= gemm_num_sms()) --warps;
    const int grid = (F + warps - 1) / warps;
    const size_t smem = per_warp * warps;
    const int qv = (D / 4 + 31) / 32;   // 16-byte pieces per lane
    // developer trace: VIDIL_SEL_TRACE=<file> dumps one line per frame of the LAST call (see the kernel's epilogue)
    static const char* trace_path = getenv("VIDIL_SEL_TRACE");
    unsigned long long* trace = nullptr;
    if (trace_path) VIDIL_CUDA_OK(cudaMalloc(&trace, static_cast<size_t>(F) * 8 * sizeof(unsigned long long)));
    auto go = [&](auto kern) -> int {
        // per launch (cheap): the attribute belongs to the current device's copy of the function
        VIDIL_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(SEL_SMEM_MAX)));
        kern<<<grid, warps * 32, smem, stream>>>(top4, ld_top4, G, img, bank, 1.0f / 1024.0f, bank_max_norm_dev, F, T, D, k, out_scores, out_idx, trace);
        return 0;
    };
    int rc;
    if (qv <= 2) rc = go(topk_select_kernel<2>);
    else if (qv <= 4) rc = go(topk_select_kernel<4>);
    else if (qv <= 6) rc = go(topk_select_kernel<6>);
    else if (qv <= 8) rc = go(topk_select_kernel<8>);
    else rc = go(topk_select_kernel<SEL_QV_MAX>);
    if (rc) return rc;
    VIDIL_CUDA_OK(cudaGetLastError());
    count_launches(1);
    if (trace) {
        std::vector<unsigned long long> h(static_cast<size_t>(F) * 8);
        VIDIL_CUDA_OK(cudaStreamSynchronize(stream));
        VIDIL_CUDA_OK(cudaMemcpy(h.data(), trace, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        cudaFree(trace);
        if (FILE* fp = fopen(trace_path, "w")) {
            for (int f = 0; f < F; ++f) {
                const unsigned long long* t = &h[static_cast<size_t>(f) * 8];
                fprintf(fp, "%d %llu %llu %llu %llu %llu %llu %llu %llu\n", f, t[0], t[1], t[2], t[3], t[4], t[5], t[6], t[7]);
            }
            fclose(fp);
        }
    }
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
