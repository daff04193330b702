// SIMT kernels of the med.py text stack (caption decoder / ITM encoder) around the tcgen05 GEMMs:
// token embedding, short-sequence self-attention with a beam-indexed K/V cache, cross-attention of a few query rows onto
// the image tokens, fused log-softmax + top-2K over the vocabulary, and the beam-search bookkeeping
// (transformers v4.15 BeamSearchScorer semantics, driven from models/blip.py:150-158).
//
// Shapes are small on the query side (<= 40 text tokens, 3 beams) and large on the key side of the cross-attention
// (197 / 577 image tokens per frame) and of the vocabulary scan (30 524 logits per row): all of these kernels are
// HBM/L2-bound streaming kernels with 128-bit loads; none of them is GEMM-shaped enough for the tensor cores.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math.h>

#include "kernels.h"
#include "ptx.cuh"

namespace vidil {
namespace {

template <typename T>
__device__ __forceinline__ float to_f(T x);
template <>
__device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 x) {
    return __bfloat162float(x);
}
template <>
__device__ __forceinline__ float to_f<__half>(__half x) {
    return __half2float(x);
}
template <typename T>
__device__ __forceinline__ T from_f(float x);
template <>
__device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float x) {
    return __float2bfloat16_rn(x);
}
template <>
__device__ __forceinline__ __half from_f<__half>(float x) {
    return __float2half_rn(x);
}

// 8 consecutive 16-bit values (one 128-bit load) -> 8 floats
template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&v)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const T* h = reinterpret_cast<const T*>(&u);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = to_f<T>(h[i]);
}

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}
__device__ __forceinline__ float warp_max(float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, o));
    return x;
}

// ---------------------------------------------------------------------------------------------------------------------
// BertEmbeddings (med.py:74-96) before its LayerNorm: resid[r,:] = word[ids[r],:] + pos[pos0 + r % T,:]
// ---------------------------------------------------------------------------------------------------------------------
__global__ void med_embed_kernel(const int32_t* __restrict__ ids, const float* __restrict__ word, const float* __restrict__ pos,
                                 float* __restrict__ resid, int64_t rows, int T, int pos0, int ids_mod, int D, int vocab, int max_pos) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    const int d4 = D / 4;
    const int64_t total = rows * d4;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t r = i / d4;
        const int c = static_cast<int>(i - r * d4);
        int id = ids[ids_mod > 0 ? r % ids_mod : r];  // ids_mod: the same short sequence (the prompt) for every frame
        id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
        int t = pos0 + static_cast<int>(r % T);
        t = t >= max_pos ? max_pos - 1 : t;
        const float4 a = reinterpret_cast<const float4*>(word + static_cast<int64_t>(id) * D)[c];
        const float4 b = reinterpret_cast<const float4*>(pos + static_cast<int64_t>(t) * D)[c];
        reinterpret_cast<float4*>(resid + r * D)[c] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Decode-step self-attention (BertSelfAttention with past_key_value, med.py:169-175): every row holds ONE new token at
// position `pos`; keys/values of earlier positions live in cache slots [row][t][2D] that are never moved — row r's history
// is found through anc[r][t] = the row that computed position t of this beam's prefix (replaces _reorder_cache,
// med.py:951-955).  One warp per (row, head): lane j scores key j (a full 128-byte K row per lane, q broadcast), the
// softmax is two warp reductions, then lane l accumulates dims 2l, 2l+1 of P.V with the key loop's addresses known up front.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int SA_WARPS = 8;
constexpr int SA_MAX_KEYS = 64;

// 8 lanes share one 128-byte K or V row (16 bytes each), 4 rows per warp-wide load.
template <typename T>
__global__ void __launch_bounds__(SA_WARPS * 32)
    med_self_attn_decode_kernel(const T* __restrict__ qkv, T* __restrict__ cache, const int32_t* __restrict__ anc, T* __restrict__ out,
                                int rows, int H, int pos, int Tmax, float scale) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    __shared__ float s_sc[SA_WARPS][SA_MAX_KEYS];
    __shared__ int s_src[SA_WARPS][SA_MAX_KEYS];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int item = blockIdx.x * SA_WARPS + w;
    if (item >= rows * H) return;
    const int r = item / H, h = item - r * H;
    const int D = H * 64;
    const int c = lane & 7, kq = lane >> 3;
    const T* own = qkv + static_cast<int64_t>(r) * 3 * D + h * 64;  // q | +D: k | +2D: v of the new token
    // the new token's K/V go to slot (r, pos): lanes 0..7 copy K, 8..15 copy V, 16 bytes each
    if (lane < 16) {
        T* dst = cache + (static_cast<int64_t>(r) * Tmax + pos) * (2 * D) + kq * D + h * 64 + c * 8;
        *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(own + (1 + kq) * D + c * 8);
    }
    float q[8];
    load8<T>(own + 8 * c, q);
#pragma unroll
    for (int e = 0; e < 8; ++e) q[e] *= scale;
    const int nkeys = pos + 1;
    float* sc = s_sc[w];
    int* srcs = s_src[w];
    for (int j0 = 0; j0 < nkeys; j0 += 4) {   // warp-uniform trip count: the shuffles below need every lane
        const int j = j0 + kq;
        const bool valid = j < nkeys;
        int src = r;
        const T* kp = own + D;
        if (valid && j != pos) {
            src = anc[static_cast<int64_t>(r) * Tmax + j];
            kp = cache + (static_cast<int64_t>(src) * Tmax + j) * (2 * D) + h * 64;
        }
        float k8[8];
        load8<T>(kp + 8 * c, k8);
        float d = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) d += q[e] * k8[e];
        d += __shfl_xor_sync(0xffffffffu, d, 1);
        d += __shfl_xor_sync(0xffffffffu, d, 2);
        d += __shfl_xor_sync(0xffffffffu, d, 4);
        if (valid && c == 0) {
            sc[j] = d;
            srcs[j] = src;
        }
    }
    __syncwarp();
    const float s0 = lane < nkeys ? sc[lane] : -INFINITY, s1 = lane + 32 < nkeys ? sc[lane + 32] : -INFINITY;
    const float mx = warp_max(fmaxf(s0, s1));
    const float p0 = lane < nkeys ? __expf(s0 - mx) : 0.f, p1 = lane + 32 < nkeys ? __expf(s1 - mx) : 0.f;
    const float inv = 1.0f / warp_sum(p0 + p1);
    __syncwarp();
    if (lane < nkeys) sc[lane] = p0 * inv;
    if (lane + 32 < nkeys) sc[lane + 32] = p1 * inv;
    __syncwarp();
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    for (int j = kq; j < nkeys; j += 4) {
        const T* vp = (j == pos) ? own + 2 * D : cache + (static_cast<int64_t>(srcs[j]) * Tmax + j) * (2 * D) + D + h * 64;
        float v8[8];
        load8<T>(vp + 8 * c, v8);
        const float p = sc[j];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += p * v8[e];
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 8);
        acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 16);
    }
    if (kq == 0) {
        T o8[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o8[e] = from_f<T>(acc[e]);
        *reinterpret_cast<uint4*>(out + static_cast<int64_t>(r) * D + h * 64 + c * 8) = *reinterpret_cast<const uint4*>(o8);
    }
}

// K/V of whole sequences (the prompt) -> cache slots (seq * beams, t): the decode steps of every beam of a frame start from
// the same prompt prefix.  qkv [n_seq*T, 3D]; one thread per 16 bytes.
template <typename T>
__global__ void med_cache_fill_kernel(const T* __restrict__ qkv, T* __restrict__ cache, int64_t n_rows, int T_seq, int D, int Tmax,
                                      int beams) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    const int chunks = 2 * D / 8;
    const int64_t total = n_rows * chunks;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t row = i / chunks;
        const int c = static_cast<int>(i - row * chunks);
        const int64_t b = row / T_seq;
        const int t = static_cast<int>(row - b * T_seq);
        const uint4 val = *reinterpret_cast<const uint4*>(qkv + row * 3 * D + D + c * 8);
        *reinterpret_cast<uint4*>(cache + ((b * beams) * Tmax + t) * (2 * static_cast<int64_t>(D)) + c * 8) = val;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// log_softmax over the vocabulary + MinLengthLogitsProcessor + beam score + the row's best NC candidates
// (the per-row half of `torch.topk(next_token_scores.view(batch, beams*V), 2*beams)`).  One CTA per candidate list.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int TK_THREADS = 256;
constexpr int TK_MAX_NC = 8;
constexpr int TK_CAP = 1024;  // candidate list capacity (shared memory)

__device__ __forceinline__ bool cand_better(float v, int i, float v2, int i2) { return v > v2 || (v == v2 && i < i2); }

// Best (value, index) of the block under cand_better; every thread gets the winner.  Two __syncthreads per call.
__device__ __forceinline__ void block_argmax(float& v, int& i, float* s_v, int* s_i) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float v2 = __shfl_xor_sync(0xffffffffu, v, o);
        const int i2 = __shfl_xor_sync(0xffffffffu, i, o);
        if (cand_better(v2, i2, v, i)) {
            v = v2;
            i = i2;
        }
    }
    __syncthreads();  // the previous call's readers are done with s_v / s_i
    if (lane == 0) {
        s_v[w] = v;
        s_i[w] = i;
    }
    __syncthreads();
    v = s_v[0];
    i = s_i[0];
#pragma unroll
    for (int k = 1; k < TK_THREADS / 32; ++k)
        if (cand_better(s_v[k], s_i[k], v, i)) {
            v = s_v[k];
            i = s_i[k];
        }
}

// Pass 1 streams the row once for the log-sum-exp and each thread's largest admissible logit; the nc-th largest of those
// 256 thread maxima is a lower bound of the row's nc-th largest value, so pass 2 (the row is in L2 by then) only has to
// collect the handful of elements at or above it, from which the block picks the nc best exactly — ties to the lowest token
// id, like a stable descending sort.  Keeping a sorted top-nc list per thread instead costs ~10x more: with 120 elements per
// thread a quarter of them trigger an insertion, and a warp diverges on almost every element.
__global__ void __launch_bounds__(TK_THREADS)
    med_logits_topk_kernel(const float* __restrict__ logits, int64_t ld, int row_mul, const float* __restrict__ beam_scores, int V,
                           int nc, int ban_token, float* __restrict__ cand_score, int32_t* __restrict__ cand_tok) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    __shared__ float s_red[TK_THREADS / 32];
    __shared__ float s_v[TK_THREADS / 32];
    __shared__ int s_i[TK_THREADS / 32];
    __shared__ float s_m, s_l;
    __shared__ int s_cnt;
    __shared__ float c_v[TK_CAP];
    __shared__ int c_i[TK_CAP];
    const int list = blockIdx.x, t = threadIdx.x, lane = t & 31, w = t >> 5;
    const float* row = logits + static_cast<int64_t>(list) * row_mul * ld;
    const float4* row4 = reinterpret_cast<const float4*>(row);
    const int v4 = V / 4;
    constexpr int TU = 4;  // independent 16-byte loads in flight per thread
    float m = -INFINITY, s = 0.f, cm = -INFINITY;
    if (t == 0) s_cnt = 0;
    for (int base = t; base < v4; base += TK_THREADS * TU) {
        float4 x4s[TU];
#pragma unroll
        for (int u = 0; u < TU; ++u) {
            const int i4 = base + u * TK_THREADS;
            x4s[u] = (i4 < v4) ? row4[i4] : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        }
#pragma unroll
        for (int u = 0; u < TU; ++u) {
            const int i4 = base + u * TK_THREADS;
            if (i4 < v4) {
                const float xs[4] = {x4s[u].x, x4s[u].y, x4s[u].z, x4s[u].w};
                const float bm = fmaxf(fmaxf(xs[0], xs[1]), fmaxf(xs[2], xs[3]));
                if (bm > m) {
                    s *= __expf(m - bm);
                    m = bm;
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    s += __expf(xs[e] - m);
                    if (4 * i4 + e != ban_token) cm = fmaxf(cm, xs[e]);
                }
            }
        }
    }
    // block max, then rescaled sum -> log-sum-exp
    const float wm = warp_max(m);
    if (lane == 0) s_red[w] = wm;
    __syncthreads();
    if (t == 0) {
        float x = s_red[0];
        for (int i = 1; i < TK_THREADS / 32; ++i) x = fmaxf(x, s_red[i]);
        s_m = x;
    }
    __syncthreads();
    const float gm = s_m;
    const float ws = warp_sum(m == -INFINITY ? 0.f : s * __expf(m - gm));
    __syncthreads();
    if (lane == 0) s_red[w] = ws;
    __syncthreads();
    if (t == 0) {
        float x = 0.f;
        for (int i = 0; i < TK_THREADS / 32; ++i) x += s_red[i];
        s_l = logf(x);
    }
    // threshold: the nc-th largest thread maximum (-inf when fewer than nc threads saw an admissible element)
    float mine = cm, thr = -INFINITY;
    for (int c = 0; c < nc; ++c) {
        float v = mine;
        int owner = t;
        block_argmax(v, owner, s_v, s_i);
        thr = v;
        if (owner == t) mine = -INFINITY;
    }
    __syncthreads();
    const float lse = s_l;
    const float add = beam_scores ? beam_scores[list] : 0.f;
    // pass 2: collect everything at or above the threshold
    for (int base = t; base < v4; base += TK_THREADS * TU) {
        float4 x4s[TU];
#pragma unroll
        for (int u = 0; u < TU; ++u) {
            const int i4 = base + u * TK_THREADS;
            x4s[u] = (i4 < v4) ? row4[i4] : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        }
#pragma unroll
        for (int u = 0; u < TU; ++u) {
            const int i4 = base + u * TK_THREADS;
            if (i4 < v4) {
                const float xs[4] = {x4s[u].x, x4s[u].y, x4s[u].z, x4s[u].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int idx = 4 * i4 + e;
                    if (xs[e] >= thr && idx != ban_token) {
                        const int pos = atomicAdd(&s_cnt, 1);
                        if (pos < TK_CAP) {
                            c_v[pos] = xs[e];
                            c_i[pos] = idx;
                        }
                    }
                }
            }
        }
    }
    __syncthreads();
    const int cnt = s_cnt;
    if (cnt <= TK_CAP) {
        for (int c = 0; c < nc; ++c) {
            float v = -INFINITY;
            int i = 0x7fffffff;
            for (int e = t; e < cnt; e += TK_THREADS)
                if (cand_better(c_v[e], c_i[e], v, i)) {
                    v = c_v[e];
                    i = c_i[e];
                }
            block_argmax(v, i, s_v, s_i);
            if (t == 0) {
                cand_score[static_cast<int64_t>(list) * nc + c] = ((v - gm) - lse) + add;  // log_softmax, then + beam score
                cand_tok[static_cast<int64_t>(list) * nc + c] = i;
            }
            for (int e = t; e < cnt; e += TK_THREADS)
                if (c_i[e] == i) {
                    c_v[e] = -INFINITY;
                    c_i[e] = 0x7fffffff;
                }
            __syncthreads();
        }
    } else {
        // more than TK_CAP elements tie at or above the threshold: exact selection by nc ordered scans of the row
        float pv = INFINITY;
        int pi = -1;
        for (int c = 0; c < nc; ++c) {
            float v = -INFINITY;
            int i = 0x7fffffff;
            for (int idx = t; idx < V; idx += TK_THREADS) {
                const float x = row[idx];
                const bool after = x < pv || (x == pv && idx > pi);
                if (after && idx != ban_token && cand_better(x, idx, v, i)) {
                    v = x;
                    i = idx;
                }
            }
            block_argmax(v, i, s_v, s_i);
            if (t == 0) {
                cand_score[static_cast<int64_t>(list) * nc + c] = ((v - gm) - lse) + add;
                cand_tok[static_cast<int64_t>(list) * nc + c] = i;
            }
            pv = v;
            pi = i;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Beam bookkeeping: BeamSearchScorer.process of transformers v4.15 for one frame per thread.
// ---------------------------------------------------------------------------------------------------------------------
__device__ void hyp_add(const BeamState& st, int b, const int32_t* toks, int len, float sum_logprobs) {
    const int K = st.beams, Tm = st.t_max;
    const double score = static_cast<double>(sum_logprobs) / pow(static_cast<double>(len), static_cast<double>(st.length_penalty));
    int n = st.hyp_n[b];
    double* hs = st.hyp_score + static_cast<int64_t>(b) * (K + 1);
    int32_t* hl = st.hyp_len + static_cast<int64_t>(b) * (K + 1);
    int32_t* ht = st.hyp_tok + static_cast<int64_t>(b) * (K + 1) * Tm;
    if (n < K || score > st.worst[b]) {
        hs[n] = score;
        hl[n] = len;
        for (int i = 0; i < len; ++i) ht[n * Tm + i] = toks[i];
        ++n;
        if (n > K) {
            // drop the lowest (score, position) entry; keep list order (BeamHypotheses.add)
            int lo = 0;
            for (int i = 1; i < n; ++i)
                if (hs[i] < hs[lo]) lo = i;
            for (int i = lo; i + 1 < n; ++i) {
                hs[i] = hs[i + 1];
                hl[i] = hl[i + 1];
                for (int k = 0; k < hl[i]; ++k) ht[i * Tm + k] = ht[(i + 1) * Tm + k];
            }
            --n;
            double wmin = hs[0];
            for (int i = 1; i < n; ++i) wmin = fmin(wmin, hs[i]);
            st.worst[b] = wmin;
        } else {
            st.worst[b] = fmin(score, st.worst[b]);
        }
        st.hyp_n[b] = n;
    }
}

__global__ void med_beam_step_kernel(BeamState st, const float* __restrict__ cand_score, const int32_t* __restrict__ cand_tok,
                                     int lists_per_frame, int nc, int V, int cur_len, int parity) {
    ptx::griddep_launch();
    ptx::griddep_wait();
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= st.frames) return;
    const int K = st.beams, Tm = st.t_max;
    const int64_t R = static_cast<int64_t>(st.frames) * K;
    const int32_t* seq_in = st.seq + parity * R * Tm;
    int32_t* seq_out = st.seq + (parity ^ 1) * R * Tm;
    const int32_t* anc_in = st.anc + parity * R * Tm;
    int32_t* anc_out = st.anc + (parity ^ 1) * R * Tm;
    if (st.done[b]) {
        for (int j = 0; j < K; ++j) {
            const int64_t r = static_cast<int64_t>(b) * K + j;
            st.beam_scores[r] = 0.f;
            st.cur_tok[r] = st.pad;
            for (int i = 0; i < cur_len; ++i) {
                seq_out[r * Tm + i] = seq_in[r * Tm + i];
                anc_out[r * Tm + i] = (i < cur_len - 1) ? anc_in[r * Tm + i] : static_cast<int32_t>(r);
            }
            seq_out[r * Tm + cur_len] = st.pad;
        }
        return;
    }
    // merge the frame's candidate lists: the 2K best of (score desc, flat index beam*V + token asc)
    constexpr int MAXC = 4 * TK_MAX_NC;
    float cs[MAXC];
    int cflat[MAXC];
    const int total = lists_per_frame * nc;
    for (int l = 0; l < lists_per_frame; ++l)
        for (int c = 0; c < nc; ++c) {
            const int64_t list = (lists_per_frame == 1) ? b : static_cast<int64_t>(b) * K + l;
            cs[l * nc + c] = cand_score[list * nc + c];
            cflat[l * nc + c] = l * V + cand_tok[list * nc + c];
        }
    float top_s[TK_MAX_NC];
    int top_f[TK_MAX_NC];
    for (int c = 0; c < nc; ++c) {
        int best = -1;
        for (int i = 0; i < total; ++i) {
            if (cflat[i] < 0) continue;
            if (best < 0 || cs[i] > cs[best] || (cs[i] == cs[best] && cflat[i] < cflat[best])) best = i;
        }
        top_s[c] = cs[best];
        top_f[c] = cflat[best];
        cflat[best] = -1;
    }
    int slot = 0;
    float new_score[4];
    int new_tok[4], new_parent[4];
    for (int rank = 0; rank < nc && slot < K; ++rank) {
        const int tok = top_f[rank] % V, src = top_f[rank] / V;
        const int64_t prow = static_cast<int64_t>(b) * K + src;
        if (tok == st.eos) {
            if (rank >= K) continue;
            hyp_add(st, b, seq_in + prow * Tm, cur_len, top_s[rank]);
        } else {
            new_score[slot] = top_s[rank];
            new_tok[slot] = tok;
            new_parent[slot] = static_cast<int>(prow);
            ++slot;
        }
    }
    // fewer than K continuations can only happen when nc < 2K; the caller guarantees nc == 2K
    for (int j = 0; j < K; ++j) {
        const int64_t r = static_cast<int64_t>(b) * K + j;
        const int jj = j < slot ? j : slot - 1;
        const int64_t p = new_parent[jj];
        st.beam_scores[r] = j < slot ? new_score[jj] : -1e9f;
        st.cur_tok[r] = new_tok[jj];
        for (int i = 0; i < cur_len; ++i) {
            seq_out[r * Tm + i] = seq_in[p * Tm + i];
            anc_out[r * Tm + i] = (i < cur_len - 1) ? anc_in[p * Tm + i] : static_cast<int32_t>(p);
        }
        seq_out[r * Tm + cur_len] = new_tok[jj];
    }
    // BeamHypotheses.is_done(best_sum_logprobs = max of the step's candidates, cur_len), early_stopping False
    if (st.hyp_n[b] >= K) {
        const double cur = static_cast<double>(top_s[0]) / pow(static_cast<double>(cur_len), static_cast<double>(st.length_penalty));
        if (st.worst[b] >= cur) {
            st.done[b] = 1;
            atomicAdd(st.n_done, 1);
        }
    }
}

// BeamSearchScorer.finalize: open beams of unfinished frames become hypotheses; the best one is written out, followed by
// eos when it fits (len < max_length), padded with pad.
__global__ void med_beam_finalize_kernel(BeamState st, int cur_len, int parity, int max_length, int32_t* __restrict__ out_tokens,
                                         int32_t* __restrict__ out_len, float* __restrict__ out_score) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= st.frames) return;
    const int K = st.beams, Tm = st.t_max;
    const int64_t R = static_cast<int64_t>(st.frames) * K;
    const int32_t* seq_in = st.seq + parity * R * Tm;
    if (!st.done[b])
        for (int j = 0; j < K; ++j) {
            const int64_t r = static_cast<int64_t>(b) * K + j;
            hyp_add(st, b, seq_in + r * Tm, cur_len, st.beam_scores[r]);
        }
    const int n = st.hyp_n[b];
    const double* hs = st.hyp_score + static_cast<int64_t>(b) * (K + 1);
    const int32_t* hl = st.hyp_len + static_cast<int64_t>(b) * (K + 1);
    const int32_t* ht = st.hyp_tok + static_cast<int64_t>(b) * (K + 1) * Tm;
    int best = 0;
    for (int i = 1; i < n; ++i)
        if (hs[i] >= hs[best]) best = i;  // sorted(key=score).pop(): among equal scores the later entry
    int len = hl[best];
    int32_t* o = out_tokens + static_cast<int64_t>(b) * max_length;
    for (int i = 0; i < len; ++i) o[i] = ht[best * Tm + i];
    if (len < max_length) o[len++] = st.eos;
    for (int i = len; i < max_length; ++i) o[i] = st.pad;
    out_len[b] = len;
    out_score[b] = static_cast<float>(hs[best]);
}

__global__ void med_beam_init_kernel(BeamState st, const int32_t* __restrict__ prompt, int prompt_len) {
    const int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    const int K = st.beams, Tm = st.t_max;
    const int64_t R = static_cast<int64_t>(st.frames) * K;
    if (r >= R) return;
    const int b = static_cast<int>(r / K);
    for (int p = 0; p < 2; ++p)
        for (int i = 0; i < Tm; ++i) {
            st.seq[(p * R + r) * Tm + i] = i < prompt_len ? prompt[i] : st.pad;
            st.anc[(p * R + r) * Tm + i] = b * K;
        }
    if (r == 0) *st.n_done = 0;
    st.beam_scores[r] = (r % K == 0) ? 0.f : -1e9f;
    st.cur_tok[r] = prompt[prompt_len - 1];
    if (r % K == 0) {
        st.hyp_n[b] = 0;
        st.worst[b] = 1e9;
        st.done[b] = 0;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Nucleus sampling step (BLIP_Decoder.generate(sample=True), blip.py:139-148 -> transformers v4.15 `sample()`): on the raw
// logits of the step, RepetitionPenaltyLogitsProcessor (tokens of the sequence so far: x < 0 ? x * p : x / p), then
// MinLengthLogitsProcessor, then the warpers — TopKLogitsWarper (config default top_k = 50: everything below the k-th largest
// value goes, ties at it stay) and TopPLogitsWarper (descending order, a token goes once the probability mass BEFORE it
// exceeds top_p) — softmax over what is left and one draw.  The draw is an inverse-CDF lookup in that descending order with a
// uniform number supplied by the caller (torch.multinomial's own stream cannot be reproduced; the distribution is the same).
// One CTA per frame; finished frames emit pad.  With one beam the ping-pong sequence buffers of the beam search hold the same
// tokens, so both are written.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int SP_CAP = 1024;  // tokens that can survive top-k (ties included)

__device__ __forceinline__ uint32_t sp_ordered(float v) {  // monotonic float -> uint
    const uint32_t u = __float_as_uint(v);
    return u ^ ((u >> 31) ? 0xffffffffu : 0x80000000u);
}

__global__ void __launch_bounds__(TK_THREADS)
    med_sample_step_kernel(BeamState st, const float* __restrict__ logits, int64_t ld, int V, int cur_len, int ban_token, int top_k,
                           float top_p, float rep_pen, const float* __restrict__ uniforms) {
    extern __shared__ uint32_t sp_seen[];  // one bit per token: is it in the sequence so far
    __shared__ float c_v[SP_CAP];
    __shared__ int c_i[SP_CAP];
    __shared__ int s_cnt;
    const int b = blockIdx.x, t = threadIdx.x;
    const int Tm = st.t_max;
    int32_t* seq0 = st.seq + static_cast<int64_t>(b) * Tm;
    int32_t* seq1 = st.seq + (static_cast<int64_t>(st.frames) + b) * Tm;
    if (st.done[b]) {  // block-uniform
        if (t == 0) {
            seq0[cur_len] = st.pad;
            seq1[cur_len] = st.pad;
            st.cur_tok[b] = st.pad;
        }
        return;
    }
    const int words = (V + 31) >> 5;
    for (int i = t; i < words; i += TK_THREADS) sp_seen[i] = 0u;
    if (t == 0) s_cnt = 0;
    __syncthreads();
    for (int i = t; i < cur_len; i += TK_THREADS) {
        const int tok = seq0[i];
        if (tok >= 0 && tok < V) atomicOr(&sp_seen[tok >> 5], 1u << (tok & 31));
    }
    __syncthreads();
    const float* row = logits + static_cast<int64_t>(b) * ld;
    auto processed = [&](int idx) {
        float x = row[idx];
        if ((sp_seen[idx >> 5] >> (idx & 31)) & 1u) x = x < 0.f ? x * rep_pen : x / rep_pen;
        return idx == ban_token ? -INFINITY : x;
    };
    // pass 1: every thread's largest processed logit; the top_k-th largest of those 256 maxima bounds the row's top_k-th
    // largest value from below (binary search over the ordered bit pattern, one block-wide count per bit)
    float cm = -INFINITY;
    for (int idx = t; idx < V; idx += TK_THREADS) cm = fmaxf(cm, processed(idx));
    const uint32_t key = sp_ordered(cm), key_ninf = sp_ordered(-INFINITY);
    const int finite_threads = __syncthreads_count(key > key_ninf);
    uint32_t thr = 0u;
    if (finite_threads >= top_k) {
        for (int bit = 31; bit >= 0; --bit) {
            const uint32_t cand = thr | (1u << bit);
            if (__syncthreads_count(key >= cand) >= top_k) thr = cand;
        }
    }
    // pass 2: collect everything at or above it
    for (int idx = t; idx < V; idx += TK_THREADS) {
        const float x = processed(idx);
        if (x != -INFINITY && sp_ordered(x) >= thr) {
            const int pos = atomicAdd(&s_cnt, 1);
            if (pos < SP_CAP) {
                c_v[pos] = x;
                c_i[pos] = idx;
            }
        }
    }
    __syncthreads();
    int n = s_cnt;
    if (n > SP_CAP) {
        // more than SP_CAP values at or above the bound (a flat row): find the exact top_k-th largest value with one count of the
        // row per bit, then keep what is at or above it — in token order, at most SP_CAP of them
        uint32_t exact = 0u;
        for (int bit = 31; bit >= 0; --bit) {
            const uint32_t cand = exact | (1u << bit);
            int mine = 0;
            for (int idx = t; idx < V; idx += TK_THREADS) {
                const float x = processed(idx);
                mine += (x != -INFINITY && sp_ordered(x) >= cand) ? 1 : 0;
            }
            __shared__ int s_tot;
            if (t == 0) s_tot = 0;
            __syncthreads();
            atomicAdd(&s_tot, mine);
            __syncthreads();
            if (s_tot >= top_k) exact = cand;
            __syncthreads();
        }
        if (t == 0) {
            int m = 0;
            for (int idx = 0; idx < V && m < SP_CAP; ++idx) {
                const float x = processed(idx);
                if (x != -INFINITY && sp_ordered(x) >= exact) {
                    c_v[m] = x;
                    c_i[m] = idx;
                    ++m;
                }
            }
            s_cnt = m;
        }
        __syncthreads();
        n = s_cnt;
    }
    // descending bitonic sort of the collected (value, token) pairs, value descending, token ascending among equals
    int P = 1;
    while (P < n) P <<= 1;
    for (int i = n + t; i < P; i += TK_THREADS) {
        c_v[i] = -INFINITY;
        c_i[i] = 0x7fffffff;
    }
    __syncthreads();
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = t; i < P; i += TK_THREADS) {
                const int j = i ^ stride;
                if (j > i) {
                    const bool first_wins = cand_better(c_v[i], c_i[i], c_v[j], c_i[j]);
                    const bool want_first = (i & size) == 0;   // this half sorted best-first
                    if (first_wins != want_first) {
                        const float tv = c_v[i];
                        const int ti = c_i[i];
                        c_v[i] = c_v[j];
                        c_i[i] = c_i[j];
                        c_v[j] = tv;
                        c_i[j] = ti;
                    }
                }
            }
            __syncthreads();
        }
    }
    if (t == 0) {
        // top-k with ties, then top-p, then the draw
        int kept = n;
        if (n > top_k) {
            const float vk = c_v[top_k - 1];
            kept = top_k;
            while (kept < n && c_v[kept] == vk) ++kept;
        }
        const float m = c_v[0];
        float total = 0.f;
        for (int j = 0; j < kept; ++j) total += expf(c_v[j] - m);
        int nucleus = 0;
        float cum = 0.f, mass = 0.f;
        for (int j = 0; j < kept; ++j) {
            if (j > 0 && cum / total > top_p) break;   // the mass before token j exceeds top_p
            const float e = expf(c_v[j] - m);
            cum += e;
            mass = cum;
            ++nucleus;
        }
        const float target = uniforms[b] * mass;
        int pick = nucleus - 1;
        float run = 0.f;
        for (int j = 0; j < nucleus; ++j) {
            run += expf(c_v[j] - m);
            if (run > target) {
                pick = j;
                break;
            }
        }
        const int tok = c_i[pick];
        seq0[cur_len] = tok;
        seq1[cur_len] = tok;
        st.cur_tok[b] = tok;
        st.beam_scores[b] += (c_v[pick] - m) - logf(mass);   // log-probability of the draw under the sampled distribution
        if (tok == st.eos) {
            st.done[b] = 1;
            atomicAdd(st.n_done, 1);
        }
    }
}

// sequences as transformers' sample() returns them: the tokens up to and including eos (or max_length), pad behind
__global__ void med_sample_finalize_kernel(BeamState st, int cur_len, int max_length, int32_t* __restrict__ out_tokens,
                                           int32_t* __restrict__ out_len, float* __restrict__ out_score) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= st.frames) return;
    const int32_t* seq = st.seq + static_cast<int64_t>(b) * st.t_max;
    int32_t* o = out_tokens + static_cast<int64_t>(b) * max_length;
    int len = cur_len;
    for (int i = 0; i < cur_len; ++i) {
        o[i] = seq[i];
        if (seq[i] == st.eos && i + 1 < len) len = i + 1;
    }
    for (int i = len; i < max_length; ++i) o[i] = st.pad;
    out_len[b] = len;
    out_score[b] = st.beam_scores[b];
}

// out[p, j] = hidden[p * T, :] . W[j, :] + bias[j]   (itm_head on the first token, blip_itm.py:56)
__global__ void med_cls_head_kernel(const float* __restrict__ hidden, const float* __restrict__ W, const float* __restrict__ bias,
                                    float* __restrict__ out, int n_seq, int T, int D, int n_out) {
    const int item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (item >= n_seq * n_out) return;
    const int p = item / n_out, j = item - p * n_out;
    const float* x = hidden + static_cast<int64_t>(p) * T * D;
    const float* wr = W + static_cast<int64_t>(j) * D;
    float d = 0.f;
    for (int i = lane; i < D; i += 32) d += x[i] * wr[i];
    d = warp_sum(d);
    if (lane == 0) out[item] = d + (bias ? bias[j] : 0.f);
}

inline int grid_for(int64_t n, int block) {
    int64_t g = (n + block - 1) / block;
    if (g > 148 * 32) g = 148 * 32;
    return static_cast<int>(g < 1 ? 1 : g);
}

}  // namespace

int med_embed_run(const int32_t* ids, const float* word, const float* pos, float* resid, int64_t rows, int T, int pos0, int ids_mod,
                  int D, int vocab, int max_pos, cudaStream_t s) {
    if (rows <= 0) return 0;
    VIDIL_CUDA_OK(launch_pdl(med_embed_kernel, dim3(grid_for(rows * (D / 4), 256)), dim3(256), 0, s, ids, word, pos, resid, rows, T, pos0,
                             ids_mod, D, vocab, max_pos));
    count_launches(1);
    return 0;
}

int med_self_attn_decode_run(const void* qkv, void* cache, const int32_t* anc, void* out, DType dt, int rows, int H, int pos, int Tmax,
                             float scale, cudaStream_t s) {
    if (rows <= 0) return 0;
    if (pos + 1 > SA_MAX_KEYS || pos >= Tmax) {
        set_error("med decode self-attention: position %d, at most %d keys / cache length %d", pos, SA_MAX_KEYS, Tmax);
        return 1;
    }
    const int grid = (rows * H + SA_WARPS - 1) / SA_WARPS;
    if (dt == DT_BF16)
        VIDIL_CUDA_OK(launch_pdl(med_self_attn_decode_kernel<__nv_bfloat16>, dim3(grid), dim3(SA_WARPS * 32), 0, s,
                                 reinterpret_cast<const __nv_bfloat16*>(qkv), reinterpret_cast<__nv_bfloat16*>(cache), anc,
                                 reinterpret_cast<__nv_bfloat16*>(out), rows, H, pos, Tmax, scale));
    else
        VIDIL_CUDA_OK(launch_pdl(med_self_attn_decode_kernel<__half>, dim3(grid), dim3(SA_WARPS * 32), 0, s,
                                 reinterpret_cast<const __half*>(qkv), reinterpret_cast<__half*>(cache), anc, reinterpret_cast<__half*>(out),
                                 rows, H, pos, Tmax, scale));
    count_launches(1);
    return 0;
}

int med_cross_kv_map_prepare(CrossKvMap& m, const void* ckv, int depth, int F, int Nv, int H) {
    typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    m.valid = false;
    m.F = F; m.Nv = Nv; m.H = H;
    const int64_t rows = static_cast<int64_t>(depth) * F * Nv;
    if (fn == nullptr || rows > 0x7fffffffLL) {
        set_error("decode cross-attention: cuTensorMapEncodeTiled unavailable or %lld K/V rows exceed the tensor-map coordinate range",
                  static_cast<long long>(rows));
        return 1;
    }
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(2 * H * 64), static_cast<cuuint64_t>(rows)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(2 * H * 64) * 2};
    const cuuint32_t box[2] = {64, static_cast<cuuint32_t>(Nv <= 256 ? Nv : cross_decode_mma_chunk_rows())};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(&m.map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(ckv), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (cross K/V) failed with CUresult %d", static_cast<int>(r));
        return 1;
    }
    m.valid = true;
    return 0;
}

int med_cross_attn_decode_run(const CrossKvMap* map, int layer, const void* q, void* out, DType dt, int F, int nq, int Nv, int H,
                              float scale, cudaStream_t s) {
    if (F <= 0) return 0;
    if (nq < 1 || nq > 4 || F > 65535) {
        set_error("med decode cross-attention: %d beams (1..4), %d frames (<= 65535)", nq, F);
        return 1;
    }
    if (map == nullptr || !map->valid || map->F != F || map->Nv != Nv || map->H != H) {
        set_error("med decode cross-attention: no tensor map prepared for %d frames x %d image tokens x %d heads", F, Nv, H);
        return 1;
    }
    return cross_decode_mma_run(map->map, layer * F * Nv, q, out, dt, F, nq, Nv, H, scale, s);
}

int med_cache_fill_run(const void* qkv, void* cache, DType dt, int64_t n_rows, int T_seq, int D, int Tmax, int beams, cudaStream_t s) {
    if (n_rows <= 0) return 0;
    const int grid = grid_for(n_rows * (2 * D / 8), 256);
    // 16-bit payload is only moved: one instantiation serves both operand types
    med_cache_fill_kernel<__half><<<grid, 256, 0, s>>>(reinterpret_cast<const __half*>(qkv), reinterpret_cast<__half*>(cache), n_rows,
                                                      T_seq, D, Tmax, beams);
    (void)dt;
    VIDIL_CUDA_OK(cudaGetLastError());
    count_launches(1);
    return 0;
}

int med_logits_topk_run(const float* logits, int64_t ld, int row_mul, const float* beam_scores, int n_lists, int V, int nc, int ban_token,
                        float* cand_score, int32_t* cand_tok, cudaStream_t s) {
    if (n_lists <= 0) return 0;
    if (nc < 1 || nc > TK_MAX_NC || V % 4 != 0 || ld % 4 != 0) {
        set_error("med top-k: nc=%d (1..%d), V=%d and ld=%lld must be multiples of 4", nc, TK_MAX_NC, V, (long long)ld);
        return 1;
    }
    VIDIL_CUDA_OK(launch_pdl(med_logits_topk_kernel, dim3(n_lists), dim3(TK_THREADS), 0, s, logits, ld, row_mul, beam_scores, V, nc, ban_token,
                             cand_score, cand_tok));
    count_launches(1);
    return 0;
}

int med_beam_init_run(const BeamState& st, const int32_t* prompt_dev, int prompt_len, cudaStream_t s) {
    const int64_t R = static_cast<int64_t>(st.frames) * st.beams;
    med_beam_init_kernel<<<static_cast<int>((R + 127) / 128), 128, 0, s>>>(st, prompt_dev, prompt_len);
    VIDIL_CUDA_OK(cudaGetLastError());
    count_launches(1);
    return 0;
}

int med_beam_step_run(const BeamState& st, const float* cand_score, const int32_t* cand_tok, int lists_per_frame, int nc, int V,
                      int cur_len, int parity, cudaStream_t s) {
    if (st.beams > 4 || nc != 2 * st.beams) {
        set_error("beam step: num_beams=%d (at most 4) with %d candidates per list (must be 2*num_beams)", st.beams, nc);
        return 1;
    }
    VIDIL_CUDA_OK(launch_pdl(med_beam_step_kernel, dim3((st.frames + 63) / 64), dim3(64), 0, s, st, cand_score, cand_tok, lists_per_frame, nc, V,
                             cur_len, parity));
    count_launches(1);
    return 0;
}

int med_sample_step_run(const BeamState& st, const float* logits, int64_t ld, int V, int cur_len, int ban_token, int top_k, float top_p,
                        float repetition_penalty, const float* uniforms_dev, cudaStream_t s) {
    if (st.beams != 1 || top_k < 1 || top_k > SP_CAP || !(top_p > 0.f) || !(repetition_penalty > 0.f)) {
        set_error("sample step: one beam, top_k in [1, %d], top_p > 0 and repetition_penalty > 0 expected (beams %d, top_k %d, top_p %g, "
                  "penalty %g)", SP_CAP, st.beams, top_k, top_p, repetition_penalty);
        return 1;
    }
    const size_t smem = static_cast<size_t>((V + 31) / 32) * sizeof(uint32_t);
    if (smem > 40 * 1024) {
        set_error("sample step: vocabulary of %d tokens exceeds the shared-memory bit set", V);
        return 1;
    }
    med_sample_step_kernel<<<st.frames, TK_THREADS, smem, s>>>(st, logits, ld, V, cur_len, ban_token, top_k, top_p, repetition_penalty, uniforms_dev);
    VIDIL_CUDA_OK(cudaGetLastError());
    count_launches(1);
    return 0;
}

int med_sample_finalize_run(const BeamState& st, int cur_len, int max_length, int32_t* out_tokens, int32_t* out_len, float* out_score,
                            cudaStream_t s) {
    med_sample_finalize_kernel<<<(st.frames + 63) / 64, 64, 0, s>>>(st, cur_len, max_length, out_tokens, out_len, out_score);
    VIDIL_CUDA_OK(cudaGetLastError());
    count_launches(1);
    return 0;
}

int med_beam_finalize_run(const BeamState& st, int cur_len, int parity, int max_length, int32_t* out_tokens, int32_t* out_len,
                          float* out_score, cudaStream_t s) {
    med_beam_finalize_kernel<<<(st.frames + 63) / 64, 64, 0, s>>>(st, cur_len, parity, max_length, out_tokens, out_len, out_score);
    VIDIL_CUDA_OK(cudaGetLastError());
    count_launches(1);
    return 0;
}

int med_cls_head_run(const float* hidden, const float* W, const float* bias, float* out, int n_seq, int T, int D, int n_out,
                     cudaStream_t s) {
    if (n_seq <= 0 || n_out <= 0) return 0;
    const int items = n_seq * n_out;
    med_cls_head_kernel<<<(items + 7) / 8, 256, 0, s>>>(hidden, W, bias, out, n_seq, T, D, n_out);
    VIDIL_CUDA_OK(cudaGetLastError());
    count_launches(1);
    return 0;
}

}  // namespace vidil
