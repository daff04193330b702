// Internal C++ interface between the C-ABI layer (api.cu) and the kernel translation units.
// Nothing here is exported; the public surface is include/vidil_b200.h.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace vidil {

// Arithmetic type of the tensor-core operands (accumulation is always fp32).
enum DType : int { DT_BF16 = 0, DT_FP16 = 1 };

// What the GEMM epilogue does with the fp32 accumulator tile (acc = A * W^T):
enum EpiMode : int {
    EPI_STORE = 0,      // out[T]   = acc + bias
    EPI_GELU = 1,       // out[T]   = gelu_erf(acc + bias)            (BLIP ViT Mlp, vit.py:35-41)
    EPI_QUICKGELU = 2,  // out[T]   = x * sigmoid(1.702 x), x = acc+bias   (CLIP quick_gelu)
    EPI_RESID = 3,      // out[f32] += acc + bias                     (residual stream, vit.py:108-109)
    EPI_PATCH = 4,      // out[f32][frame*(P+1)+1+p] = acc + bias + pos[1+p]   (vit.py:182-187)
    EPI_STORE_F32 = 5,  // out[f32] = acc + bias
    EPI_TOP4 = 6,       // out[f32][row][col / 32][4] = the four largest acc of every 32-column group, the column's position in the
                        // group packed into the 5 low mantissa bits (similarity + top-k: the [M,N] scores are never written)
    EPI_COUNT = 7
};

void set_error(const char* fmt, ...);
const char* get_error();
// Every launcher adds the number of kernels it enqueued (reported by vidil_kernel_launch_count()).
void count_launches(int n);
int64_t launch_count();

// Whether kernels are launched with programmatic stream serialization (PDL); VIDIL_PDL=0 in the environment switches it off.
bool pdl_enabled();
void pdl_scope(int delta);  // +1 / -1 around a region whose launches should use PDL (unless VIDIL_PDL forces it either way)

// <<<grid, block, smem, stream>>> with the PDL attribute: the kernel may begin before its predecessor in the stream has finished
// and must call ptx::griddep_wait() before it touches global memory.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

#define VIDIL_CUDA_OK(expr)                                                                  \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            ::vidil::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return 1;                                                                        \
        }                                                                                    \
    } while (0)

// ---------------------------------------------------------------------------------------
// GEMM  D[M,N] = A[M,K] * W[N,K]^T  (both operands K-major, 16-bit), tcgen05 + TMA.
// ---------------------------------------------------------------------------------------
struct GemmProblem {
    DType dt = DT_BF16;
    int epi = EPI_STORE;
    int cta_group = 2;  // 1: one CTA per 128x256 tile; 2: CTA pair per 256x256 tile (cta_group::2)
    int M = 0, N = 0, K = 0;
    const void* A = nullptr;  // [M, lda] 16-bit
    int64_t lda = 0;
    const void* W = nullptr;  // [N, ldw] 16-bit
    int64_t ldw = 0;
    const float* bias = nullptr;  // [N] or null
    void* out = nullptr;          // [M, ldo] T or fp32 depending on epi
    int64_t ldo = 0;
    const float* pos = nullptr;  // EPI_PATCH: [P+1, N] fp32
    int patches_per_frame = 0;   // EPI_PATCH
    // filled by gemm_prepare():
    CUtensorMap map_a;
    CUtensorMap map_w;
    CUtensorMap map_out;  // epilogue TMA store / reduce-add target (all modes but EPI_PATCH)
    bool prepared = false;
};

int gemm_prepare(GemmProblem& p);                          // validates + encodes the TMA maps
int gemm_run(const GemmProblem& p, cudaStream_t stream);   // launches; asynchronous
int gemm_num_sms();

// ---------------------------------------------------------------------------------------
// LayerNorm over the last dim of an fp32 matrix (vit.py:94,99,192; eps 1e-6 BLIP / 1e-5 CLIP).
//   in : fp32 rows at in + r * in_row_stride      out: T or fp32 rows at out + r * D
// ---------------------------------------------------------------------------------------
int layernorm_run(const float* in, int64_t in_row_stride, const float* gamma, const float* beta, void* out,
                  bool out_f32, DType dt, int rows, int D, float eps, cudaStream_t stream);

// Post-LayerNorm of the BERT-style text stack (med.py:245,316): resid[r,:] = LN(resid[r,:]) in place (fp32) and, when
// out16 != null, the same row as the 16-bit operand of the next GEMM.
int layernorm_post_run(float* resid, const float* gamma, const float* beta, void* out16, DType dt, int rows, int D, float eps,
                       cudaStream_t stream);
// LayerNorm of a 16-bit matrix into a 16-bit matrix (BertPredictionHeadTransform, med.py:513-517)
int layernorm16_run(const void* in16, const float* gamma, const float* beta, void* out16, DType dt, int rows, int D, float eps,
                    cudaStream_t stream);

// ---------------------------------------------------------------------------------------
// Fused softmax attention over packed qkv [B, N, 3, H, 64] -> out [B, N, H*64] (vit.py:72-83).
// ---------------------------------------------------------------------------------------
int attention_run(const void* qkv, void* out, DType dt, int B, int N, int H, float scale, cudaStream_t stream);

// tcgen05 version for sequences of at most 208 tokens (attention_tc.cu).  The TMA descriptors depend only on the
// buffer addresses and the shape, so a forward plan encodes them once.
struct AttentionMaps {
    CUtensorMap q, kv, out;
    CUtensorMap kv16;      // attention_tc257.cu: the 16-row K / V box holding token 256
    bool is_257 = false;   // prepared for attention_tc257_run
    const void* qkv = nullptr;
    void* out_ptr = nullptr;
    int B = 0, N = 0, H = 0;
    DType dt = DT_BF16;
    bool causal = false;  // query i attends to keys 0..i only (CLIP text tower)
    // long-sequence kernel (attention_tcl.cu): key-block size, number of key blocks, query tiles done on tcgen05
    bool is_long = false;
    int KB = 0, nkb = 0, n_qt = 0, n_tail = 0;
};
// Developer hook: a device buffer of 5*16*8 int64 that CTA 0 fills with clock64 stamps of its pipeline events.
void attention_set_trace(long long* dev_buf);
bool attention_tc_supported(int N);
int attention_tc_prepare(AttentionMaps& m, const void* qkv, void* out, DType dt, int B, int N, int H);
int attention_tc_run(const AttentionMaps& m, float scale, cudaStream_t stream);
// CLIP ViT-L/14's 257 tokens: single-pass kernel with the 257th key / query handled by SIMT (attention_tc257.cu)
bool attention_tc257_supported(int N);
int attention_tc257_prepare(AttentionMaps& m, const void* qkv, void* out, DType dt, int B, int N, int H);
int attention_tc257_run(const AttentionMaps& m, float scale, cudaStream_t stream);
// tcgen05 version with a key-block loop for 128 < N <= 768 (attention_tcl.cu); non-causal
bool attention_tcl_supported(int N);
int attention_tcl_prepare(AttentionMaps& m, const void* qkv, void* out, DType dt, int B, int N, int H);
int attention_tcl_run(const AttentionMaps& m, float scale, cudaStream_t stream);

// ---------------------------------------------------------------------------------------
// Elementwise / layout kernels.
// ---------------------------------------------------------------------------------------
// fp32 -> T cast of a dense [rows, cols] matrix into [rows, ld_out] (pad columns zeroed).
int cast_run(const float* in, void* out, DType dt, int64_t rows, int64_t cols, int64_t ld_out,
             cudaStream_t stream);
// NCHW fp32 frames -> patch matrix [B*P, Kpad] T, column = c*ps*ps + i*ps + j (Conv2d weight order).
int im2col_run(const float* frames, void* patches, DType dt, int B, int C, int img, int ps, int Kpad,
               cudaStream_t stream);
// resid[b*(P+1) + 0, :] = cls + pos[0]  (vit.py:184-187)
int cls_pos_run(const float* cls, const float* pos, float* resid, int B, int tokens, int D, cudaStream_t stream);
// out[r,:] = in[r,:] / ||in[r,:]||_2
int l2norm_run(const float* in, float* out, int rows, int D, cudaStream_t stream);
// resid[r, :] = tok[ids[r], :] + pos[r % L, :]  and  out[b, :] = in[b*L + idx[b], :]  (CLIP text tower)
int embed_tokens_run(const int32_t* ids, const float* tok, const float* pos, float* resid, int64_t rows, int L, int D,
                     int vocab, cudaStream_t stream);
int gather_rows_run(const float* in, const int32_t* idx, float* out, int B, int L, int D, cudaStream_t stream);
// T -> fp32 (operator-level tests read 16-bit results back through this)
int uncast_run(const void* in, float* out, DType dt, int64_t n, cudaStream_t stream);

// ---------------------------------------------------------------------------------------
// Frame pre-processing (run_video_CapFilt.py:128-137): uint8 HWC -> PIL-identical bicubic resize -> /255 -> normalise.
// ---------------------------------------------------------------------------------------
size_t preprocess_workspace_bytes(int B, int H, int W, int S);
int preprocess_run(const uint8_t* frames, int B, int H, int W, int S, const float* mean, const float* stdv, float* out,
                   void* workspace, size_t workspace_bytes, cudaStream_t stream);
// transformers' CLIPImageProcessor (run_visual_tokenization.py:138-140): shortest edge -> S bicubic, centre crop S x S,
// rescale by the double 1/255, normalise.
size_t clip_preprocess_workspace_bytes(int B, int H, int W, int S);
int clip_preprocess_run(const uint8_t* frames, int B, int H, int W, int S, const float* mean, const float* stdv, float* out,
                        void* workspace, size_t workspace_bytes, cudaStream_t stream);

// ---------------------------------------------------------------------------------------
// Similarity top-k (run_visual_tokenization.py:276,306).
// ---------------------------------------------------------------------------------------
// Exact top-k from the EPI_TOP4 output of the similarity GEMM: top4 [F][G][4] fp32 (G = ceil(T / 32) groups, index-tagged
// approximate scores), img [F, D] / bank [T, D] fp32 originals.  Every entry within 2 * eps_scale * |img row| * bank_max_norm of
// the k-th largest approximate score is re-scored in fp32; a group whose fourth-best qualifies is re-scored completely (a
// fifth member may hide behind it).
int topk_select_run(const float* top4, int ld_top4, int G, const float* img, const float* bank, const float* bank_max_norm_dev, int F, int T, int D,
                    int k, float* out_scores, int32_t* out_idx, cudaStream_t stream);
// max over rows of ||row||_2 of an fp32 [rows, D] matrix -> out[0] (device)
int max_row_norm_run(const float* m, int rows, int D, float* out, cudaStream_t stream);

// ---------------------------------------------------------------------------------------
// med.py text stack (caption decoder / ITM encoder): SIMT kernels around the GEMMs (med.cu).
// ---------------------------------------------------------------------------------------
// Attention with separate query and key/value sequences on the mma.sync pipeline (attention.cu): group g = query rows
// [g*nq, (g+1)*nq) at q + row*q_stride + head*64 attends to the nk keys/values of frame f = frame_of_group[g] (g when null) at
// k|v + (f*nk + j)*kv_stride + head*64.  key_mask int32 [groups][nk] (0 = masked, -10000 added) or null; causal: key j > row i
// masked.  out rows at out + row*out_stride + head*64.
int attention_x_run(const void* q, int64_t q_stride, const void* k, const void* v, int64_t kv_stride, const int32_t* frame_of_group,
                    const int32_t* key_mask, void* out, int64_t out_stride, DType dt, int groups, int nq, int nk, int H, bool causal,
                    float scale, cudaStream_t stream);

enum MedAttnMode : int {
    MED_ATTN_FULL = 0,    // every query sees every key of its sequence that the padding mask allows (encoder, blip_itm.py:49)
    MED_ATTN_CAUSAL = 1,  // query i sees keys 0..i of its sequence (decoder over whole sequences / the prompt)
    MED_ATTN_DECODE = 2   // one new token per row at position `pos`; earlier keys come from the cache via the ancestry table
};
// Device-side state of a beam search over `frames` frames x `beams` beams (row r = frame * beams + beam).
struct BeamState {
    int frames = 0, beams = 0, t_max = 0, eos = 0, pad = 0;
    float length_penalty = 1.f;
    int32_t* seq = nullptr;        // [2][R][t_max] token sequences (ping-pong across steps)
    int32_t* anc = nullptr;        // [2][R][t_max] row whose cache slot holds position t of this beam's history
    float* beam_scores = nullptr;  // [R] sum of log-probs
    int32_t* cur_tok = nullptr;    // [R] token fed to the next step
    int32_t* hyp_n = nullptr;      // [frames] finished hypotheses kept (<= beams)
    double* hyp_score = nullptr;   // [frames][beams+1]
    int32_t* hyp_len = nullptr;    // [frames][beams+1]
    int32_t* hyp_tok = nullptr;    // [frames][beams+1][t_max]
    double* worst = nullptr;       // [frames]
    int32_t* done = nullptr;       // [frames]
    int32_t* n_done = nullptr;     // [1] number of frames whose search has finished (the host polls it to stop early)
};
// resid[r,:] = word[ids[ids_mod ? r % ids_mod : r],:] + pos[pos0 + r % T,:]
int med_embed_run(const int32_t* ids, const float* word, const float* pos, float* resid, int64_t rows, int T, int pos0, int ids_mod,
                  int D, int vocab, int max_pos, cudaStream_t s);
// decode step: qkv [rows, 3D] (one new token per row at position pos) -> out [rows, D]; cache [R][Tmax][2D] (K then V) receives the
// new K/V at slot (row, pos); earlier keys are read from slot (anc[row][t], t)
int med_self_attn_decode_run(const void* qkv, void* cache, const int32_t* anc, void* out, DType dt, int rows, int H, int pos, int Tmax,
                             float scale, cudaStream_t s);
// decode step: the nq beams of frame f (rows f*nq.. of q [F*nq, D]) attend to kv [F, Nv, 2D] -> out [F*nq, D].  With a prepared
// CrossKvMap (one tensor map over the cross K/V of all layers, [depth*F*Nv, 2D]) the K/V tiles come in by TMA.
struct CrossKvMap {
    CUtensorMap map;  // 128B swizzle, 64 columns wide: one [Nv, 64] box up to 256 tokens per frame, [128, 64] boxes beyond
    bool valid = false;
    int F = 0, Nv = 0, H = 0;
};
int cross_decode_mma_chunk_rows();
int cross_decode_mma_run(const CUtensorMap& kv_map_sw128, int row0, const void* q, void* out, DType dt, int F, int nq, int Nv, int H,
                         float scale, cudaStream_t stream);
int med_cross_kv_map_prepare(CrossKvMap& m, const void* ckv, int depth, int F, int Nv, int H);
int med_cross_attn_decode_run(const CrossKvMap* map, int layer, const void* q, void* out, DType dt, int F, int nq, int Nv, int H,
                              float scale, cudaStream_t s);
// K/V of whole sequences, qkv [n_seq*T_seq, 3D] -> cache slots (seq*beams, t)
int med_cache_fill_run(const void* qkv, void* cache, DType dt, int64_t n_rows, int T_seq, int D, int Tmax, int beams, cudaStream_t s);
// list l scans logits row l*row_mul: log_softmax, ban_token excluded (-1: none), + beam_scores[l] -> nc best (score, token)
int med_logits_topk_run(const float* logits, int64_t ld, int row_mul, const float* beam_scores, int n_lists, int V, int nc, int ban_token,
                        float* cand_score, int32_t* cand_tok, cudaStream_t s);
int med_beam_init_run(const BeamState& st, const int32_t* prompt_dev, int prompt_len, cudaStream_t s);
int med_beam_step_run(const BeamState& st, const float* cand_score, const int32_t* cand_tok, int lists_per_frame, int nc, int V,
                      int cur_len, int parity, cudaStream_t s);
// nucleus sampling (one beam): frame b draws its next token from logits row b after repetition penalty, MinLength (ban_token),
// top-k and top-p, by inverse CDF with uniforms_dev[b]; finished frames emit pad
int med_sample_step_run(const BeamState& st, const float* logits, int64_t ld, int V, int cur_len, int ban_token, int top_k, float top_p,
                        float repetition_penalty, const float* uniforms_dev, cudaStream_t s);
int med_sample_finalize_run(const BeamState& st, int cur_len, int max_length, int32_t* out_tokens, int32_t* out_len, float* out_score,
                            cudaStream_t s);
int med_beam_finalize_run(const BeamState& st, int cur_len, int parity, int max_length, int32_t* out_tokens, int32_t* out_len,
                          float* out_score, cudaStream_t s);
int med_cls_head_run(const float* hidden, const float* W, const float* bias, float* out, int n_seq, int T, int D, int n_out,
                     cudaStream_t s);

}  // namespace vidil
