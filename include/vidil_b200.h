/*
 * vidil_b200 — C ABI of the B200-native frame-encoding path of VidIL.
 *
 * One shared library (vidil_b200/_C/libvidil_b200.so), plain pointers and sizes only.  It replaces the
 * arithmetic behind these reference call sites (paths relative to the VidIL repository):
 *
 *   models/vit.py:180-194        VisionTransformer.forward          -> vidil_vit_forward
 *   models/blip.py:128, blip_itm.py:43   visual_encoder(image)      -> vidil_vit_forward
 *   run_visual_tokenization.py:141-142   CLIPModel(**inputs).image_embeds -> vidil_clip_forward
 *   run_visual_tokenization.py:276,306   image_embeds @ text_embeds.t(); np.argsort(...)[::-1][:k]
 *                                                                    -> vidil_sim_topk
 *   run_visual_tokenization.py:84-96     CLIPModel(**inputs).text_embeds  -> vidil_clip_text_forward
 *   run_video_CapFilt.py:128-137         process_frame (PIL resize + ToTensor + Normalize) -> vidil_preprocess_frames
 *   run_visual_tokenization.py:138-140   processor(images=frames) = transformers' CLIPImageProcessor (shortest edge 224
 *                                        bicubic, centre crop, rescale, normalise)       -> vidil_clip_preprocess_frames
 *   models/blip.py:127-167 BLIP_Decoder.generate(sample=False) from the image tokens on: models/med.py:811-955
 *                                        BertLMHeadModel + transformers' beam search      -> vidil_med_generate
 *                                        ... + transformers' nucleus sampling             -> vidil_med_sample
 *   models/blip_itm.py:49-57 text_encoder(mode multimodal) + itm_head; models/med.py:871-893 teacher-forced logits
 *                                                                                         -> vidil_med_forward
 *
 * Conventions
 *   - Every function returns 0 on success, non-zero on failure; vidil_last_error() then returns a
 *     NUL-terminated description (thread-local, valid until the next call on that thread).
 *   - All *device* pointers are borrowed for the duration of the call only.  The library owns nothing
 *     but the packed 16-bit weights inside an encoder handle.  No allocation happens in a forward call.
 *   - Work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream) and the
 *     call returns without synchronising, except the *_host variants, which synchronise before returning.
 *   - One process per GPU; a handle is bound to the device that was current at creation and is not
 *     thread-safe.
 *   - There is no CPU fallback: on a machine without an sm_100 GPU every compute entry point fails.
 */
#ifndef VIDIL_B200_H_
#define VIDIL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VIDIL_B200_ABI_VERSION 2

/* Tensor-core operand type.  Accumulation, LayerNorm/softmax statistics and the residual stream are fp32. */
enum { VIDIL_DTYPE_BF16 = 0, VIDIL_DTYPE_FP16 = 1 };
/* MLP activation: exact erf GELU (nn.GELU, vit.py:26) or CLIP's quick_gelu x*sigmoid(1.702x). */
enum { VIDIL_ACT_GELU_ERF = 0, VIDIL_ACT_QUICK_GELU = 1 };

/* Architecture of one vision tower.  BLIP ViT (models/blip.py:298-326): patch 16, patch_bias 1, pre_ln 0,
 * ln_eps 1e-6, act GELU_ERF, proj_dim 0.  CLIP ViT-L/14: patch 14, patch_bias 0, pre_ln 1, ln_eps 1e-5,
 * act QUICK_GELU, proj_dim 768. */
typedef struct vidil_encoder_cfg {
    int32_t img_size;    /* square input side, e.g. 224 or 384 */
    int32_t patch_size;  /* 16 (BLIP) or 14 (CLIP); must be even and divide img_size */
    int32_t embed_dim;   /* D: 768 or 1024 (multiple of 128, head_dim must be 64) */
    int32_t depth;       /* number of transformer blocks */
    int32_t num_heads;   /* D / 64 */
    int32_t mlp_dim;     /* 4*D */
    float   ln_eps;
    int32_t act;         /* VIDIL_ACT_* */
    int32_t patch_bias;  /* 1: patch conv has a bias */
    int32_t pre_ln;      /* 1: LayerNorm after the position add (CLIP pre_layrnorm) */
    int32_t proj_dim;    /* >0: CLS -> post LN -> bias-free projection to proj_dim -> L2 normalise (CLIP) */
    int32_t dtype;       /* VIDIL_DTYPE_* */
    int32_t cta_group;   /* 1 or 2: CTAs cooperating on one tcgen05 tile; 0 = library default (2) */
} vidil_encoder_cfg;

typedef struct vidil_encoder vidil_encoder;

int32_t vidil_abi_version(void);
const char* vidil_last_error(void);

/* Number of CUDA kernels this library has launched in this process (all streams).  bench.py reads it
 * before/after the timed region to report gpu_launches. */
int64_t vidil_kernel_launch_count(void);

/* ---- encoder handle -------------------------------------------------------------------------- */
int32_t vidil_encoder_create(const vidil_encoder_cfg* cfg, vidil_encoder** out);
void    vidil_encoder_destroy(vidil_encoder* enc);

/* Load one parameter from a device fp32 tensor (contiguous, `numel` elements).  `name` uses the
 * reference ViT state_dict keys (models/vit.py):
 *   cls_token [D]  pos_embed [(P+1)*D]  patch_embed.proj.weight [D*3*ps*ps]  patch_embed.proj.bias [D]
 *   blocks.<i>.norm1.weight|bias  blocks.<i>.attn.qkv.weight [3D*D]|bias [3D]
 *   blocks.<i>.attn.proj.weight|bias  blocks.<i>.norm2.weight|bias
 *   blocks.<i>.mlp.fc1.weight|bias  blocks.<i>.mlp.fc2.weight|bias  norm.weight|bias
 * plus, for CLIP-style towers: pre_norm.weight|bias (pre_ln) and head.proj.weight [proj_dim*D] (proj_dim).
 * Matrices are converted to the 16-bit operand type into handle-owned buffers; vectors stay fp32. */
int32_t vidil_encoder_load(vidil_encoder* enc, const char* name, const float* dev_ptr, int64_t numel, void* stream);
/* Fails (and names the first missing parameter) unless every parameter has been loaded. */
int32_t vidil_encoder_check_loaded(const vidil_encoder* enc);

size_t  vidil_encoder_workspace_bytes(const vidil_encoder* enc, int32_t batch);
int32_t vidil_encoder_tokens(const vidil_encoder* enc); /* P + 1 */

/* VisionTransformer.forward: frames fp32 NCHW [B,3,S,S] (device) -> tokens fp32 [B, P+1, D] (device),
 * all tokens after the final LayerNorm (vit.py:192-194). */
int32_t vidil_vit_forward(vidil_encoder* enc, const float* frames, int32_t batch, float* out_tokens,
                          void* workspace, size_t workspace_bytes, void* stream);
/* Same forward, tokens written in the handle's 16-bit operand type (bf16 / fp16) instead of fp32: for consumers that take
 * 16-bit tokens anyway (this library's own text stack casts them on entry) and for host transfers — half the D2H bytes. */
int32_t vidil_vit_forward16(vidil_encoder* enc, const float* frames, int32_t batch, void* out_tokens16, void* workspace,
                            size_t workspace_bytes, void* stream);

/* CLIP vision tower + visual_projection + L2 normalisation: frames -> image_embeds fp32 [B, proj_dim].
 * If out_hidden != NULL it also receives last_hidden_state (before post_layernorm) fp32 [B, P+1, D]. */
int32_t vidil_clip_forward(vidil_encoder* enc, const float* frames, int32_t batch, float* out_embeds,
                           float* out_hidden, void* workspace, size_t workspace_bytes, void* stream);

/* Host-buffer variants (the reference-facing call with CPU tensors): copy frames host->device, run the
 * forward above, copy the result device->host, synchronise.  Host buffers should be pinned for full
 * PCIe bandwidth.  Device staging and workspace are the caller's (dev_scratch, at least
 * vidil_encoder_host_scratch_bytes(enc, batch) bytes). */
size_t  vidil_encoder_host_scratch_bytes(const vidil_encoder* enc, int32_t batch);
int32_t vidil_vit_forward_host(vidil_encoder* enc, const float* frames_host, int32_t batch, float* out_tokens_host,
                               void* dev_scratch, size_t dev_scratch_bytes, void* stream);
int32_t vidil_clip_forward_host(vidil_encoder* enc, const float* frames_host, int32_t batch, float* out_embeds_host,
                                void* dev_scratch, size_t dev_scratch_bytes, void* stream);

/* Pipelined variant for a stream of batches: submit() enqueues H2D (copy stream) -> forward (`stream`) -> D2H
 * (second copy stream) for one batch into one of two slots and returns at once; wait(slot) blocks until that
 * slot's result is in out_host.  Submitting batch k+1 to the other slot before waiting for batch k overlaps its
 * H2D with batch k's forward and batch k's D2H with batch k+1's forward.  The handle decides which forward runs
 * (vit: tokens, clip: image_embeds).  dev_scratch: vidil_encoder_host_pipeline_scratch_bytes(enc, batch) bytes,
 * the same pointer for both slots.  The slot offsets inside dev_scratch are fixed by the batch of the call that first
 * binds a (pointer, size) pair: later calls with a smaller batch (the ragged last batch of a stream) reuse them; a
 * call with a different scratch or a larger batch first waits for everything in flight, then re-binds.  The caller
 * must not touch a slot's host buffers between submit and wait, nor free dev_scratch before both slots are waited. */
size_t  vidil_encoder_host_pipeline_scratch_bytes(const vidil_encoder* enc, int32_t batch);
int32_t vidil_encoder_host_submit(vidil_encoder* enc, const float* frames_host, int32_t batch, float* out_host,
                                  int32_t slot, void* dev_scratch, size_t dev_scratch_bytes, void* stream);
/* submit() with the ViT tokens delivered to the host in the handle's 16-bit operand type (out_host16: batch * tokens *
 * embed_dim 16-bit values). */
int32_t vidil_encoder_host_submit16(vidil_encoder* enc, const float* frames_host, int32_t batch, void* out_host16, int32_t slot,
                                    void* dev_scratch, size_t dev_scratch_bytes, void* stream);
int32_t vidil_encoder_host_wait(vidil_encoder* enc, int32_t slot);

/* ---- CLIP text tower (the phrase bank of run_visual_tokenization.py:84-96) ----------------------- */
/* transformers' CLIPTextModel + text_projection + L2 normalisation.  Parameter names for vidil_text_encoder_load:
 *   token_embedding [V*D]  position_embedding [P*D]  blocks.<i>.* as for the image towers (q,k,v fused into attn.qkv)
 *   norm.weight|bias (final_layer_norm)  head.proj.weight [proj_dim*D] (text_projection).
 * input_ids int32 [B, L] and eos_pos int32 [B] (the pooled position of each sequence: argmax of the ids for the
 * legacy eos_token_id == 2 configs, else the first eos_token_id) are device pointers; out_embeds fp32 [B, proj_dim].
 * Attention is causal; padding after the EOS token cannot influence the pooled row and needs no mask. */
typedef struct vidil_text_cfg {
    int32_t vocab_size;
    int32_t max_positions; /* <= 208 */
    int32_t embed_dim;     /* 64 * num_heads */
    int32_t depth;
    int32_t num_heads;
    int32_t mlp_dim;
    float   ln_eps;
    int32_t act;           /* VIDIL_ACT_* */
    int32_t proj_dim;
    int32_t dtype;         /* VIDIL_DTYPE_* */
    int32_t cta_group;     /* 0 = default (2) */
} vidil_text_cfg;
typedef struct vidil_text_encoder vidil_text_encoder;
int32_t vidil_text_encoder_create(const vidil_text_cfg* cfg, vidil_text_encoder** out);
void    vidil_text_encoder_destroy(vidil_text_encoder* enc);
int32_t vidil_text_encoder_load(vidil_text_encoder* enc, const char* name, const float* dev_ptr, int64_t numel, void* stream);
int32_t vidil_text_encoder_check_loaded(const vidil_text_encoder* enc);
size_t  vidil_text_encoder_workspace_bytes(const vidil_text_encoder* enc, int32_t batch, int32_t seq_len);
int32_t vidil_clip_text_forward(vidil_text_encoder* enc, const int32_t* input_ids, const int32_t* eos_pos, int32_t batch,
                                int32_t seq_len, float* out_embeds, void* workspace, size_t workspace_bytes, void* stream);

/* ---- med.py text stack: caption decoder and ITM filter (run_video_CapFilt.py:95-123) ------------ */
/* BertModel with cross-attention onto the image tokens (models/med.py), optionally followed by the LM head
 * (BertLMHeadModel, lm_head = 1) or by a Linear on the first token (BLIP_ITM.itm_head, cls_out = 2).
 * Parameter names for vidil_med_load (matrices [out, in] as nn.Linear stores them; q,k,v of the self-attention and
 * k,v of the cross-attention are concatenated along the output dim by the caller):
 *   word_embeddings [V*D]  position_embeddings [P*D]  emb_ln.weight|bias
 *   layer.<i>.self.qkv.weight [3D*D]|bias  layer.<i>.self.out.weight [D*D]|bias  layer.<i>.self.ln.weight|bias
 *   layer.<i>.cross.q.weight [D*D]|bias    layer.<i>.cross.kv.weight [2D*E]|bias  layer.<i>.cross.out.weight|bias
 *   layer.<i>.cross.ln.weight|bias         layer.<i>.ffn.fc1.weight [I*D]|bias    layer.<i>.ffn.fc2.weight [D*I]|bias
 *   layer.<i>.ffn.ln.weight|bias
 *   lm_head: head.dense.weight [D*D]|bias  head.ln.weight|bias  head.decoder.weight [V*D]  head.decoder.bias [V]
 *   cls_out: cls.weight [cls_out*D]  cls.bias [cls_out] */
typedef struct vidil_med_cfg {
    int32_t vocab_size;     /* multiple of 4 (30524 for BLIP) */
    int32_t max_positions;
    int32_t hidden;         /* D = 64 * num_heads; 128/256/512/768/1024 */
    int32_t depth;
    int32_t num_heads;
    int32_t mlp_dim;        /* I */
    int32_t encoder_width;  /* E: width of the image tokens (vision_width, blip.py:96), multiple of 64 */
    float   ln_eps;         /* 1e-12 */
    int32_t lm_head;        /* 1: BertOnlyMLMHead present */
    int32_t cls_out;        /* >0: Linear(D, cls_out) on token 0 */
    int32_t dtype;          /* VIDIL_DTYPE_* */
    int32_t cta_group;      /* 0 = library default */
} vidil_med_cfg;
typedef struct vidil_med vidil_med;
int32_t vidil_med_create(const vidil_med_cfg* cfg, vidil_med** out);
void    vidil_med_destroy(vidil_med* med);
int32_t vidil_med_load(vidil_med* med, const char* name, const float* dev_ptr, int64_t numel, void* stream);
int32_t vidil_med_check_loaded(const vidil_med* med);

/* Whole-sequence forward.  image_embeds fp32 [n_frames, n_img_tokens, E]; input_ids int32 [n_seq, seq_len];
 * frame_of_seq int32 [n_seq] (NULL: sequence i reads frame i, n_seq == n_frames); all device pointers.
 * seqs_per_frame > 0 (frame_of_seq must be NULL, n_seq == n_frames * seqs_per_frame): sequences are ordered frame-major, sequence i
 * reads frame i / seqs_per_frame, and the cross-attention handles all sequences of a frame in one query group (the captions of a
 * video against one of its frames, run_video_CapFilt.py:107-126).
 *   causal = 1: decoder (BertLMHeadModel.forward is_decoder=True, all-ones attention mask, med.py:871-893)
 *   causal = 0: encoder with the padding mask attention_mask int32 [n_seq, seq_len] (blip_itm.py:49-54)
 * Outputs, each optional (NULL): out_hidden fp32 [n_seq, seq_len, D] (last_hidden_state), out_logits fp32
 * [n_seq, seq_len, V] (needs lm_head), out_cls fp32 [n_seq, cls_out] (needs cls_out). */
size_t  vidil_med_forward_workspace_bytes(const vidil_med* med, int32_t n_seq, int32_t seq_len, int32_t n_frames,
                                          int32_t n_img_tokens);
int32_t vidil_med_forward(vidil_med* med, const float* image_embeds, int32_t n_frames, int32_t n_img_tokens,
                          const int32_t* input_ids, const int32_t* attention_mask, const int32_t* frame_of_seq,
                          int32_t seqs_per_frame, int32_t n_seq, int32_t seq_len, int32_t causal, float* out_hidden, float* out_logits,
                          float* out_cls, void* workspace, size_t workspace_bytes, void* stream);

/* Beam-search captioning, BLIP_Decoder.generate(sample=False) after the ViT (blip.py:130-167): the prompt (HOST pointer,
 * prompt_len ids, id 0 already replaced by bos as blip.py:136-137 does) is decoded once per frame, then
 * max_length - prompt_len - 1 single-token steps run on n_frames * num_beams rows with a K/V cache that is never
 * re-ordered (a per-beam ancestry table replaces transformers' _reorder_cache, med.py:951-955).  Cross-attention K/V of the
 * image tokens are projected once per frame.  Search rules: transformers v4.15 beam_search / BeamSearchScorer with
 * early_stopping False, repetition_penalty 1.0.  Outputs (device): out_tokens int32 [n_frames, max_length] = best
 * hypothesis incl. the prompt, followed by eos when it fits, padded with pad; out_lengths int32 [n_frames];
 * out_scores fp32 [n_frames] (sum of log-probs / len^length_penalty).  num_beams <= 4, max_length <= 64.
 * Like transformers' loop the search stops as soon as every frame is done: from min_length on the call synchronises `stream`
 * every other step to read one counter; the outputs are enqueued on `stream` as everywhere else. */
size_t  vidil_med_generate_workspace_bytes(const vidil_med* med, int32_t n_frames, int32_t n_img_tokens, int32_t num_beams,
                                           int32_t max_length, int32_t prompt_len);
int32_t vidil_med_generate(vidil_med* med, const float* image_embeds, int32_t n_frames, int32_t n_img_tokens,
                           const int32_t* prompt_ids_host, int32_t prompt_len, int32_t num_beams, int32_t max_length,
                           int32_t min_length, int32_t eos_token, int32_t pad_token, float length_penalty, int32_t* out_tokens,
                           int32_t* out_lengths, float* out_scores, void* workspace, size_t workspace_bytes, void* stream);

/* Nucleus sampling, BLIP_Decoder.generate(sample=True) after the ViT (blip.py:139-148; run_video_CapFilt.py:104 with
 * generation_mode "sample"): one sequence per frame, same prompt handling, K/V cache and early stop as vidil_med_generate with
 * one beam (workspace: vidil_med_generate_workspace_bytes(..., num_beams = 1, ...)).  Every step applies, in transformers
 * v4.15 order, RepetitionPenaltyLogitsProcessor (blip.py passes 1.1), MinLengthLogitsProcessor, TopKLogitsWarper (the
 * PretrainedConfig default top_k = 50 that BLIP inherits) and TopPLogitsWarper, then draws from the softmax of what is left.
 * The draw is an inverse-CDF lookup in descending-probability order with the caller's uniform numbers — `uniforms` fp32
 * (device) [max_length - prompt_len, n_frames] in [0, 1), step-major — so a run is reproducible from its random stream;
 * torch.multinomial's own stream cannot be reproduced, the distribution is the same.  Outputs as vidil_med_generate, except
 * that out_tokens holds the sequence up to and including eos (no hypothesis selection) and out_scores the sum of the drawn
 * tokens' log-probabilities under the sampled distributions.  top_k in [1, 1024]. */
int32_t vidil_med_sample(vidil_med* med, const float* image_embeds, int32_t n_frames, int32_t n_img_tokens,
                         const int32_t* prompt_ids_host, int32_t prompt_len, int32_t max_length, int32_t min_length, int32_t eos_token,
                         int32_t pad_token, int32_t top_k, float top_p, float repetition_penalty, const float* uniforms,
                         int32_t* out_tokens, int32_t* out_lengths, float* out_scores, void* workspace, size_t workspace_bytes,
                         void* stream);

/* The search alone over given logits (parity of the bookkeeping): step_logits fp32 [n_steps, n_frames*num_beams, V]
 * (device) plays the decoder; step 0 reads the row of beam 0 of every frame. */
size_t  vidil_op_beam_search_workspace_bytes(int32_t n_frames, int32_t num_beams, int32_t max_length);
int32_t vidil_op_beam_search(const float* step_logits, int32_t n_steps, int32_t n_frames, int32_t num_beams, int32_t V,
                             const int32_t* prompt_ids_host, int32_t prompt_len, int32_t max_length, int32_t min_length,
                             int32_t eos_token, int32_t pad_token, float length_penalty, int32_t* out_tokens,
                             int32_t* out_lengths, float* out_scores, void* workspace, size_t workspace_bytes, void* stream);

/* The sampling alone over given logits (parity of the processors / warpers and the draw): step_logits fp32
 * [n_steps, n_frames, V] (device) plays the decoder; workspace: vidil_op_beam_search_workspace_bytes(n_frames, 1, max_length). */
int32_t vidil_op_sample(const float* step_logits, int32_t n_steps, int32_t n_frames, int32_t V, const int32_t* prompt_ids_host,
                        int32_t prompt_len, int32_t max_length, int32_t min_length, int32_t eos_token, int32_t pad_token, int32_t top_k,
                        float top_p, float repetition_penalty, const float* uniforms, int32_t* out_tokens, int32_t* out_lengths,
                        float* out_scores, void* workspace, size_t workspace_bytes, void* stream);

/* ---- frame pre-processing (run_video_CapFilt.py:128-137 process_frame) ------------------------- */
/* frames_u8: device uint8 [B, H, W, 3] (decoded RGB frames, HWC) -> out: device fp32 [B, 3, S, S]:
 * Pillow's antialiased BICUBIC resize to S x S reproduced bit for bit (22-bit fixed-point weights, uint8 intermediate
 * between the horizontal and the vertical pass), then x / 255 and (x - mean) / std in float32 as torchvision's ToTensor
 * and Normalize do.  mean3 / std3 are HOST pointers to 3 floats. */
size_t  vidil_preprocess_workspace_bytes(int32_t batch, int32_t in_h, int32_t in_w, int32_t out_size);
int32_t vidil_preprocess_frames(const uint8_t* frames_u8, int32_t batch, int32_t in_h, int32_t in_w, int32_t out_size,
                                const float* mean3, const float* std3, float* out, void* workspace, size_t workspace_bytes,
                                void* stream);

/* The CLIP side (run_visual_tokenization.py:138-140: `processor(images=frames, return_tensors="pt")`, transformers'
 * CLIPImageProcessor with its PIL backend): resize so that the SHORTER side becomes S (the longer one int(S * long / short)),
 * Pillow BICUBIC as above; keep the centre S x S window (top = (rh - S) / 2, left = (rw - S) / 2, integer division);
 * rescale by the double 1/255 rounded to float32; (x - mean) / std in float32.  Bit-identical to the processor's output
 * (tests/golden/clip_preprocess.npz).  Only the columns / rows of the window are computed. */
size_t  vidil_clip_preprocess_workspace_bytes(int32_t batch, int32_t in_h, int32_t in_w, int32_t out_size);
int32_t vidil_clip_preprocess_frames(const uint8_t* frames_u8, int32_t batch, int32_t in_h, int32_t in_w, int32_t out_size,
                                     const float* mean3, const float* std3, float* out, void* workspace,
                                     size_t workspace_bytes, void* stream);

/* ---- per-kernel-class device timing (bench.py's roofline figures) ---------------------------- */
/* With profiling on, every kernel a forward enqueues is bracketed by CUDA events on the caller's stream.
 * vidil_encoder_read_profile synchronises those events, adds up elapsed time / algorithmic FLOPs / algorithmic
 * bytes / launch counts per class since the last read, writes them to *out and resets the counters. */
enum { VIDIL_KCLASS_GEMM = 0, VIDIL_KCLASS_ATTENTION = 1, VIDIL_KCLASS_LAYERNORM = 2, VIDIL_KCLASS_OTHER = 3,
       VIDIL_KCLASS_COUNT = 4 };
typedef struct vidil_kernel_stats {
    double  ms[VIDIL_KCLASS_COUNT];
    double  flops[VIDIL_KCLASS_COUNT];
    double  bytes[VIDIL_KCLASS_COUNT];
    int64_t launches[VIDIL_KCLASS_COUNT];
} vidil_kernel_stats;
int32_t vidil_encoder_set_profiling(vidil_encoder* enc, int32_t enable);
int32_t vidil_encoder_read_profile(vidil_encoder* enc, vidil_kernel_stats* out);
/* The same for a text-stack handle.  Classes there: GEMM = every projection incl. the cross K/V and the vocabulary GEMM;
 * ATTENTION = the cross-attention onto the image tokens (bytes = each query group's K and V tiles once + its q and out rows);
 * LAYERNORM = the post-LayerNorm kernel; OTHER = text self-attention, vocabulary log-softmax/top-k. */
int32_t vidil_med_set_profiling(vidil_med* med, int32_t enable);
int32_t vidil_med_read_profile(vidil_med* med, vidil_kernel_stats* out);

/* ---- similarity + top-k ----------------------------------------------------------------------- */
/* Replaces run_visual_tokenization.py:276 (`sims = image_embeds @ text_embeds.t()`), :299 (D2H of [F,T]) and :306
 * (`np.argsort(score)[::-1][:k]` per frame).  img fp32 [F,D], bank fp32 [T,D] (device, D a multiple of 64) -> out_scores
 * fp32 [F,k], out_idx int32 [F,k]: for each frame the k phrases with the largest FP32 dot product, best first (k <= 12;
 * exact ties are reported index-descending).  The [F,T] matrix is never written: a tcgen05 GEMM on fp16 copies of the
 * operands keeps the two best scores of every 32-phrase group in its epilogue, and candidates are then re-scored in fp32 from
 * the original embeddings until the fp32 ranking is certain (error bound 2^-10 |img| max|bank row| on the fp16 product), so
 * the indices are those of the fp32 ranking for any data, near-duplicate phrases included.  Any T that fits memory.
 *
 * A phrase bank is constant for a whole run: vidil_sim_bank_create converts it once (fp16 copy for the tensor cores, fp32
 * copy for the re-scoring, largest row norm) and vidil_sim_bank_topk uses the handle.  vidil_sim_topk is the one-shot form
 * (the bank is converted inside the call, in the workspace). */
typedef struct vidil_sim_bank vidil_sim_bank;
int32_t vidil_sim_bank_create(const float* bank, int32_t T, int32_t D, void* stream, vidil_sim_bank** out);
void    vidil_sim_bank_destroy(vidil_sim_bank* bank);
size_t  vidil_sim_bank_topk_workspace_bytes(const vidil_sim_bank* bank, int32_t F);
int32_t vidil_sim_bank_topk(const vidil_sim_bank* bank, const float* img, int32_t F, int32_t k, float* out_scores,
                            int32_t* out_idx, void* workspace, size_t workspace_bytes, void* stream);
size_t  vidil_sim_topk_workspace_bytes(int32_t F, int32_t T, int32_t D);
int32_t vidil_sim_topk(const float* img, const float* bank, int32_t F, int32_t T, int32_t D, int32_t k,
                       float* out_scores, int32_t* out_idx, void* workspace, size_t workspace_bytes, void* stream);

/* ---- single operators (parity tests call these; same kernels the forwards use) --------------- */
/* out = epilogue(A[M,K] @ W[N,K]^T): A, W fp32 on device are cast to `dtype` first (workspace).  epi:
 * 0 store T, 1 gelu_erf T, 2 quick_gelu T, 3 out_f32 += , 4 patch scatter (pos, patches_per_frame), 5 store f32.
 * out_is_f32 tells how `out` is typed for epi 0..2 the result is written as fp32 after a T round trip. */
size_t  vidil_op_linear_workspace_bytes(int32_t M, int32_t N, int32_t K);
int32_t vidil_op_linear(const float* A, const float* W, const float* bias, float* out, int32_t M, int32_t N, int32_t K,
                        int32_t epi, int32_t dtype, int32_t cta_group, const float* pos, int32_t patches_per_frame,
                        void* workspace, size_t workspace_bytes, void* stream);
int32_t vidil_op_layernorm(const float* in, const float* gamma, const float* beta, float* out, int32_t rows, int32_t D,
                           float eps, void* stream);
/* qkv fp32 [B,N,3,H,64] -> out fp32 [B,N,H*64]; operands are rounded to `dtype` as in the forward. */
size_t  vidil_op_attention_workspace_bytes(int32_t B, int32_t N, int32_t H);
int32_t vidil_op_attention(const float* qkv, float* out, int32_t B, int32_t N, int32_t H, float scale, int32_t dtype,
                           void* workspace, size_t workspace_bytes, void* stream);
/* Same with a causal mask (query i attends to keys 0..i): the CLIP text tower's attention
 * (run_visual_tokenization.py:84-96 through transformers' CLIPTextTransformer).  N <= 208. */
int32_t vidil_op_attention_causal(const float* qkv, float* out, int32_t B, int32_t N, int32_t H, float scale, int32_t dtype,
                                  void* workspace, size_t workspace_bytes, void* stream);

/* ---- developer hooks ---------------------------------------------------------------------------- */
/* dev_buf: device buffer of 5*16*8 int64 (or NULL to switch off).  While set, CTA 0 of the tcgen05 attention kernel
 * stamps clock64() at its pipeline synchronisation points for its first 16 items (tools/attn_trace.py prints them). */
void vidil_debug_set_attention_trace(void* dev_buf);

#ifdef __cplusplus
}
#endif
#endif /* VIDIL_B200_H_ */
