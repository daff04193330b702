// C ABI of the med.py text stack (include/vidil_b200.h, "med.py text stack"): the caption decoder that
// BLIP_Decoder.generate drives (models/blip.py:127-167, models/med.py:811-955) and the multimodal encoder + ITM head that
// BLIP_ITM.forward drives (models/blip_itm.py:41-58), both fed by the image tokens of the ViT.
//
// Dense projections run on the tcgen05 GEMM of gemm.cu (fp32 residual stream, 16-bit operands, bias / exact-erf GELU /
// residual-add epilogues); BERT's post-LayerNorm, the short text attentions, the vocabulary scan and the beam bookkeeping
// are the SIMT kernels of layernorm.cu / med.cu.  Nothing here allocates on the forward path: every buffer is carved from
// the caller's workspace.
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <memory>
#include <string>
#include <vector>

#include "../../include/vidil_b200.h"
#include "kernels.h"

using namespace vidil;

namespace {

constexpr size_t ALIGN = 1024;
inline size_t align_up(size_t x) { return (x + ALIGN - 1) / ALIGN * ALIGN; }

struct MBuf {
    void* p = nullptr;
    size_t bytes = 0;
    bool loaded = false;
    MBuf() = default;
    MBuf(const MBuf&) = delete;
    MBuf& operator=(const MBuf&) = delete;
    ~MBuf() {
        if (p) cudaFree(p);
    }
    int alloc(size_t n) {
        VIDIL_CUDA_OK(cudaMalloc(&p, n));
        bytes = n;
        return 0;
    }
    const float* f() const { return reinterpret_cast<const float*>(p); }
};

struct MedLayer {
    MBuf sqkv_w, sqkv_b, so_w, so_b, sln_w, sln_b;
    MBuf cq_w, cq_b, ckv_w, ckv_b, co_w, co_b, cln_w, cln_b;
    MBuf fc1_w, fc1_b, fc2_w, fc2_b, fln_w, fln_b;
};

struct MSlot {
    MBuf* buf = nullptr;
    int64_t numel = 0;
    bool matrix = false;
    int rows = 0, cols = 0;
};

// Carves aligned regions out of a workspace; with base == nullptr it only measures.
struct Carver {
    uint8_t* base;
    size_t off = 0;
    explicit Carver(void* b) : base(reinterpret_cast<uint8_t*>(b)) {}
    template <typename T>
    T* take(size_t count) {
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += align_up(count * sizeof(T));
        return p;
    }
    void* take_bytes(size_t n) { return take<uint8_t>(n); }
};

struct StackBufs {
    float* resid = nullptr;
    void *xn = nullptr, *qkv = nullptr, *attn = nullptr, *qc = nullptr, *hidden = nullptr;
};

struct StackPlan {
    int rows = 0;
    std::vector<GemmProblem> sqkv, so, cq, co, fc1, fc2;
};

}  // namespace

// Event-pair records of the launches enqueued while profiling is on (read back by vidil_med_read_profile).
struct MedProfRec {
    cudaEvent_t a, b;
    int cls;
    double flops, bytes;
};

struct vidil_med {
    vidil_med_cfg cfg;
    DType dt = DT_BF16;
    bool profiling = false;
    std::vector<MedProfRec> prof;
    std::vector<cudaEvent_t> free_events;
    ~vidil_med() {
        for (auto& r : prof) {
            cudaEventDestroy(r.a);
            cudaEventDestroy(r.b);
        }
        for (auto ev : free_events) cudaEventDestroy(ev);
    }
    MBuf word, pos, eln_w, eln_b;
    std::vector<std::unique_ptr<MedLayer>> layers;
    MBuf hd_w, hd_b, hln_w, hln_b, dec_w, dec_b, cls_w, cls_b;
};

namespace {

bool med_slot(vidil_med* m, const char* name, MSlot& s) {
    const vidil_med_cfg& c = m->cfg;
    const int D = c.hidden, I = c.mlp_dim, E = c.encoder_width;
    auto vec = [&](MBuf& b, int64_t n) { s = MSlot{&b, n, false, 0, 0}; return true; };
    auto mat = [&](MBuf& b, int r, int k) { s = MSlot{&b, static_cast<int64_t>(r) * k, true, r, k}; return true; };
    if (!strcmp(name, "word_embeddings")) return vec(m->word, static_cast<int64_t>(c.vocab_size) * D);
    if (!strcmp(name, "position_embeddings")) return vec(m->pos, static_cast<int64_t>(c.max_positions) * D);
    if (!strcmp(name, "emb_ln.weight")) return vec(m->eln_w, D);
    if (!strcmp(name, "emb_ln.bias")) return vec(m->eln_b, D);
    if (c.lm_head) {
        if (!strcmp(name, "head.dense.weight")) return mat(m->hd_w, D, D);
        if (!strcmp(name, "head.dense.bias")) return vec(m->hd_b, D);
        if (!strcmp(name, "head.ln.weight")) return vec(m->hln_w, D);
        if (!strcmp(name, "head.ln.bias")) return vec(m->hln_b, D);
        if (!strcmp(name, "head.decoder.weight")) return mat(m->dec_w, c.vocab_size, D);
        if (!strcmp(name, "head.decoder.bias")) return vec(m->dec_b, c.vocab_size);
    }
    if (c.cls_out > 0) {
        if (!strcmp(name, "cls.weight")) return vec(m->cls_w, static_cast<int64_t>(c.cls_out) * D);
        if (!strcmp(name, "cls.bias")) return vec(m->cls_b, c.cls_out);
    }
    int idx = -1, consumed = 0;
    if (sscanf(name, "layer.%d.%n", &idx, &consumed) == 1 && consumed > 0 && idx >= 0 && idx < c.depth) {
        const char* r = name + consumed;
        MedLayer& ly = *m->layers[idx];
        if (!strcmp(r, "self.qkv.weight")) return mat(ly.sqkv_w, 3 * D, D);
        if (!strcmp(r, "self.qkv.bias")) return vec(ly.sqkv_b, 3 * D);
        if (!strcmp(r, "self.out.weight")) return mat(ly.so_w, D, D);
        if (!strcmp(r, "self.out.bias")) return vec(ly.so_b, D);
        if (!strcmp(r, "self.ln.weight")) return vec(ly.sln_w, D);
        if (!strcmp(r, "self.ln.bias")) return vec(ly.sln_b, D);
        if (!strcmp(r, "cross.q.weight")) return mat(ly.cq_w, D, D);
        if (!strcmp(r, "cross.q.bias")) return vec(ly.cq_b, D);
        if (!strcmp(r, "cross.kv.weight")) return mat(ly.ckv_w, 2 * D, E);
        if (!strcmp(r, "cross.kv.bias")) return vec(ly.ckv_b, 2 * D);
        if (!strcmp(r, "cross.out.weight")) return mat(ly.co_w, D, D);
        if (!strcmp(r, "cross.out.bias")) return vec(ly.co_b, D);
        if (!strcmp(r, "cross.ln.weight")) return vec(ly.cln_w, D);
        if (!strcmp(r, "cross.ln.bias")) return vec(ly.cln_b, D);
        if (!strcmp(r, "ffn.fc1.weight")) return mat(ly.fc1_w, I, D);
        if (!strcmp(r, "ffn.fc1.bias")) return vec(ly.fc1_b, I);
        if (!strcmp(r, "ffn.fc2.weight")) return mat(ly.fc2_w, D, I);
        if (!strcmp(r, "ffn.fc2.bias")) return vec(ly.fc2_b, D);
        if (!strcmp(r, "ffn.ln.weight")) return vec(ly.fln_w, D);
        if (!strcmp(r, "ffn.ln.bias")) return vec(ly.fln_b, D);
    }
    return false;
}

void med_param_names(const vidil_med* m, std::vector<std::string>& out) {
    out = {"word_embeddings", "position_embeddings", "emb_ln.weight", "emb_ln.bias"};
    static const char* per_layer[] = {"self.qkv.weight", "self.qkv.bias", "self.out.weight", "self.out.bias", "self.ln.weight",
                                      "self.ln.bias", "cross.q.weight", "cross.q.bias", "cross.kv.weight", "cross.kv.bias",
                                      "cross.out.weight", "cross.out.bias", "cross.ln.weight", "cross.ln.bias", "ffn.fc1.weight",
                                      "ffn.fc1.bias", "ffn.fc2.weight", "ffn.fc2.bias", "ffn.ln.weight", "ffn.ln.bias"};
    for (int i = 0; i < m->cfg.depth; ++i)
        for (const char* n : per_layer) out.push_back("layer." + std::to_string(i) + "." + n);
    if (m->cfg.lm_head)
        for (const char* n : {"head.dense.weight", "head.dense.bias", "head.ln.weight", "head.ln.bias", "head.decoder.weight",
                              "head.decoder.bias"})
            out.push_back(n);
    if (m->cfg.cls_out > 0) {
        out.push_back("cls.weight");
        out.push_back("cls.bias");
    }
}

GemmProblem make_gemm(const vidil_med* m, int epi, int M, int N, int K, const void* A, int64_t lda, const MBuf& W, const MBuf* bias,
                      void* out, int64_t ldo) {
    GemmProblem g;
    g.dt = m->dt;
    g.cta_group = m->cfg.cta_group;
    g.epi = epi;
    g.M = M; g.N = N; g.K = K;
    g.A = A; g.lda = lda;
    g.W = W.p; g.ldw = K;
    g.bias = bias ? bias->f() : nullptr;
    g.out = out; g.ldo = ldo;
    return g;
}

int med_event(vidil_med* m, cudaEvent_t* ev) {
    if (!m->free_events.empty()) {
        *ev = m->free_events.back();
        m->free_events.pop_back();
        return 0;
    }
    VIDIL_CUDA_OK(cudaEventCreate(ev));
    return 0;
}

// Runs one launcher; with profiling on, brackets it with an event pair on the same stream (class: VIDIL_KCLASS_*; here
// ATTENTION = cross-attention onto the image tokens, OTHER = self-attention, vocabulary scan, beam bookkeeping, embedding).
template <typename Fn>
int med_timed(const vidil_med* cm, cudaStream_t s, int cls, double flops, double bytes, Fn&& fn) {
    vidil_med* m = const_cast<vidil_med*>(cm);
    if (!m->profiling) return fn();
    MedProfRec r{nullptr, nullptr, cls, flops, bytes};
    if (med_event(m, &r.a) || med_event(m, &r.b)) return 1;
    VIDIL_CUDA_OK(cudaEventRecord(r.a, s));
    const int rc = fn();
    VIDIL_CUDA_OK(cudaEventRecord(r.b, s));
    m->prof.push_back(r);
    return rc;
}

int gemm_timed(const vidil_med* m, cudaStream_t s, const GemmProblem& g) {
    const double out = (g.epi == EPI_RESID) ? 8.0 : (g.epi == EPI_STORE_F32) ? 4.0 : 2.0;
    return med_timed(m, s, VIDIL_KCLASS_GEMM, 2.0 * g.M * g.N * g.K, 2.0 * g.M * g.K + 2.0 * g.N * g.K + out * g.M * g.N,
                     [&] { return gemm_run(g, s); });
}

void carve_stack(Carver& cv, const vidil_med* m, size_t rows, StackBufs& b) {
    const size_t D = m->cfg.hidden;
    b.resid = cv.take<float>(rows * D);
    b.xn = cv.take_bytes(rows * D * 2);
    b.qkv = cv.take_bytes(rows * 3 * D * 2);
    b.attn = cv.take_bytes(rows * D * 2);
    b.qc = cv.take_bytes(rows * D * 2);
    b.hidden = cv.take_bytes(rows * m->cfg.mlp_dim * 2);
}

int plan_stack(const vidil_med* m, StackPlan& pl, const StackBufs& b, int rows) {
    const vidil_med_cfg& c = m->cfg;
    const int D = c.hidden, I = c.mlp_dim;
    pl.rows = rows;
    for (int i = 0; i < c.depth; ++i) {
        const MedLayer& ly = *m->layers[i];
        pl.sqkv.push_back(make_gemm(m, EPI_STORE, rows, 3 * D, D, b.xn, D, ly.sqkv_w, &ly.sqkv_b, b.qkv, 3 * D));
        pl.so.push_back(make_gemm(m, EPI_RESID, rows, D, D, b.attn, D, ly.so_w, &ly.so_b, b.resid, D));
        pl.cq.push_back(make_gemm(m, EPI_STORE, rows, D, D, b.xn, D, ly.cq_w, &ly.cq_b, b.qc, D));
        pl.co.push_back(make_gemm(m, EPI_RESID, rows, D, D, b.attn, D, ly.co_w, &ly.co_b, b.resid, D));
        pl.fc1.push_back(make_gemm(m, EPI_GELU, rows, I, D, b.xn, D, ly.fc1_w, &ly.fc1_b, b.hidden, I));
        pl.fc2.push_back(make_gemm(m, EPI_RESID, rows, D, I, b.hidden, I, ly.fc2_w, &ly.fc2_b, b.resid, D));
        if (gemm_prepare(pl.sqkv.back()) || gemm_prepare(pl.so.back()) || gemm_prepare(pl.cq.back()) || gemm_prepare(pl.co.back()) ||
            gemm_prepare(pl.fc1.back()) || gemm_prepare(pl.fc2.back()))
            return 1;
    }
    return 0;
}

// Launches inside the scope use programmatic dependent launch (kernels.h): measured on the decode chain of ~2400 short kernels
// (6-140 us each): 67.6 -> 64.6 ms per 1024-frame call; the whole-sequence passes (ITM, prompt) and the ViT lose 1-4 % with it.
struct PdlScope {
    PdlScope() { pdl_scope(+1); }
    ~PdlScope() { pdl_scope(-1); }
    PdlScope(const PdlScope&) = delete;
    PdlScope& operator=(const PdlScope&) = delete;
};

struct AttnArgs {
    int mode = MED_ATTN_FULL;
    int T_seq = 1;                   // tokens per sequence among the rows (1 in decode mode)
    int pos = 0;                     // decode: position of the new token
    int groups = 0, nq = 0;          // cross-attention: query groups and rows per group
    const int32_t* frame_of_group = nullptr;
    const void* cross_kv = nullptr;  // [depth][F][Nv][2D]
    const CrossKvMap* kv_map = nullptr;  // decode: TMA view of cross_kv
    size_t cross_layer_elems = 0;
    int Nv = 0;
    void* cache = nullptr;           // [depth][R][Tmax][2D] or null
    size_t cache_layer_elems = 0;
    const int32_t* anc = nullptr;
    const int32_t* mask = nullptr;
    int beams = 1, Tmax = 0;
};

// embeddings LayerNorm, then depth x BertLayer (med.py:333-384).  resid holds word+position embeddings on entry and
// last_hidden_state (fp32) on exit; xn its 16-bit copy.
int run_stack(const vidil_med* m, const StackPlan& pl, const StackBufs& b, const AttnArgs& a, cudaStream_t s) {
    const vidil_med_cfg& c = m->cfg;
    const int D = c.hidden, H = c.num_heads, rows = pl.rows;
    const float eps = c.ln_eps;
    if (med_timed(m, s, VIDIL_KCLASS_LAYERNORM, 0.0, 10.0 * rows * D, [&] { return layernorm_post_run(b.resid, m->eln_w.f(), m->eln_b.f(), b.xn, m->dt, rows, D, eps, s); })) return 1;
    for (int i = 0; i < c.depth; ++i) {
        const MedLayer& ly = *m->layers[i];
        if (gemm_timed(m, s, pl.sqkv[i])) return 1;
        uint8_t* cache_l = a.cache ? reinterpret_cast<uint8_t*>(a.cache) + static_cast<size_t>(i) * a.cache_layer_elems * 2 : nullptr;
        const uint8_t* qkv8 = reinterpret_cast<const uint8_t*>(b.qkv);
        if (a.mode == MED_ATTN_DECODE) {
            if (med_timed(m, s, VIDIL_KCLASS_OTHER, 0.0, 0.0, [&] {
                    return med_self_attn_decode_run(b.qkv, cache_l, a.anc, b.attn, m->dt, rows, H, a.pos, a.Tmax, 0.125f, s);
                }))
                return 1;
        } else {
            if (med_timed(m, s, VIDIL_KCLASS_OTHER, 4.0 * rows * a.T_seq * D, 0.0, [&] {
                    return attention_x_run(b.qkv, 3 * D, qkv8 + static_cast<size_t>(D) * 2, qkv8 + static_cast<size_t>(2 * D) * 2, 3 * D,
                                           nullptr, a.mask, b.attn, D, m->dt, rows / a.T_seq, a.T_seq, a.T_seq, H,
                                           a.mode == MED_ATTN_CAUSAL, 0.125f, s);
                }))
                return 1;
            if (cache_l && med_cache_fill_run(b.qkv, cache_l, m->dt, rows, a.T_seq, D, a.Tmax, a.beams, s)) return 1;
        }
        if (gemm_timed(m, s, pl.so[i])) return 1;
        if (med_timed(m, s, VIDIL_KCLASS_LAYERNORM, 0.0, 10.0 * rows * D, [&] { return layernorm_post_run(b.resid, ly.sln_w.f(), ly.sln_b.f(), b.xn, m->dt, rows, D, eps, s); })) return 1;
        if (gemm_timed(m, s, pl.cq[i])) return 1;
        const void* kv_l = reinterpret_cast<const uint8_t*>(a.cross_kv) + static_cast<size_t>(i) * a.cross_layer_elems * 2;
        // algorithmic bytes of the cross-attention: each group's K and V tiles once (2 * Nv * 64 * 2 B per head) + q and out rows
        const double x_bytes = static_cast<double>(a.groups) * H * (2.0 * a.Nv * 128 + 2.0 * a.nq * 128);
        const double x_flops = 4.0 * a.groups * a.nq * static_cast<double>(a.Nv) * D;
        if (med_timed(m, s, VIDIL_KCLASS_ATTENTION, x_flops, x_bytes, [&] {
                if (a.mode == MED_ATTN_DECODE && a.frame_of_group == nullptr)
                    return med_cross_attn_decode_run(a.kv_map, i, b.qc, b.attn, m->dt, a.groups, a.nq, a.Nv, H, 0.125f, s);
                return attention_x_run(b.qc, D, kv_l, reinterpret_cast<const uint8_t*>(kv_l) + static_cast<size_t>(D) * 2, 2 * D,
                                       a.frame_of_group, nullptr, b.attn, D, m->dt, a.groups, a.nq, a.Nv, H, false, 0.125f, s);
            }))
            return 1;
        if (gemm_timed(m, s, pl.co[i])) return 1;
        if (med_timed(m, s, VIDIL_KCLASS_LAYERNORM, 0.0, 10.0 * rows * D, [&] { return layernorm_post_run(b.resid, ly.cln_w.f(), ly.cln_b.f(), b.xn, m->dt, rows, D, eps, s); })) return 1;
        if (gemm_timed(m, s, pl.fc1[i])) return 1;
        if (gemm_timed(m, s, pl.fc2[i])) return 1;
        if (med_timed(m, s, VIDIL_KCLASS_LAYERNORM, 0.0, 10.0 * rows * D, [&] { return layernorm_post_run(b.resid, ly.fln_w.f(), ly.fln_b.f(), b.xn, m->dt, rows, D, eps, s); })) return 1;
    }
    return 0;
}

// Cross-attention keys/values of every layer for every frame: image tokens fp32 -> 16-bit once, then one GEMM per layer.
int run_cross_kv(const vidil_med* m, const float* image_embeds, int F, int Nv, void* img16, void* ckv, cudaStream_t s) {
    const vidil_med_cfg& c = m->cfg;
    const int D = c.hidden, E = c.encoder_width;
    const int64_t Mi = static_cast<int64_t>(F) * Nv;
    if (cast_run(image_embeds, img16, m->dt, Mi, E, E, s)) return 1;
    for (int i = 0; i < c.depth; ++i) {
        const MedLayer& ly = *m->layers[i];
        void* out = reinterpret_cast<uint8_t*>(ckv) + static_cast<size_t>(i) * Mi * 2 * D * 2;
        GemmProblem g = make_gemm(m, EPI_STORE, static_cast<int>(Mi), 2 * D, E, img16, E, ly.ckv_w, &ly.ckv_b, out, 2 * D);
        if (gemm_prepare(g) || gemm_timed(m, s, g)) return 1;
    }
    return 0;
}

// BertOnlyMLMHead (med.py:501-541) on M rows of a 16-bit matrix with row stride lda -> fp32 logits [M, V]
int run_lm_head(const vidil_med* m, const void* A, int64_t lda, int M, void* head_t, void* head_ln, float* logits, cudaStream_t s) {
    const vidil_med_cfg& c = m->cfg;
    const int D = c.hidden, V = c.vocab_size;
    GemmProblem t = make_gemm(m, EPI_GELU, M, D, D, A, lda, m->hd_w, &m->hd_b, head_t, D);
    if (gemm_prepare(t) || gemm_timed(m, s, t)) return 1;
    if (layernorm16_run(head_t, m->hln_w.f(), m->hln_b.f(), head_ln, m->dt, M, D, c.ln_eps, s)) return 1;
    GemmProblem d = make_gemm(m, EPI_STORE_F32, M, V, D, head_ln, D, m->dec_w, &m->dec_b, logits, V);
    if (gemm_prepare(d) || gemm_timed(m, s, d)) return 1;
    return 0;
}

struct ForwardWs {
    StackBufs b;
    void *img16 = nullptr, *ckv = nullptr, *head_t = nullptr, *head_ln = nullptr;
    size_t total = 0;
};

ForwardWs forward_ws(const vidil_med* m, void* base, int n_seq, int T, int F, int Nv) {
    const vidil_med_cfg& c = m->cfg;
    const size_t rows = static_cast<size_t>(n_seq) * T, D = c.hidden;
    Carver cv(base);
    ForwardWs w;
    carve_stack(cv, m, rows, w.b);
    w.img16 = cv.take_bytes(static_cast<size_t>(F) * Nv * c.encoder_width * 2);
    w.ckv = cv.take_bytes(static_cast<size_t>(c.depth) * F * Nv * 2 * D * 2);
    if (c.lm_head) {
        w.head_t = cv.take_bytes(rows * D * 2);
        w.head_ln = cv.take_bytes(rows * D * 2);
    }
    w.total = cv.off;
    return w;
}

struct BeamWs {
    BeamState st;
    float* cand_score = nullptr;
    int32_t* cand_tok = nullptr;
    int32_t* prompt = nullptr;
};

void carve_beam(Carver& cv, int F, int K, int Tm, BeamWs& w) {
    const size_t R = static_cast<size_t>(F) * K;
    w.st.frames = F;
    w.st.beams = K;
    w.st.t_max = Tm;
    w.st.seq = cv.take<int32_t>(2 * R * Tm);
    w.st.anc = cv.take<int32_t>(2 * R * Tm);
    w.st.beam_scores = cv.take<float>(R);
    w.st.cur_tok = cv.take<int32_t>(R);
    w.st.hyp_n = cv.take<int32_t>(F);
    w.st.hyp_score = cv.take<double>(static_cast<size_t>(F) * (K + 1));
    w.st.hyp_len = cv.take<int32_t>(static_cast<size_t>(F) * (K + 1));
    w.st.hyp_tok = cv.take<int32_t>(static_cast<size_t>(F) * (K + 1) * Tm);
    w.st.worst = cv.take<double>(F);
    w.st.done = cv.take<int32_t>(F);
    w.st.n_done = cv.take<int32_t>(1);
    w.cand_score = cv.take<float>(R * 2 * K);
    w.cand_tok = cv.take<int32_t>(R * 2 * K);
    w.prompt = cv.take<int32_t>(Tm);
}

struct GenerateWs {
    StackBufs b;
    void *img16 = nullptr, *ckv = nullptr, *cache = nullptr, *head_t = nullptr, *head_ln = nullptr;
    float* logits = nullptr;
    BeamWs beam;
    size_t total = 0;
};

GenerateWs generate_ws(const vidil_med* m, void* base, int F, int Nv, int K, int Tm, int Lp) {
    const vidil_med_cfg& c = m->cfg;
    const size_t D = c.hidden, R = static_cast<size_t>(F) * K;
    const size_t rows = std::max(R, static_cast<size_t>(F) * Lp);
    Carver cv(base);
    GenerateWs w;
    carve_stack(cv, m, rows, w.b);
    w.img16 = cv.take_bytes(static_cast<size_t>(F) * Nv * c.encoder_width * 2);
    w.ckv = cv.take_bytes(static_cast<size_t>(c.depth) * F * Nv * 2 * D * 2);
    w.cache = cv.take_bytes(static_cast<size_t>(c.depth) * R * Tm * 2 * D * 2);
    w.head_t = cv.take_bytes(R * D * 2);
    w.head_ln = cv.take_bytes(R * D * 2);
    w.logits = cv.take<float>(R * c.vocab_size);
    carve_beam(cv, F, K, Tm, w.beam);
    w.total = cv.off;
    return w;
}

int check_beam_args(int F, int K, int V, int Lp, int max_length, int min_length, const void* prompt) {
    if (F <= 0 || K < 1 || K > 4 || prompt == nullptr || Lp < 1 || max_length <= Lp || max_length > 64 || min_length < 0 || V < 2 * K + 1 ||
        V % 4 != 0) {
        set_error("beam search: n_frames=%d num_beams=%d (1..4) prompt_len=%d max_length=%d (prompt_len < max_length <= 64) V=%d "
                  "(multiple of 4, > 2*num_beams)", F, K, Lp, max_length, V);
        return 1;
    }
    return 0;
}

}  // namespace

extern "C" {

int32_t vidil_med_create(const vidil_med_cfg* cfg, vidil_med** out) {
    if (cfg == nullptr || out == nullptr) {
        set_error("vidil_med_create: null argument");
        return 1;
    }
    *out = nullptr;
    vidil_med_cfg c = *cfg;
    if (c.cta_group == 0) c.cta_group = 2;
    if (c.vocab_size <= 0 || c.vocab_size % 4 != 0 || c.max_positions <= 0) {
        set_error("unsupported text geometry: vocab_size=%d (multiple of 4) max_positions=%d", c.vocab_size, c.max_positions);
        return 1;
    }
    if (c.num_heads * 64 != c.hidden || (c.hidden != 128 && c.hidden != 256 && c.hidden != 512 && c.hidden != 768 && c.hidden != 1024)) {
        set_error("unsupported width: hidden=%d num_heads=%d (head_dim must be 64; LayerNorm covers 128/256/512/768/1024)", c.hidden,
                  c.num_heads);
        return 1;
    }
    if (c.depth <= 0 || c.mlp_dim <= 0 || c.mlp_dim % 64 != 0 || c.encoder_width <= 0 || c.encoder_width % 64 != 0 || c.cls_out < 0 ||
        c.cls_out > 64) {
        set_error("unsupported depth=%d / mlp_dim=%d / encoder_width=%d (multiples of 64) / cls_out=%d", c.depth, c.mlp_dim,
                  c.encoder_width, c.cls_out);
        return 1;
    }
    if ((c.dtype != VIDIL_DTYPE_BF16 && c.dtype != VIDIL_DTYPE_FP16) || (c.cta_group != 1 && c.cta_group != 2)) {
        set_error("unsupported dtype %d / cta_group %d", c.dtype, c.cta_group);
        return 1;
    }
    if (gemm_num_sms() == 0) {
        if (get_error()[0] == 0) set_error("no sm_100 CUDA device available");
        return 1;
    }
    std::unique_ptr<vidil_med> m(new vidil_med());
    m->cfg = c;
    m->dt = (c.dtype == VIDIL_DTYPE_BF16) ? DT_BF16 : DT_FP16;
    for (int i = 0; i < c.depth; ++i) m->layers.emplace_back(new MedLayer());
    std::vector<std::string> names;
    med_param_names(m.get(), names);
    for (auto& n : names) {
        MSlot s;
        if (!med_slot(m.get(), n.c_str(), s)) {
            set_error("internal: no slot for %s", n.c_str());
            return 1;
        }
        if (s.buf->alloc(static_cast<size_t>(s.numel) * (s.matrix ? 2 : 4))) return 1;
    }
    *out = m.release();
    return 0;
}

void vidil_med_destroy(vidil_med* med) { delete med; }

int32_t vidil_med_load(vidil_med* med, const char* name, const float* dev_ptr, int64_t numel, void* stream) {
    if (med == nullptr || name == nullptr || dev_ptr == nullptr) {
        set_error("vidil_med_load: null argument");
        return 1;
    }
    MSlot s;
    if (!med_slot(med, name, s)) {
        set_error("vidil_med_load: unknown parameter '%s'", name);
        return 1;
    }
    if (numel != s.numel) {
        set_error("vidil_med_load: '%s' has %lld elements, expected %lld", name, (long long)numel, (long long)s.numel);
        return 1;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (s.matrix) {
        if (cast_run(dev_ptr, s.buf->p, med->dt, s.rows, s.cols, s.cols, st)) return 1;
    } else {
        VIDIL_CUDA_OK(cudaMemcpyAsync(s.buf->p, dev_ptr, static_cast<size_t>(numel) * 4, cudaMemcpyDeviceToDevice, st));
    }
    s.buf->loaded = true;
    return 0;
}

int32_t vidil_med_check_loaded(const vidil_med* med) {
    if (med == nullptr) {
        set_error("null handle");
        return 1;
    }
    std::vector<std::string> names;
    med_param_names(med, names);
    for (auto& n : names) {
        MSlot s;
        med_slot(const_cast<vidil_med*>(med), n.c_str(), s);
        if (!s.buf->loaded) {
            set_error("parameter '%s' has not been loaded", n.c_str());
            return 1;
        }
    }
    return 0;
}

int32_t vidil_med_set_profiling(vidil_med* med, int32_t enable) {
    if (med == nullptr) {
        set_error("null handle");
        return 1;
    }
    med->profiling = enable != 0;
    return 0;
}

int32_t vidil_med_read_profile(vidil_med* med, vidil_kernel_stats* out) {
    if (med == nullptr || out == nullptr) {
        set_error("vidil_med_read_profile: null argument");
        return 1;
    }
    memset(out, 0, sizeof(*out));
    for (auto& r : med->prof) {
        VIDIL_CUDA_OK(cudaEventSynchronize(r.b));
        float ms = 0.f;
        VIDIL_CUDA_OK(cudaEventElapsedTime(&ms, r.a, r.b));
        out->ms[r.cls] += ms;
        out->flops[r.cls] += r.flops;
        out->bytes[r.cls] += r.bytes;
        out->launches[r.cls] += 1;
        med->free_events.push_back(r.a);
        med->free_events.push_back(r.b);
    }
    med->prof.clear();
    return 0;
}

size_t vidil_med_forward_workspace_bytes(const vidil_med* med, int32_t n_seq, int32_t seq_len, int32_t n_frames, int32_t n_img_tokens) {
    if (med == nullptr || n_seq <= 0 || seq_len <= 0 || n_frames <= 0 || n_img_tokens <= 0) return 0;
    return forward_ws(med, nullptr, n_seq, seq_len, n_frames, n_img_tokens).total;
}

int32_t vidil_med_forward(vidil_med* med, const float* image_embeds, int32_t n_frames, int32_t n_img_tokens, const int32_t* input_ids,
                          const int32_t* attention_mask, const int32_t* frame_of_seq, int32_t seqs_per_frame, int32_t n_seq, int32_t seq_len,
                          int32_t causal, float* out_hidden, float* out_logits, float* out_cls, void* workspace, size_t workspace_bytes,
                          void* stream) {
    if (med == nullptr || image_embeds == nullptr || input_ids == nullptr || workspace == nullptr) {
        set_error("vidil_med_forward: null argument");
        return 1;
    }
    const vidil_med_cfg& c = med->cfg;
    if (n_seq <= 0 || seq_len <= 0 || seq_len > c.max_positions || seq_len > 128 || n_frames <= 0 || n_img_tokens <= 0) {
        set_error("vidil_med_forward: n_seq=%d seq_len=%d (<= min(128, max_positions %d)) n_frames=%d n_img_tokens=%d", n_seq, seq_len,
                  c.max_positions, n_frames, n_img_tokens);
        return 1;
    }
    if (seqs_per_frame > 0 && (frame_of_seq != nullptr || n_seq != n_frames * seqs_per_frame)) {
        set_error("vidil_med_forward: seqs_per_frame=%d needs frame_of_seq == NULL and n_seq == n_frames * seqs_per_frame (%d != %d * %d)",
                  seqs_per_frame, n_seq, n_frames, seqs_per_frame);
        return 1;
    }
    if (seqs_per_frame <= 0 && frame_of_seq == nullptr && n_seq != n_frames) {
        set_error("vidil_med_forward: %d sequences for %d frames need frame_of_seq or seqs_per_frame", n_seq, n_frames);
        return 1;
    }
    if ((out_logits != nullptr && !c.lm_head) || (out_cls != nullptr && c.cls_out <= 0)) {
        set_error("vidil_med_forward: this handle has no %s", out_logits ? "LM head" : "cls head");
        return 1;
    }
    if (vidil_med_check_loaded(med)) return 1;
    ForwardWs w = forward_ws(med, workspace, n_seq, seq_len, n_frames, n_img_tokens);
    if (workspace_bytes < w.total || (reinterpret_cast<uintptr_t>(workspace) & (ALIGN - 1)) ||
        (reinterpret_cast<uintptr_t>(out_logits) & 15)) {
        set_error("vidil_med_forward: workspace too small (%zu < %zu) or not %zu-byte aligned, or logits not 16-byte aligned",
                  workspace_bytes, w.total, ALIGN);
        return 1;
    }
    if (gemm_num_sms() == 0) return 1;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const int rows = n_seq * seq_len, D = c.hidden;
    if (run_cross_kv(med, image_embeds, n_frames, n_img_tokens, w.img16, w.ckv, s)) return 1;
    if (med_embed_run(input_ids, med->word.f(), med->pos.f(), w.b.resid, rows, seq_len, 0, 0, D, c.vocab_size, c.max_positions, s))
        return 1;
    StackPlan pl;
    if (plan_stack(med, pl, w.b, rows)) return 1;
    AttnArgs a;
    a.mode = causal ? MED_ATTN_CAUSAL : MED_ATTN_FULL;
    a.T_seq = seq_len;
    a.groups = seqs_per_frame > 0 ? n_frames : n_seq;          // cross-attention query groups
    a.nq = seqs_per_frame > 0 ? seqs_per_frame * seq_len : seq_len;
    a.frame_of_group = frame_of_seq;
    a.cross_kv = w.ckv;
    a.cross_layer_elems = static_cast<size_t>(n_frames) * n_img_tokens * 2 * D;
    a.Nv = n_img_tokens;
    a.mask = causal ? nullptr : attention_mask;
    if (run_stack(med, pl, w.b, a, s)) return 1;
    if (out_hidden)
        VIDIL_CUDA_OK(cudaMemcpyAsync(out_hidden, w.b.resid, static_cast<size_t>(rows) * D * 4, cudaMemcpyDeviceToDevice, s));
    if (out_logits && run_lm_head(med, w.b.xn, D, rows, w.head_t, w.head_ln, out_logits, s)) return 1;
    if (out_cls && med_cls_head_run(w.b.resid, med->cls_w.f(), med->cls_b.f(), out_cls, n_seq, seq_len, D, c.cls_out, s)) return 1;
    return 0;
}

size_t vidil_med_generate_workspace_bytes(const vidil_med* med, int32_t n_frames, int32_t n_img_tokens, int32_t num_beams,
                                          int32_t max_length, int32_t prompt_len) {
    if (med == nullptr || n_frames <= 0 || n_img_tokens <= 0 || num_beams <= 0 || max_length <= 0 || prompt_len <= 0) return 0;
    return generate_ws(med, nullptr, n_frames, n_img_tokens, num_beams, max_length, prompt_len).total;
}

}  // extern "C"

namespace {

// nucleus sampling instead of beam search (blip.py:139-148): one sequence per frame, the draw of step i of frame f uses
// uniforms[i * n_frames + f]
struct SampleCfg {
    int top_k;
    float top_p, repetition_penalty;
    const float* uniforms;
};

int generate_core(vidil_med* med, const float* image_embeds, int32_t n_frames, int32_t n_img_tokens, const int32_t* prompt_ids_host,
                  int32_t prompt_len, int32_t num_beams, int32_t max_length, int32_t min_length, int32_t eos_token, int32_t pad_token,
                  float length_penalty, const SampleCfg* smp, int32_t* out_tokens, int32_t* out_lengths, float* out_scores,
                  void* workspace, size_t workspace_bytes, void* stream) {
    if (med == nullptr || image_embeds == nullptr || out_tokens == nullptr || out_lengths == nullptr || out_scores == nullptr ||
        workspace == nullptr) {
        set_error("vidil_med_generate: null argument");
        return 1;
    }
    const vidil_med_cfg& c = med->cfg;
    if (!c.lm_head) {
        set_error("vidil_med_generate: this handle has no LM head");
        return 1;
    }
    const int F = n_frames, K = num_beams, Tm = max_length, Lp = prompt_len, V = c.vocab_size, D = c.hidden;
    if (check_beam_args(F, K, V, Lp, max_length, min_length, prompt_ids_host)) return 1;
    if (n_img_tokens <= 0 || max_length > c.max_positions) {
        set_error("vidil_med_generate: n_img_tokens=%d max_length=%d (max_positions %d)", n_img_tokens, max_length, c.max_positions);
        return 1;
    }
    if (vidil_med_check_loaded(med)) return 1;
    GenerateWs w = generate_ws(med, workspace, F, n_img_tokens, K, Tm, Lp);
    if (workspace_bytes < w.total || (reinterpret_cast<uintptr_t>(workspace) & (ALIGN - 1))) {
        set_error("vidil_med_generate: workspace too small (%zu < %zu) or not %zu-byte aligned", workspace_bytes, w.total, ALIGN);
        return 1;
    }
    if (gemm_num_sms() == 0) return 1;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    BeamState& st = w.beam.st;
    st.eos = eos_token;
    st.pad = pad_token;
    st.length_penalty = length_penalty;
    const int R = F * K, nc = 2 * K;

    VIDIL_CUDA_OK(cudaMemcpyAsync(w.beam.prompt, prompt_ids_host, static_cast<size_t>(Lp) * 4, cudaMemcpyHostToDevice, s));
    if (med_beam_init_run(st, w.beam.prompt, Lp, s)) return 1;
    if (run_cross_kv(med, image_embeds, F, n_img_tokens, w.img16, w.ckv, s)) return 1;

    AttnArgs a;
    a.groups = F;
    a.cross_kv = w.ckv;
    a.cross_layer_elems = static_cast<size_t>(F) * n_img_tokens * 2 * D;
    a.Nv = n_img_tokens;
    a.cache = w.cache;
    a.cache_layer_elems = static_cast<size_t>(R) * Tm * 2 * D;
    a.beams = K;
    a.Tmax = Tm;

    // the prompt, once per frame (the beams of a frame are identical until the first expansion)
    {
        const int rows = F * Lp;
        StackPlan pl;
        if (plan_stack(med, pl, w.b, rows)) return 1;
        if (med_embed_run(w.beam.prompt, med->word.f(), med->pos.f(), w.b.resid, rows, Lp, 0, Lp, D, V, c.max_positions, s)) return 1;
        a.mode = MED_ATTN_CAUSAL;
        a.T_seq = Lp;
        a.nq = Lp;
        a.anc = st.anc;
        if (run_stack(med, pl, w.b, a, s)) return 1;
        const uint8_t* last = reinterpret_cast<const uint8_t*>(w.b.xn) + static_cast<size_t>(Lp - 1) * D * 2;
        if (run_lm_head(med, last, static_cast<int64_t>(Lp) * D, F, w.head_t, w.head_ln, w.logits, s)) return 1;
        if (smp) {
            if (med_sample_step_run(st, w.logits, V, V, Lp, Lp < min_length ? eos_token : -1, smp->top_k, smp->top_p, smp->repetition_penalty,
                                    smp->uniforms, s))
                return 1;
        } else {
            if (med_logits_topk_run(w.logits, V, 1, nullptr, F, V, nc, Lp < min_length ? eos_token : -1, w.beam.cand_score, w.beam.cand_tok, s))
                return 1;
            if (med_beam_step_run(st, w.beam.cand_score, w.beam.cand_tok, 1, nc, V, Lp, 0, s)) return 1;
        }
    }
    int parity = 1, cur_len = Lp + 1;
    if (cur_len < max_length) {
        StackPlan pl;
        if (plan_stack(med, pl, w.b, R)) return 1;
        CrossKvMap kv_map;
        if (med_cross_kv_map_prepare(kv_map, w.ckv, c.depth, F, n_img_tokens, c.num_heads)) return 1;
        a.kv_map = &kv_map;
        PdlScope pdl;
        a.mode = MED_ATTN_DECODE;
        a.T_seq = 1;
        a.nq = K;
        for (; cur_len < max_length; ++cur_len) {
            a.pos = cur_len - 1;
            a.anc = st.anc + static_cast<size_t>(parity) * R * Tm;
            if (med_embed_run(st.cur_tok, med->word.f(), med->pos.f(), w.b.resid, R, 1, cur_len - 1, 0, D, V, c.max_positions, s)) return 1;
            if (run_stack(med, pl, w.b, a, s)) return 1;
            if (run_lm_head(med, w.b.xn, D, R, w.head_t, w.head_ln, w.logits, s)) return 1;
            if (smp) {
                if (med_timed(med, s, VIDIL_KCLASS_OTHER, 0.0, 8.0 * R * V, [&] {
                        return med_sample_step_run(st, w.logits, V, V, cur_len, cur_len < min_length ? eos_token : -1, smp->top_k, smp->top_p,
                                                   smp->repetition_penalty, smp->uniforms + static_cast<size_t>(cur_len - Lp) * F, s);
                    }))
                    return 1;
            } else {
                if (med_timed(med, s, VIDIL_KCLASS_OTHER, 0.0, 4.0 * R * V, [&] {
                        return med_logits_topk_run(w.logits, V, 1, st.beam_scores, R, V, nc, cur_len < min_length ? eos_token : -1,
                                                   w.beam.cand_score, w.beam.cand_tok, s);
                    }))
                    return 1;
                if (med_beam_step_run(st, w.beam.cand_score, w.beam.cand_tok, K, nc, V, cur_len, parity, s)) return 1;
            }
            parity ^= 1;
            // `if beam_scorer.is_done: break` of transformers' beam_search (`if unfinished_sequences.max() == 0: break` of sample()):
            // every other step from min_length on (no frame can finish earlier) the host reads the count of finished frames; the
            // remaining steps would only pad.  This synchronises the stream — once per ~8 ms of queued work.
            if (cur_len >= min_length && cur_len + 1 < max_length && ((cur_len - min_length) & 1) == 0) {
                int32_t n_done = 0;
                VIDIL_CUDA_OK(cudaMemcpyAsync(&n_done, st.n_done, sizeof(n_done), cudaMemcpyDeviceToHost, s));
                VIDIL_CUDA_OK(cudaStreamSynchronize(s));
                if (n_done >= F) {
                    ++cur_len;
                    break;
                }
            }
        }
    }
    if (smp) return med_sample_finalize_run(st, cur_len, max_length, out_tokens, out_lengths, out_scores, s);
    return med_beam_finalize_run(st, cur_len, parity, max_length, out_tokens, out_lengths, out_scores, s);
}

}  // namespace

extern "C" {

int32_t vidil_med_generate(vidil_med* med, const float* image_embeds, int32_t n_frames, int32_t n_img_tokens,
                           const int32_t* prompt_ids_host, int32_t prompt_len, int32_t num_beams, int32_t max_length,
                           int32_t min_length, int32_t eos_token, int32_t pad_token, float length_penalty, int32_t* out_tokens,
                           int32_t* out_lengths, float* out_scores, void* workspace, size_t workspace_bytes, void* stream) {
    return generate_core(med, image_embeds, n_frames, n_img_tokens, prompt_ids_host, prompt_len, num_beams, max_length, min_length,
                         eos_token, pad_token, length_penalty, nullptr, out_tokens, out_lengths, out_scores, workspace, workspace_bytes,
                         stream);
}

int32_t vidil_med_sample(vidil_med* med, const float* image_embeds, int32_t n_frames, int32_t n_img_tokens,
                         const int32_t* prompt_ids_host, int32_t prompt_len, int32_t max_length, int32_t min_length, int32_t eos_token,
                         int32_t pad_token, int32_t top_k, float top_p, float repetition_penalty, const float* uniforms,
                         int32_t* out_tokens, int32_t* out_lengths, float* out_scores, void* workspace, size_t workspace_bytes,
                         void* stream) {
    if (uniforms == nullptr) {
        set_error("vidil_med_sample: null argument");
        return 1;
    }
    if (top_k < 1 || top_k > 1024 || !(top_p > 0.f) || !(repetition_penalty > 0.f)) {
        set_error("vidil_med_sample: top_k=%d (1..1024), top_p=%g (> 0), repetition_penalty=%g (> 0)", top_k, top_p, repetition_penalty);
        return 1;
    }
    const SampleCfg smp{top_k, top_p, repetition_penalty, uniforms};
    return generate_core(med, image_embeds, n_frames, n_img_tokens, prompt_ids_host, prompt_len, 1, max_length, min_length, eos_token,
                         pad_token, 1.0f, &smp, out_tokens, out_lengths, out_scores, workspace, workspace_bytes, stream);
}

size_t vidil_op_beam_search_workspace_bytes(int32_t n_frames, int32_t num_beams, int32_t max_length) {
    if (n_frames <= 0 || num_beams <= 0 || max_length <= 0) return 0;
    Carver cv(nullptr);
    BeamWs w;
    carve_beam(cv, n_frames, num_beams, max_length, w);
    return cv.off;
}

int32_t vidil_op_beam_search(const float* step_logits, int32_t n_steps, int32_t n_frames, int32_t num_beams, int32_t V,
                             const int32_t* prompt_ids_host, int32_t prompt_len, int32_t max_length, int32_t min_length,
                             int32_t eos_token, int32_t pad_token, float length_penalty, int32_t* out_tokens, int32_t* out_lengths,
                             float* out_scores, void* workspace, size_t workspace_bytes, void* stream) {
    if (step_logits == nullptr || out_tokens == nullptr || out_lengths == nullptr || out_scores == nullptr || workspace == nullptr) {
        set_error("vidil_op_beam_search: null argument");
        return 1;
    }
    const int F = n_frames, K = num_beams, Lp = prompt_len;
    if (check_beam_args(F, K, V, Lp, max_length, min_length, prompt_ids_host)) return 1;
    if (n_steps != max_length - Lp) {
        set_error("vidil_op_beam_search: %d steps given, max_length - prompt_len = %d needed", n_steps, max_length - Lp);
        return 1;
    }
    Carver cv(workspace);
    BeamWs w;
    carve_beam(cv, F, K, max_length, w);
    if (workspace_bytes < cv.off || (reinterpret_cast<uintptr_t>(workspace) & (ALIGN - 1))) {
        set_error("vidil_op_beam_search: workspace too small (%zu < %zu) or not %zu-byte aligned", workspace_bytes, cv.off, ALIGN);
        return 1;
    }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    BeamState& st = w.st;
    st.eos = eos_token;
    st.pad = pad_token;
    st.length_penalty = length_penalty;
    const int R = F * K, nc = 2 * K;
    VIDIL_CUDA_OK(cudaMemcpyAsync(w.prompt, prompt_ids_host, static_cast<size_t>(Lp) * 4, cudaMemcpyHostToDevice, s));
    if (med_beam_init_run(st, w.prompt, Lp, s)) return 1;
    int parity = 0, cur_len = Lp;
    for (int step = 0; step < n_steps; ++step, ++cur_len) {
        const float* lg = step_logits + static_cast<size_t>(step) * R * V;
        const int ban = cur_len < min_length ? eos_token : -1;
        if (step == 0) {
            if (med_logits_topk_run(lg, V, K, nullptr, F, V, nc, ban, w.cand_score, w.cand_tok, s)) return 1;
            if (med_beam_step_run(st, w.cand_score, w.cand_tok, 1, nc, V, cur_len, parity, s)) return 1;
        } else {
            if (med_logits_topk_run(lg, V, 1, st.beam_scores, R, V, nc, ban, w.cand_score, w.cand_tok, s)) return 1;
            if (med_beam_step_run(st, w.cand_score, w.cand_tok, K, nc, V, cur_len, parity, s)) return 1;
        }
        parity ^= 1;
    }
    return med_beam_finalize_run(st, cur_len, parity, max_length, out_tokens, out_lengths, out_scores, s);
}

int32_t vidil_op_sample(const float* step_logits, int32_t n_steps, int32_t n_frames, int32_t V, const int32_t* prompt_ids_host,
                        int32_t prompt_len, int32_t max_length, int32_t min_length, int32_t eos_token, int32_t pad_token, int32_t top_k,
                        float top_p, float repetition_penalty, const float* uniforms, int32_t* out_tokens, int32_t* out_lengths,
                        float* out_scores, void* workspace, size_t workspace_bytes, void* stream) {
    if (step_logits == nullptr || uniforms == nullptr || out_tokens == nullptr || out_lengths == nullptr || out_scores == nullptr ||
        workspace == nullptr) {
        set_error("vidil_op_sample: null argument");
        return 1;
    }
    const int F = n_frames, Lp = prompt_len;
    if (check_beam_args(F, 1, V, Lp, max_length, min_length, prompt_ids_host)) return 1;
    if (n_steps != max_length - Lp) {
        set_error("vidil_op_sample: %d steps given, max_length - prompt_len = %d needed", n_steps, max_length - Lp);
        return 1;
    }
    Carver cv(workspace);
    BeamWs w;
    carve_beam(cv, F, 1, max_length, w);
    if (workspace_bytes < cv.off || (reinterpret_cast<uintptr_t>(workspace) & (ALIGN - 1))) {
        set_error("vidil_op_sample: workspace too small (%zu < %zu) or not %zu-byte aligned", workspace_bytes, cv.off, ALIGN);
        return 1;
    }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    BeamState& st = w.st;
    st.eos = eos_token;
    st.pad = pad_token;
    st.length_penalty = 1.0f;
    VIDIL_CUDA_OK(cudaMemcpyAsync(w.prompt, prompt_ids_host, static_cast<size_t>(Lp) * 4, cudaMemcpyHostToDevice, s));
    if (med_beam_init_run(st, w.prompt, Lp, s)) return 1;
    int cur_len = Lp;
    for (int step = 0; step < n_steps; ++step, ++cur_len) {
        if (med_sample_step_run(st, step_logits + static_cast<size_t>(step) * F * V, V, V, cur_len, cur_len < min_length ? eos_token : -1, top_k,
                                top_p, repetition_penalty, uniforms + static_cast<size_t>(step) * F, s))
            return 1;
    }
    return med_sample_finalize_run(st, cur_len, max_length, out_tokens, out_lengths, out_scores, s);
}

}  // extern "C"
