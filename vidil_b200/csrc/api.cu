// C ABI (include/vidil_b200.h): encoder handle, weight packing, forward plans, similarity/top-k and the
// operator-level entry points used by the parity tests.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <memory>
#include <string>
#include <vector>

#include "../../include/vidil_b200.h"
#include "kernels.h"

namespace vidil {

static thread_local std::string g_error;
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
}
const char* get_error() { return g_error.c_str(); }
void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int64_t launch_count() { return g_launches.load(std::memory_order_relaxed); }
// PDL is off unless VIDIL_PDL=1, or a caller that measured a gain turns it on for its scope (the caption decoder's chain of
// ~2400 short kernels; on the ViT path, whose kernels run 50-300 us each under the power cap, it costs 1 %).
static const int g_pdl_env = [] {
    const char* e = getenv("VIDIL_PDL");
    return e == nullptr ? -1 : (e[0] == '0' ? 0 : 1);
}();
static thread_local int g_pdl_scope = 0;
bool pdl_enabled() { return g_pdl_env >= 0 ? g_pdl_env == 1 : g_pdl_scope > 0; }
void pdl_scope(int delta) { g_pdl_scope += delta; }

namespace {

constexpr size_t ALIGN = 1024;
inline size_t align_up(size_t x) { return (x + ALIGN - 1) / ALIGN * ALIGN; }
inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    bool loaded = false;
    ~DevBuf() {
        if (p) cudaFree(p);
    }
    int alloc(size_t n) {
        VIDIL_CUDA_OK(cudaMalloc(&p, n));
        bytes = n;
        return 0;
    }
};

struct Layer {
    DevBuf ln1_w, ln1_b, qkv_w, qkv_b, proj_w, proj_b, ln2_w, ln2_b, fc1_w, fc1_b, fc2_w, fc2_b;
};

// All GEMMs of one forward at a given batch size and workspace address, TMA maps already encoded.
struct Plan {
    int batch = 0;
    const void* ws = nullptr;
    uint64_t stamp = 0;
    float* resid = nullptr;
    void* xn = nullptr;
    void* qkv = nullptr;
    void* attn = nullptr;
    void* hidden = nullptr;
    void* patches = nullptr;  // aliases hidden
    float* pre = nullptr;     // CLIP: pre-LayerNorm fp32 embeddings, aliases qkv
    void* cls_ln = nullptr;
    float* head_out = nullptr;
    GemmProblem patch;
    std::vector<GemmProblem> qkv_g, proj_g, fc1_g, fc2_g;
    GemmProblem head;
    bool attn_tc = false;  // tcgen05 attention (tokens <= 208); otherwise the streaming mma.sync kernel
    AttentionMaps attn_maps;
};

}  // namespace
}  // namespace vidil

using namespace vidil;

// Event-pair records of the launches enqueued while profiling is on (read back by vidil_encoder_read_profile).
struct ProfRec {
    cudaEvent_t a, b;
    int cls;
    double flops, bytes;
};

struct vidil_encoder {
    vidil_encoder_cfg cfg;
    int device = 0;
    int grid = 0, patches = 0, tokens = 0, kpatch = 0, kpatch_pad = 0;
    DType dt = DT_BF16;
    DevBuf cls, pos, patch_w, patch_b, pre_w, pre_b, norm_w, norm_b, head_w;
    std::vector<std::unique_ptr<Layer>> layers;
    std::vector<std::unique_ptr<Plan>> plans;
    uint64_t clock = 0;
    bool profiling = false;
    std::vector<ProfRec> prof;
    std::vector<cudaEvent_t> free_events;
    // host-buffer pipeline (vidil_encoder_host_submit / _wait): copy engines run on their own streams
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_compute[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    cudaEvent_t ev_bind = nullptr;
    // geometry of the two-slot pipeline, fixed when a scratch buffer is bound: slot offsets must not move while the other slot
    // is in flight, so they are derived from the capacity batch, never from the batch of the current call
    void* pipe_scratch = nullptr;
    size_t pipe_bytes = 0;
    int pipe_cap = 0;
    ~vidil_encoder() {
        if (copy_in) cudaStreamDestroy(copy_in);
        if (copy_out) cudaStreamDestroy(copy_out);
        if (ev_bind) cudaEventDestroy(ev_bind);
        for (int i = 0; i < 2; ++i) {
            if (ev_in[i]) cudaEventDestroy(ev_in[i]);
            if (ev_compute[i]) cudaEventDestroy(ev_compute[i]);
            if (ev_out[i]) cudaEventDestroy(ev_out[i]);
        }
        for (auto& r : prof) {
            cudaEventDestroy(r.a);
            cudaEventDestroy(r.b);
        }
        for (auto ev : free_events) cudaEventDestroy(ev);
    }
};

namespace {

struct WsLayout {
    size_t resid, xn, qkv, attn, hidden, cls_ln, head_out, total;
};

WsLayout ws_layout(const vidil_encoder* e, int B) {
    const size_t M = static_cast<size_t>(B) * e->tokens;
    const size_t D = e->cfg.embed_dim;
    WsLayout w;
    size_t off = 0;
    w.resid = off;
    off += align_up(M * D * 4);
    w.xn = off;
    off += align_up(M * D * 2);
    w.qkv = off;
    off += align_up(M * 3 * D * 2);
    w.attn = off;
    off += align_up(M * D * 2);
    size_t hid = M * e->cfg.mlp_dim * 2;
    const size_t pat = static_cast<size_t>(B) * e->patches * e->kpatch_pad * 2;
    if (pat > hid) hid = pat;
    w.hidden = off;
    off += align_up(hid);
    w.cls_ln = off;
    off += align_up(static_cast<size_t>(B) * D * 2);
    w.head_out = off;
    off += align_up(static_cast<size_t>(B) * (e->cfg.proj_dim > 0 ? e->cfg.proj_dim : 1) * 4);
    w.total = off;
    return w;
}

int build_plan(vidil_encoder* e, Plan& pl, int B, void* ws) {
    const vidil_encoder_cfg& c = e->cfg;
    const WsLayout L = ws_layout(e, B);
    uint8_t* base = reinterpret_cast<uint8_t*>(ws);
    pl.batch = B;
    pl.ws = ws;
    pl.resid = reinterpret_cast<float*>(base + L.resid);
    pl.xn = base + L.xn;
    pl.qkv = base + L.qkv;
    pl.attn = base + L.attn;
    pl.hidden = base + L.hidden;
    pl.patches = pl.hidden;
    pl.pre = reinterpret_cast<float*>(pl.qkv);
    pl.cls_ln = base + L.cls_ln;
    pl.head_out = reinterpret_cast<float*>(base + L.head_out);
    const int M = B * e->tokens, D = c.embed_dim;
    const int cg = c.cta_group;

    GemmProblem g;
    g.dt = e->dt;
    g.cta_group = cg;

    // patch embedding: [B*P, Kpad] x [D, Kpad]^T -> token rows 1..P of every frame (+bias +pos)
    pl.patch = g;
    pl.patch.epi = EPI_PATCH;
    pl.patch.M = B * e->patches;
    pl.patch.N = D;
    pl.patch.K = e->kpatch_pad;
    pl.patch.A = pl.patches;
    pl.patch.lda = e->kpatch_pad;
    pl.patch.W = e->patch_w.p;
    pl.patch.ldw = e->kpatch_pad;
    pl.patch.bias = c.patch_bias ? reinterpret_cast<const float*>(e->patch_b.p) : nullptr;
    pl.patch.out = c.pre_ln ? pl.pre : pl.resid;
    pl.patch.ldo = D;
    pl.patch.pos = reinterpret_cast<const float*>(e->pos.p);
    pl.patch.patches_per_frame = e->patches;
    if (gemm_prepare(pl.patch)) return 1;

    pl.qkv_g.assign(c.depth, g);
    pl.proj_g.assign(c.depth, g);
    pl.fc1_g.assign(c.depth, g);
    pl.fc2_g.assign(c.depth, g);
    for (int i = 0; i < c.depth; ++i) {
        Layer& ly = *e->layers[i];
        GemmProblem& q = pl.qkv_g[i];
        q.epi = EPI_STORE;
        q.M = M; q.N = 3 * D; q.K = D;
        q.A = pl.xn; q.lda = D;
        q.W = ly.qkv_w.p; q.ldw = D;
        q.bias = reinterpret_cast<const float*>(ly.qkv_b.p);
        q.out = pl.qkv; q.ldo = 3 * D;
        if (gemm_prepare(q)) return 1;

        GemmProblem& p = pl.proj_g[i];
        p.epi = EPI_RESID;
        p.M = M; p.N = D; p.K = D;
        p.A = pl.attn; p.lda = D;
        p.W = ly.proj_w.p; p.ldw = D;
        p.bias = reinterpret_cast<const float*>(ly.proj_b.p);
        p.out = pl.resid; p.ldo = D;
        if (gemm_prepare(p)) return 1;

        GemmProblem& f1 = pl.fc1_g[i];
        f1.epi = (c.act == VIDIL_ACT_QUICK_GELU) ? EPI_QUICKGELU : EPI_GELU;
        f1.M = M; f1.N = c.mlp_dim; f1.K = D;
        f1.A = pl.xn; f1.lda = D;
        f1.W = ly.fc1_w.p; f1.ldw = D;
        f1.bias = reinterpret_cast<const float*>(ly.fc1_b.p);
        f1.out = pl.hidden; f1.ldo = c.mlp_dim;
        if (gemm_prepare(f1)) return 1;

        GemmProblem& f2 = pl.fc2_g[i];
        f2.epi = EPI_RESID;
        f2.M = M; f2.N = D; f2.K = c.mlp_dim;
        f2.A = pl.hidden; f2.lda = c.mlp_dim;
        f2.W = ly.fc2_w.p; f2.ldw = c.mlp_dim;
        f2.bias = reinterpret_cast<const float*>(ly.fc2_b.p);
        f2.out = pl.resid; f2.ldo = D;
        if (gemm_prepare(f2)) return 1;
    }
    if (attention_tc_supported(e->tokens)) {
        pl.attn_tc = true;
        if (attention_tc_prepare(pl.attn_maps, pl.qkv, pl.attn, e->dt, B, e->tokens, c.num_heads)) return 1;
    } else if (attention_tc257_supported(e->tokens)) {
        pl.attn_tc = true;
        if (attention_tc257_prepare(pl.attn_maps, pl.qkv, pl.attn, e->dt, B, e->tokens, c.num_heads)) return 1;
    } else if (attention_tcl_supported(e->tokens)) {
        pl.attn_tc = true;
        if (attention_tcl_prepare(pl.attn_maps, pl.qkv, pl.attn, e->dt, B, e->tokens, c.num_heads)) return 1;
    }
    if (c.proj_dim > 0) {
        pl.head = g;
        pl.head.epi = EPI_STORE_F32;
        pl.head.M = B; pl.head.N = c.proj_dim; pl.head.K = D;
        pl.head.A = pl.cls_ln; pl.head.lda = D;
        pl.head.W = e->head_w.p; pl.head.ldw = D;
        pl.head.bias = nullptr;
        pl.head.out = pl.head_out; pl.head.ldo = c.proj_dim;
        if (gemm_prepare(pl.head)) return 1;
    }
    return 0;
}

Plan* get_plan(vidil_encoder* e, int B, void* ws) {
    for (auto& p : e->plans)
        if (p->batch == B && p->ws == ws) {
            p->stamp = ++e->clock;
            return p.get();
        }
    std::unique_ptr<Plan> pl(new Plan());
    if (build_plan(e, *pl, B, ws)) return nullptr;
    pl->stamp = ++e->clock;
    if (e->plans.size() >= 8) {  // evict the least recently used plan
        size_t victim = 0;
        for (size_t i = 1; i < e->plans.size(); ++i)
            if (e->plans[i]->stamp < e->plans[victim]->stamp) victim = i;
        e->plans[victim] = std::move(pl);
        return e->plans[victim].get();
    }
    e->plans.push_back(std::move(pl));
    return e->plans.back().get();
}

int check_common(vidil_encoder* e, const void* frames, int B, const void* out, void* ws, size_t ws_bytes) {
    if (e == nullptr || frames == nullptr || out == nullptr || ws == nullptr) {
        set_error("null argument");
        return 1;
    }
    if (B <= 0) {
        set_error("batch must be positive, got %d", B);
        return 1;
    }
    if (vidil_encoder_check_loaded(e)) return 1;
    const size_t need = ws_layout(e, B).total;
    if (ws_bytes < need) {
        set_error("workspace too small: %zu bytes given, %zu needed for batch %d", ws_bytes, need, B);
        return 1;
    }
    if ((reinterpret_cast<uintptr_t>(ws) & (ALIGN - 1)) || (reinterpret_cast<uintptr_t>(frames) & 15) ||
        (reinterpret_cast<uintptr_t>(out) & 15)) {
        set_error("workspace must be %zu-byte aligned and frames/out 16-byte aligned", ALIGN);
        return 1;
    }
    if (gemm_num_sms() == 0) {
        if (get_error()[0] == 0) set_error("no sm_100 CUDA device available");
        return 1;
    }
    return 0;
}

int prof_event(vidil_encoder* e, cudaEvent_t* ev) {
    if (!e->free_events.empty()) {
        *ev = e->free_events.back();
        e->free_events.pop_back();
        return 0;
    }
    VIDIL_CUDA_OK(cudaEventCreate(ev));
    return 0;
}

// Runs one launcher; with profiling on, brackets it with an event pair on the same stream.
template <typename Fn>
int timed(vidil_encoder* e, cudaStream_t s, int cls, double flops, double bytes, Fn&& fn) {
    if (!e->profiling) return fn();
    ProfRec r{nullptr, nullptr, cls, flops, bytes};
    if (prof_event(e, &r.a) || prof_event(e, &r.b)) return 1;
    VIDIL_CUDA_OK(cudaEventRecord(r.a, s));
    const int rc = fn();
    VIDIL_CUDA_OK(cudaEventRecord(r.b, s));
    e->prof.push_back(r);
    return rc;
}

inline double gemm_flops(const GemmProblem& g) { return 2.0 * g.M * g.N * g.K; }
// operands once + output once (fp32 read-modify-write for the residual epilogue)
inline double gemm_bytes(const GemmProblem& g) {
    const double out = (g.epi == EPI_RESID) ? 8.0 : (g.epi == EPI_PATCH || g.epi == EPI_STORE_F32) ? 4.0 : 2.0;
    return 2.0 * g.M * g.K + 2.0 * g.N * g.K + out * g.M * g.N;
}

int ln_timed(vidil_encoder* e, cudaStream_t s, const float* in, int64_t stride, const DevBuf& w, const DevBuf& b, void* out,
             bool out_f32, int rows) {
    const int D = e->cfg.embed_dim;
    return timed(e, s, VIDIL_KCLASS_LAYERNORM, 8.0 * rows * D, static_cast<double>(rows) * D * (4 + (out_f32 ? 4 : 2)), [&] {
        return layernorm_run(in, stride, reinterpret_cast<const float*>(w.p), reinterpret_cast<const float*>(b.p), out,
                             out_f32, e->dt, rows, D, e->cfg.ln_eps, s);
    });
}

int gemm_timed(vidil_encoder* e, cudaStream_t s, const GemmProblem& g) {
    return timed(e, s, VIDIL_KCLASS_GEMM, gemm_flops(g), gemm_bytes(g), [&] { return gemm_run(g, s); });
}

// Everything up to (not including) the final LayerNorm: leaves the last block's output in pl.resid.
int run_trunk(vidil_encoder* e, Plan& pl, const float* frames, cudaStream_t s) {
    const vidil_encoder_cfg& c = e->cfg;
    const int B = pl.batch, M = B * e->tokens, D = c.embed_dim;
    const double patch_elems = static_cast<double>(B) * e->patches * e->kpatch_pad;
    if (timed(e, s, VIDIL_KCLASS_OTHER, 0.0, 4.0 * B * 3 * c.img_size * c.img_size + 2.0 * patch_elems, [&]() -> int {
            if (e->kpatch_pad != e->kpatch)  // zero the K padding once per call (CLIP 588 -> 640)
                VIDIL_CUDA_OK(cudaMemsetAsync(pl.patches, 0, static_cast<size_t>(patch_elems) * 2, s));
            return im2col_run(frames, pl.patches, e->dt, B, 3, c.img_size, c.patch_size, e->kpatch_pad, s);
        }))
        return 1;
    float* emb = c.pre_ln ? pl.pre : pl.resid;
    if (timed(e, s, VIDIL_KCLASS_OTHER, 0.0, 4.0 * B * D, [&] {
            return cls_pos_run(reinterpret_cast<const float*>(e->cls.p), reinterpret_cast<const float*>(e->pos.p), emb, B,
                               e->tokens, D, s);
        }))
        return 1;
    if (gemm_timed(e, s, pl.patch)) return 1;
    if (c.pre_ln && ln_timed(e, s, pl.pre, D, e->pre_w, e->pre_b, pl.resid, true, M)) return 1;
    const float scale = 0.125f;  // head_dim ** -0.5 with head_dim = 64 (vit.py:49)
    const double attn_flops = 4.0 * B * c.num_heads * static_cast<double>(e->tokens) * e->tokens * 64;
    const double attn_bytes = 2.0 * M * 4.0 * D;  // qkv read once, out written once
    for (int i = 0; i < c.depth; ++i) {
        Layer& ly = *e->layers[i];
        if (ln_timed(e, s, pl.resid, D, ly.ln1_w, ly.ln1_b, pl.xn, false, M)) return 1;
        if (gemm_timed(e, s, pl.qkv_g[i])) return 1;
        if (timed(e, s, VIDIL_KCLASS_ATTENTION, attn_flops, attn_bytes,
                  [&] {
                      if (!pl.attn_tc) return attention_run(pl.qkv, pl.attn, e->dt, B, e->tokens, c.num_heads, scale, s);
                      if (pl.attn_maps.is_257) return attention_tc257_run(pl.attn_maps, scale, s);
                      return pl.attn_maps.is_long ? attention_tcl_run(pl.attn_maps, scale, s)
                                                  : attention_tc_run(pl.attn_maps, scale, s);
                  }))
            return 1;
        if (gemm_timed(e, s, pl.proj_g[i])) return 1;
        if (ln_timed(e, s, pl.resid, D, ly.ln2_w, ly.ln2_b, pl.xn, false, M)) return 1;
        if (gemm_timed(e, s, pl.fc1_g[i])) return 1;
        if (gemm_timed(e, s, pl.fc2_g[i])) return 1;
    }
    return 0;
}

struct Slot {
    DevBuf* buf;
    int64_t numel;
    bool matrix;
    int rows, cols, ld;  // matrices: logical [rows, cols] packed into [rows, ld]
};

bool find_slot(vidil_encoder* e, const char* name, Slot& s) {
    const vidil_encoder_cfg& c = e->cfg;
    const int D = c.embed_dim, H = c.mlp_dim;
    auto vec = [&](DevBuf& b, int64_t n) { s = Slot{&b, n, false, 0, 0, 0}; return true; };
    auto mat = [&](DevBuf& b, int r, int k, int ld) { s = Slot{&b, static_cast<int64_t>(r) * k, true, r, k, ld}; return true; };
    if (!strcmp(name, "cls_token")) return vec(e->cls, D);
    if (!strcmp(name, "pos_embed")) return vec(e->pos, static_cast<int64_t>(e->tokens) * D);
    if (!strcmp(name, "patch_embed.proj.weight")) return mat(e->patch_w, D, e->kpatch, e->kpatch_pad);
    if (!strcmp(name, "patch_embed.proj.bias") && c.patch_bias) return vec(e->patch_b, D);
    if (!strcmp(name, "pre_norm.weight") && c.pre_ln) return vec(e->pre_w, D);
    if (!strcmp(name, "pre_norm.bias") && c.pre_ln) return vec(e->pre_b, D);
    if (!strcmp(name, "norm.weight")) return vec(e->norm_w, D);
    if (!strcmp(name, "norm.bias")) return vec(e->norm_b, D);
    if (!strcmp(name, "head.proj.weight") && c.proj_dim > 0) return mat(e->head_w, c.proj_dim, D, D);
    int idx = -1, consumed = 0;
    if (sscanf(name, "blocks.%d.%n", &idx, &consumed) == 1 && consumed > 0 && idx >= 0 && idx < c.depth) {
        const char* r = name + consumed;
        Layer& ly = *e->layers[idx];
        if (!strcmp(r, "norm1.weight")) return vec(ly.ln1_w, D);
        if (!strcmp(r, "norm1.bias")) return vec(ly.ln1_b, D);
        if (!strcmp(r, "attn.qkv.weight")) return mat(ly.qkv_w, 3 * D, D, D);
        if (!strcmp(r, "attn.qkv.bias")) return vec(ly.qkv_b, 3 * D);
        if (!strcmp(r, "attn.proj.weight")) return mat(ly.proj_w, D, D, D);
        if (!strcmp(r, "attn.proj.bias")) return vec(ly.proj_b, D);
        if (!strcmp(r, "norm2.weight")) return vec(ly.ln2_w, D);
        if (!strcmp(r, "norm2.bias")) return vec(ly.ln2_b, D);
        if (!strcmp(r, "mlp.fc1.weight")) return mat(ly.fc1_w, H, D, D);
        if (!strcmp(r, "mlp.fc1.bias")) return vec(ly.fc1_b, H);
        if (!strcmp(r, "mlp.fc2.weight")) return mat(ly.fc2_w, D, H, H);
        if (!strcmp(r, "mlp.fc2.bias")) return vec(ly.fc2_b, D);
    }
    return false;
}

struct NamedBuf {
    std::string name;
    const DevBuf* buf;
};

void list_params(const vidil_encoder* e, std::vector<NamedBuf>& out) {
    const vidil_encoder_cfg& c = e->cfg;
    out.push_back({"cls_token", &e->cls});
    out.push_back({"pos_embed", &e->pos});
    out.push_back({"patch_embed.proj.weight", &e->patch_w});
    if (c.patch_bias) out.push_back({"patch_embed.proj.bias", &e->patch_b});
    if (c.pre_ln) {
        out.push_back({"pre_norm.weight", &e->pre_w});
        out.push_back({"pre_norm.bias", &e->pre_b});
    }
    for (int i = 0; i < c.depth; ++i) {
        const Layer& ly = *e->layers[i];
        const std::string p = "blocks." + std::to_string(i) + ".";
        out.push_back({p + "norm1.weight", &ly.ln1_w});
        out.push_back({p + "norm1.bias", &ly.ln1_b});
        out.push_back({p + "attn.qkv.weight", &ly.qkv_w});
        out.push_back({p + "attn.qkv.bias", &ly.qkv_b});
        out.push_back({p + "attn.proj.weight", &ly.proj_w});
        out.push_back({p + "attn.proj.bias", &ly.proj_b});
        out.push_back({p + "norm2.weight", &ly.ln2_w});
        out.push_back({p + "norm2.bias", &ly.ln2_b});
        out.push_back({p + "mlp.fc1.weight", &ly.fc1_w});
        out.push_back({p + "mlp.fc1.bias", &ly.fc1_b});
        out.push_back({p + "mlp.fc2.weight", &ly.fc2_w});
        out.push_back({p + "mlp.fc2.bias", &ly.fc2_b});
    }
    out.push_back({"norm.weight", &e->norm_w});
    out.push_back({"norm.bias", &e->norm_b});
    if (c.proj_dim > 0) out.push_back({"head.proj.weight", &e->head_w});
}

}  // namespace

extern "C" {

int32_t vidil_abi_version(void) { return VIDIL_B200_ABI_VERSION; }
const char* vidil_last_error(void) { return get_error(); }
int64_t vidil_kernel_launch_count(void) { return launch_count(); }

int32_t vidil_encoder_create(const vidil_encoder_cfg* cfg, vidil_encoder** out) {
    if (cfg == nullptr || out == nullptr) {
        set_error("vidil_encoder_create: null argument");
        return 1;
    }
    *out = nullptr;
    vidil_encoder_cfg c = *cfg;
    if (c.cta_group == 0) c.cta_group = 2;
    if (c.img_size <= 0 || c.patch_size <= 0 || c.patch_size % 2 != 0 || c.img_size % c.patch_size != 0) {
        set_error("unsupported geometry: img_size=%d patch_size=%d (patch must be even and divide the image)", c.img_size,
                  c.patch_size);
        return 1;
    }
    if (c.embed_dim <= 0 || c.embed_dim % 128 != 0 || c.num_heads * 64 != c.embed_dim) {
        set_error("unsupported width: embed_dim=%d num_heads=%d (need embed_dim %% 128 == 0 and head_dim 64)", c.embed_dim,
                  c.num_heads);
        return 1;
    }
    if (c.embed_dim != 128 && c.embed_dim != 256 && c.embed_dim != 512 && c.embed_dim != 768 && c.embed_dim != 1024 &&
        c.embed_dim != 1280) {
        set_error("unsupported embed_dim=%d (LayerNorm kernel covers 128/256/512/768/1024/1280)", c.embed_dim);
        return 1;
    }
    if (c.depth <= 0 || c.mlp_dim <= 0 || c.mlp_dim % 64 != 0) {
        set_error("unsupported depth=%d / mlp_dim=%d (mlp_dim must be a multiple of 64)", c.depth, c.mlp_dim);
        return 1;
    }
    if (c.dtype != VIDIL_DTYPE_BF16 && c.dtype != VIDIL_DTYPE_FP16) {
        set_error("unsupported dtype %d", c.dtype);
        return 1;
    }
    if (c.act != VIDIL_ACT_GELU_ERF && c.act != VIDIL_ACT_QUICK_GELU) {
        set_error("unsupported activation %d", c.act);
        return 1;
    }
    if (c.cta_group != 1 && c.cta_group != 2) {
        set_error("cta_group must be 0, 1 or 2");
        return 1;
    }
    if (c.proj_dim < 0 || c.proj_dim % 4 != 0) {
        set_error("proj_dim must be a non-negative multiple of 4");
        return 1;
    }
    if (gemm_num_sms() == 0) {
        if (get_error()[0] == 0) set_error("no sm_100 CUDA device available");
        return 1;
    }
    std::unique_ptr<vidil_encoder> e(new vidil_encoder());
    e->cfg = c;
    VIDIL_CUDA_OK(cudaGetDevice(&e->device));
    e->dt = (c.dtype == VIDIL_DTYPE_BF16) ? DT_BF16 : DT_FP16;
    e->grid = c.img_size / c.patch_size;
    e->patches = e->grid * e->grid;
    e->tokens = e->patches + 1;
    e->kpatch = 3 * c.patch_size * c.patch_size;
    e->kpatch_pad = round_up(e->kpatch, 64);
    for (int i = 0; i < c.depth; ++i) e->layers.emplace_back(new Layer());

    // allocate every parameter buffer up front (fp32 vectors, 16-bit matrices)
    std::vector<NamedBuf> names;
    list_params(e.get(), names);
    for (auto& nb : names) {
        Slot s;
        if (!find_slot(e.get(), nb.name.c_str(), s)) {
            set_error("internal: no slot for %s", nb.name.c_str());
            return 1;
        }
        const size_t bytes = s.matrix ? static_cast<size_t>(s.rows) * s.ld * 2 : static_cast<size_t>(s.numel) * 4;
        if (s.buf->alloc(bytes)) return 1;
    }
    *out = e.release();
    return 0;
}

void vidil_encoder_destroy(vidil_encoder* enc) { delete enc; }

int32_t vidil_encoder_load(vidil_encoder* enc, const char* name, const float* dev_ptr, int64_t numel, void* stream) {
    if (enc == nullptr || name == nullptr || dev_ptr == nullptr) {
        set_error("vidil_encoder_load: null argument");
        return 1;
    }
    Slot s;
    if (!find_slot(enc, name, s)) {
        set_error("vidil_encoder_load: unknown parameter '%s' for this configuration", name);
        return 1;
    }
    if (numel != s.numel) {
        set_error("vidil_encoder_load: '%s' has %lld elements, expected %lld", name, (long long)numel, (long long)s.numel);
        return 1;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (s.matrix) {
        if (cast_run(dev_ptr, s.buf->p, enc->dt, s.rows, s.cols, s.ld, st)) return 1;
    } else {
        VIDIL_CUDA_OK(cudaMemcpyAsync(s.buf->p, dev_ptr, static_cast<size_t>(numel) * 4, cudaMemcpyDeviceToDevice, st));
    }
    s.buf->loaded = true;
    return 0;
}

int32_t vidil_encoder_check_loaded(const vidil_encoder* enc) {
    if (enc == nullptr) {
        set_error("null encoder");
        return 1;
    }
    std::vector<NamedBuf> names;
    list_params(enc, names);
    for (auto& nb : names)
        if (!nb.buf->loaded) {
            set_error("parameter '%s' has not been loaded", nb.name.c_str());
            return 1;
        }
    return 0;
}

size_t vidil_encoder_workspace_bytes(const vidil_encoder* enc, int32_t batch) {
    if (enc == nullptr || batch <= 0) return 0;
    return ws_layout(enc, batch).total;
}

int32_t vidil_encoder_tokens(const vidil_encoder* enc) { return enc ? enc->tokens : 0; }

static int vit_forward_any(vidil_encoder* enc, const float* frames, int32_t batch, void* out_tokens, bool out_f32, void* workspace,
                           size_t workspace_bytes, void* stream) {
    if (check_common(enc, frames, batch, out_tokens, workspace, workspace_bytes)) return 1;
    if (enc->cfg.proj_dim > 0 || enc->cfg.pre_ln) {
        set_error("vidil_vit_forward: this handle is a CLIP-style tower; use vidil_clip_forward");
        return 1;
    }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    Plan* pl = get_plan(enc, batch, workspace);
    if (pl == nullptr) return 1;
    if (run_trunk(enc, *pl, frames, s)) return 1;
    return ln_timed(enc, s, pl->resid, enc->cfg.embed_dim, enc->norm_w, enc->norm_b, out_tokens, out_f32, batch * enc->tokens);
}

int32_t vidil_vit_forward(vidil_encoder* enc, const float* frames, int32_t batch, float* out_tokens, void* workspace,
                          size_t workspace_bytes, void* stream) {
    return vit_forward_any(enc, frames, batch, out_tokens, true, workspace, workspace_bytes, stream);
}

int32_t vidil_vit_forward16(vidil_encoder* enc, const float* frames, int32_t batch, void* out_tokens16, void* workspace,
                            size_t workspace_bytes, void* stream) {
    return vit_forward_any(enc, frames, batch, out_tokens16, false, workspace, workspace_bytes, stream);
}

int32_t vidil_clip_forward(vidil_encoder* enc, const float* frames, int32_t batch, float* out_embeds, float* out_hidden,
                           void* workspace, size_t workspace_bytes, void* stream) {
    if (check_common(enc, frames, batch, out_embeds, workspace, workspace_bytes)) return 1;
    if (enc->cfg.proj_dim <= 0) {
        set_error("vidil_clip_forward: handle has no projection head (proj_dim == 0)");
        return 1;
    }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    Plan* pl = get_plan(enc, batch, workspace);
    if (pl == nullptr) return 1;
    if (run_trunk(enc, *pl, frames, s)) return 1;
    const int D = enc->cfg.embed_dim;
    const size_t M = static_cast<size_t>(batch) * enc->tokens;
    if (out_hidden != nullptr)
        VIDIL_CUDA_OK(cudaMemcpyAsync(out_hidden, pl->resid, M * D * 4, cudaMemcpyDeviceToDevice, s));
    // post_layernorm on the CLS row of every frame only, then the bias-free projection and L2 normalisation
    if (ln_timed(enc, s, pl->resid, static_cast<int64_t>(enc->tokens) * D, enc->norm_w, enc->norm_b, pl->cls_ln, false, batch))
        return 1;
    if (gemm_timed(enc, s, pl->head)) return 1;
    return timed(enc, s, VIDIL_KCLASS_OTHER, 0.0, 8.0 * batch * enc->cfg.proj_dim,
                 [&] { return l2norm_run(pl->head_out, out_embeds, batch, enc->cfg.proj_dim, s); });
}

int32_t vidil_encoder_set_profiling(vidil_encoder* enc, int32_t enable) {
    if (enc == nullptr) {
        set_error("null encoder");
        return 1;
    }
    enc->profiling = enable != 0;
    return 0;
}

int32_t vidil_encoder_read_profile(vidil_encoder* enc, vidil_kernel_stats* out) {
    if (enc == nullptr || out == nullptr) {
        set_error("vidil_encoder_read_profile: null argument");
        return 1;
    }
    memset(out, 0, sizeof(*out));
    for (auto& r : enc->prof) {
        VIDIL_CUDA_OK(cudaEventSynchronize(r.b));
        float ms = 0.f;
        VIDIL_CUDA_OK(cudaEventElapsedTime(&ms, r.a, r.b));
        out->ms[r.cls] += ms;
        out->flops[r.cls] += r.flops;
        out->bytes[r.cls] += r.bytes;
        out->launches[r.cls] += 1;
        enc->free_events.push_back(r.a);
        enc->free_events.push_back(r.b);
    }
    enc->prof.clear();
    return 0;
}

size_t vidil_encoder_host_scratch_bytes(const vidil_encoder* enc, int32_t batch) {
    if (enc == nullptr || batch <= 0) return 0;
    const size_t in = align_up(static_cast<size_t>(batch) * 3 * enc->cfg.img_size * enc->cfg.img_size * 4);
    const size_t out_tok = align_up(static_cast<size_t>(batch) * enc->tokens * enc->cfg.embed_dim * 4);
    return in + out_tok + ws_layout(enc, batch).total;
}

static int host_forward(vidil_encoder* enc, const float* frames_host, int32_t batch, float* out_host, bool clip,
                        void* dev_scratch, size_t dev_scratch_bytes, void* stream) {
    if (enc == nullptr || frames_host == nullptr || out_host == nullptr || dev_scratch == nullptr || batch <= 0) {
        set_error("host forward: null argument or empty batch");
        return 1;
    }
    const size_t need = vidil_encoder_host_scratch_bytes(enc, batch);
    if (dev_scratch_bytes < need) {
        set_error("device scratch too small: %zu bytes given, %zu needed", dev_scratch_bytes, need);
        return 1;
    }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const size_t in_bytes = static_cast<size_t>(batch) * 3 * enc->cfg.img_size * enc->cfg.img_size * 4;
    const size_t out_tok_bytes = static_cast<size_t>(batch) * enc->tokens * enc->cfg.embed_dim * 4;
    uint8_t* base = reinterpret_cast<uint8_t*>(dev_scratch);
    float* d_in = reinterpret_cast<float*>(base);
    float* d_out = reinterpret_cast<float*>(base + align_up(in_bytes));
    void* ws = base + align_up(in_bytes) + align_up(out_tok_bytes);
    const size_t ws_bytes = dev_scratch_bytes - align_up(in_bytes) - align_up(out_tok_bytes);
    VIDIL_CUDA_OK(cudaMemcpyAsync(d_in, frames_host, in_bytes, cudaMemcpyHostToDevice, s));
    size_t out_bytes;
    if (clip) {
        if (vidil_clip_forward(enc, d_in, batch, d_out, nullptr, ws, ws_bytes, stream)) return 1;
        out_bytes = static_cast<size_t>(batch) * enc->cfg.proj_dim * 4;
    } else {
        if (vidil_vit_forward(enc, d_in, batch, d_out, ws, ws_bytes, stream)) return 1;
        out_bytes = out_tok_bytes;
    }
    VIDIL_CUDA_OK(cudaMemcpyAsync(out_host, d_out, out_bytes, cudaMemcpyDeviceToHost, s));
    VIDIL_CUDA_OK(cudaStreamSynchronize(s));
    return 0;
}

int32_t vidil_vit_forward_host(vidil_encoder* enc, const float* frames_host, int32_t batch, float* out_tokens_host,
                               void* dev_scratch, size_t dev_scratch_bytes, void* stream) {
    return host_forward(enc, frames_host, batch, out_tokens_host, false, dev_scratch, dev_scratch_bytes, stream);
}
int32_t vidil_clip_forward_host(vidil_encoder* enc, const float* frames_host, int32_t batch, float* out_embeds_host,
                                void* dev_scratch, size_t dev_scratch_bytes, void* stream) {
    return host_forward(enc, frames_host, batch, out_embeds_host, true, dev_scratch, dev_scratch_bytes, stream);
}

// ---- pipelined host-buffer encoding --------------------------------------------------------------------
size_t vidil_encoder_host_pipeline_scratch_bytes(const vidil_encoder* enc, int32_t batch) {
    if (enc == nullptr || batch <= 0) return 0;
    const size_t in = align_up(static_cast<size_t>(batch) * 3 * enc->cfg.img_size * enc->cfg.img_size * 4);
    const size_t out_tok = align_up(static_cast<size_t>(batch) * enc->tokens * enc->cfg.embed_dim * 4);
    return 2 * (in + out_tok) + ws_layout(enc, batch).total;
}

static int host_submit_any(vidil_encoder* enc, const float* frames_host, int32_t batch, void* out_host, bool out16, int32_t slot,
                           void* dev_scratch, size_t dev_scratch_bytes, void* stream) {
    if (enc == nullptr || frames_host == nullptr || out_host == nullptr || dev_scratch == nullptr || batch <= 0) {
        set_error("vidil_encoder_host_submit: null argument or empty batch");
        return 1;
    }
    if (slot < 0 || slot > 1) {
        set_error("vidil_encoder_host_submit: slot must be 0 or 1");
        return 1;
    }
    if (dev_scratch_bytes < vidil_encoder_host_pipeline_scratch_bytes(enc, batch)) {
        set_error("device scratch too small: %zu bytes given, %zu needed", dev_scratch_bytes,
                  vidil_encoder_host_pipeline_scratch_bytes(enc, batch));
        return 1;
    }
    if (enc->copy_in == nullptr) {
        VIDIL_CUDA_OK(cudaStreamCreateWithFlags(&enc->copy_in, cudaStreamNonBlocking));
        VIDIL_CUDA_OK(cudaStreamCreateWithFlags(&enc->copy_out, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            VIDIL_CUDA_OK(cudaEventCreateWithFlags(&enc->ev_in[i], cudaEventDisableTiming));
            VIDIL_CUDA_OK(cudaEventCreateWithFlags(&enc->ev_compute[i], cudaEventDisableTiming));
            VIDIL_CUDA_OK(cudaEventCreateWithFlags(&enc->ev_out[i], cudaEventDisableTiming));
        }
        VIDIL_CUDA_OK(cudaEventCreateWithFlags(&enc->ev_bind, cudaEventDisableTiming));
    }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (dev_scratch != enc->pipe_scratch || dev_scratch_bytes != enc->pipe_bytes || batch > enc->pipe_cap) {
        // (Re)binding the pipeline to a scratch buffer or to a larger capacity moves the slot offsets: drain everything
        // in flight first (rare: once per stream of batches), then order the copy engines after whatever the caller's
        // stream still has enqueued on this memory.
        for (int i = 0; i < 2; ++i) VIDIL_CUDA_OK(cudaEventSynchronize(enc->ev_out[i]));
        VIDIL_CUDA_OK(cudaStreamSynchronize(enc->copy_in));
        VIDIL_CUDA_OK(cudaStreamSynchronize(enc->copy_out));
        VIDIL_CUDA_OK(cudaEventRecord(enc->ev_bind, s));
        VIDIL_CUDA_OK(cudaStreamWaitEvent(enc->copy_in, enc->ev_bind, 0));
        VIDIL_CUDA_OK(cudaStreamWaitEvent(enc->copy_out, enc->ev_bind, 0));
        enc->pipe_scratch = dev_scratch;
        enc->pipe_bytes = dev_scratch_bytes;
        enc->pipe_cap = batch;
    }
    const bool clip = enc->cfg.proj_dim > 0;
    const int cap = enc->pipe_cap;  // batch <= cap: a ragged last batch uses the same slot offsets as the full ones
    const size_t in_bytes = static_cast<size_t>(batch) * 3 * enc->cfg.img_size * enc->cfg.img_size * 4;
    const size_t out_tok_bytes = static_cast<size_t>(batch) * enc->tokens * enc->cfg.embed_dim * 4;
    const size_t cap_in = align_up(static_cast<size_t>(cap) * 3 * enc->cfg.img_size * enc->cfg.img_size * 4);
    const size_t cap_out = align_up(static_cast<size_t>(cap) * enc->tokens * enc->cfg.embed_dim * 4);
    const size_t per_slot = cap_in + cap_out;
    uint8_t* base = reinterpret_cast<uint8_t*>(dev_scratch);
    float* d_in = reinterpret_cast<float*>(base + slot * per_slot);
    float* d_out = reinterpret_cast<float*>(base + slot * per_slot + cap_in);
    void* ws = base + 2 * per_slot;
    const size_t ws_bytes = dev_scratch_bytes - 2 * per_slot;
    // H2D: may start as soon as the previous forward that read this slot's frames is done
    VIDIL_CUDA_OK(cudaStreamWaitEvent(enc->copy_in, enc->ev_compute[slot], 0));
    VIDIL_CUDA_OK(cudaMemcpyAsync(d_in, frames_host, in_bytes, cudaMemcpyHostToDevice, enc->copy_in));
    VIDIL_CUDA_OK(cudaEventRecord(enc->ev_in[slot], enc->copy_in));
    // forward: needs the frames, and the previous D2H out of this slot's result buffer
    VIDIL_CUDA_OK(cudaStreamWaitEvent(s, enc->ev_in[slot], 0));
    VIDIL_CUDA_OK(cudaStreamWaitEvent(s, enc->ev_out[slot], 0));
    size_t out_bytes;
    if (clip) {
        if (vidil_clip_forward(enc, d_in, batch, d_out, nullptr, ws, ws_bytes, stream)) return 1;
        out_bytes = static_cast<size_t>(batch) * enc->cfg.proj_dim * 4;
    } else {
        if (vit_forward_any(enc, d_in, batch, d_out, !out16, ws, ws_bytes, stream)) return 1;
        out_bytes = out16 ? out_tok_bytes / 2 : out_tok_bytes;
    }
    VIDIL_CUDA_OK(cudaEventRecord(enc->ev_compute[slot], s));
    // D2H on the other copy engine
    VIDIL_CUDA_OK(cudaStreamWaitEvent(enc->copy_out, enc->ev_compute[slot], 0));
    VIDIL_CUDA_OK(cudaMemcpyAsync(out_host, d_out, out_bytes, cudaMemcpyDeviceToHost, enc->copy_out));
    VIDIL_CUDA_OK(cudaEventRecord(enc->ev_out[slot], enc->copy_out));
    return 0;
}

int32_t vidil_encoder_host_submit(vidil_encoder* enc, const float* frames_host, int32_t batch, float* out_host, int32_t slot,
                                  void* dev_scratch, size_t dev_scratch_bytes, void* stream) {
    return host_submit_any(enc, frames_host, batch, out_host, false, slot, dev_scratch, dev_scratch_bytes, stream);
}

int32_t vidil_encoder_host_submit16(vidil_encoder* enc, const float* frames_host, int32_t batch, void* out_host16, int32_t slot,
                                    void* dev_scratch, size_t dev_scratch_bytes, void* stream) {
    if (enc != nullptr && enc->cfg.proj_dim > 0) {
        set_error("vidil_encoder_host_submit16: 16-bit output is for token outputs (ViT handles); CLIP embeds stay fp32");
        return 1;
    }
    return host_submit_any(enc, frames_host, batch, out_host16, true, slot, dev_scratch, dev_scratch_bytes, stream);
}

int32_t vidil_encoder_host_wait(vidil_encoder* enc, int32_t slot) {
    if (enc == nullptr || slot < 0 || slot > 1) {
        set_error("vidil_encoder_host_wait: bad argument");
        return 1;
    }
    if (enc->ev_out[slot] == nullptr) return 0;  // nothing was ever submitted
    VIDIL_CUDA_OK(cudaEventSynchronize(enc->ev_out[slot]));
    return 0;
}

// ---- similarity + top-k ---------------------------------------------------------------------------
// image_embeds @ text_embeds.t() -> per-frame top-k (run_visual_tokenization.py:276,306) without ever writing the [F, T]
// score matrix: the fp16 tcgen05 GEMM's epilogue keeps the four best scores of every 32-phrase group (EPI_TOP4, 16 bytes per
// group instead of 128), topk_select_kernel re-scores candidates in fp32 until the fp32 ranking is certain.
}  // extern "C"

struct vidil_sim_bank {
    int T = 0, D = 0;
    DevBuf bank16, bank32, max_norm;
    // the GEMM of the last call (TMA maps encoded): a run scores batch after batch of the same size in the same workspace
    mutable GemmProblem plan;
    mutable int plan_F = 0;
    mutable const void* plan_ws = nullptr;
};

namespace {

inline int sim_groups(int T) { return (T + 31) / 32; }
inline int sim_ld(int T) { return 4 * sim_groups(T); }  // floats per row of the EPI_TOP4 output

struct SimWs {
    size_t img16, top2, total;
};
SimWs sim_ws(int F, int T, int D) {
    SimWs w;
    w.img16 = 0;
    w.top2 = align_up(static_cast<size_t>(F) * D * 2);
    w.total = w.top2 + align_up(static_cast<size_t>(F) * sim_ld(T) * sizeof(float));
    return w;
}

int sim_topk_core(const float* img, const void* bank16, const float* bank32, const float* max_norm_dev, int F, int T, int D, int k,
                  float* out_scores, int32_t* out_idx, uint8_t* ws, cudaStream_t s, const vidil_sim_bank* cache = nullptr) {
    const SimWs L = sim_ws(F, T, D);
    void* img_h = ws + L.img16;
    float* top2 = reinterpret_cast<float*>(ws + L.top2);
    const int G = sim_groups(T);
    // fp16 operands: unit-norm embeddings sit well inside fp16 range and keep 3 more mantissa bits than bf16
    if (cast_run(img, img_h, DT_FP16, F, D, D, s)) return 1;
    GemmProblem local;
    GemmProblem& g = cache ? cache->plan : local;
    if (cache == nullptr || cache->plan_F != F || cache->plan_ws != ws) {
        g = GemmProblem();
        g.dt = DT_FP16;
        g.epi = EPI_TOP4;
        g.cta_group = 2;
        g.M = F; g.N = T; g.K = D;
        g.A = img_h; g.lda = D;
        g.W = bank16; g.ldw = D;
        g.out = top2; g.ldo = sim_ld(T);
        if (gemm_prepare(g)) return 1;
        if (cache) {
            cache->plan_F = F;
            cache->plan_ws = ws;
        }
    }
    if (gemm_run(g, s)) return 1;
    return topk_select_run(top2, sim_ld(T), G, img, bank32, max_norm_dev, F, T, D, k, out_scores, out_idx, s);
}

int sim_check(const char* who, int F, int T, int D) {
    if (F <= 0 || T <= 0 || D <= 0 || D % 64 != 0) {
        set_error("%s: need F, T > 0 and D a positive multiple of 64 (F=%d T=%d D=%d)", who, F, T, D);
        return 1;
    }
    if (gemm_num_sms() == 0) {
        if (get_error()[0] == 0) set_error("no sm_100 CUDA device available");
        return 1;
    }
    return 0;
}

}  // namespace

extern "C" {

int32_t vidil_sim_bank_create(const float* bank, int32_t T, int32_t D, void* stream, vidil_sim_bank** out) {
    if (bank == nullptr || out == nullptr) {
        set_error("vidil_sim_bank_create: null argument");
        return 1;
    }
    *out = nullptr;
    if (sim_check("vidil_sim_bank_create", 1, T, D)) return 1;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    std::unique_ptr<vidil_sim_bank> b(new vidil_sim_bank());
    b->T = T;
    b->D = D;
    const size_t n = static_cast<size_t>(T) * D;
    if (b->bank16.alloc(n * 2) || b->bank32.alloc(n * 4) || b->max_norm.alloc(sizeof(float))) return 1;
    VIDIL_CUDA_OK(cudaMemcpyAsync(b->bank32.p, bank, n * 4, cudaMemcpyDeviceToDevice, s));
    if (cast_run(bank, b->bank16.p, DT_FP16, T, D, D, s)) return 1;
    if (max_row_norm_run(bank, T, D, reinterpret_cast<float*>(b->max_norm.p), s)) return 1;
    *out = b.release();
    return 0;
}

void vidil_sim_bank_destroy(vidil_sim_bank* bank) { delete bank; }

size_t vidil_sim_bank_topk_workspace_bytes(const vidil_sim_bank* bank, int32_t F) {
    if (bank == nullptr || F <= 0) return 0;
    return sim_ws(F, bank->T, bank->D).total;
}

int32_t vidil_sim_bank_topk(const vidil_sim_bank* bank, const float* img, int32_t F, int32_t k, float* out_scores,
                            int32_t* out_idx, void* workspace, size_t workspace_bytes, void* stream) {
    if (bank == nullptr || img == nullptr || out_scores == nullptr || out_idx == nullptr || workspace == nullptr) {
        set_error("vidil_sim_bank_topk: null argument");
        return 1;
    }
    if (sim_check("vidil_sim_bank_topk", F, bank->T, bank->D)) return 1;
    if (workspace_bytes < sim_ws(F, bank->T, bank->D).total || (reinterpret_cast<uintptr_t>(workspace) & (ALIGN - 1))) {
        set_error("vidil_sim_bank_topk: workspace too small or not %zu-byte aligned", ALIGN);
        return 1;
    }
    return sim_topk_core(img, bank->bank16.p, reinterpret_cast<const float*>(bank->bank32.p),
                         reinterpret_cast<const float*>(bank->max_norm.p), F, bank->T, bank->D, k, out_scores, out_idx,
                         reinterpret_cast<uint8_t*>(workspace), reinterpret_cast<cudaStream_t>(stream), bank);
}

// One-shot form: the bank is converted inside the call (workspace), nothing is cached.
size_t vidil_sim_topk_workspace_bytes(int32_t F, int32_t T, int32_t D) {
    if (F <= 0 || T <= 0 || D <= 0) return 0;
    return sim_ws(F, T, D).total + align_up(static_cast<size_t>(T) * D * 2) + ALIGN;
}

int32_t vidil_sim_topk(const float* img, const float* bank, int32_t F, int32_t T, int32_t D, int32_t k, float* out_scores,
                       int32_t* out_idx, void* workspace, size_t workspace_bytes, void* stream) {
    if (img == nullptr || bank == nullptr || out_scores == nullptr || out_idx == nullptr || workspace == nullptr) {
        set_error("vidil_sim_topk: null argument");
        return 1;
    }
    if (sim_check("vidil_sim_topk", F, T, D)) return 1;
    if (workspace_bytes < vidil_sim_topk_workspace_bytes(F, T, D) || (reinterpret_cast<uintptr_t>(workspace) & (ALIGN - 1))) {
        set_error("vidil_sim_topk: workspace too small or not %zu-byte aligned", ALIGN);
        return 1;
    }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    uint8_t* base = reinterpret_cast<uint8_t*>(workspace);
    const size_t core = sim_ws(F, T, D).total;
    void* bank_h = base + core;
    float* max_norm = reinterpret_cast<float*>(base + core + align_up(static_cast<size_t>(T) * D * 2));
    if (cast_run(bank, bank_h, DT_FP16, T, D, D, s)) return 1;
    if (max_row_norm_run(bank, T, D, max_norm, s)) return 1;
    return sim_topk_core(img, bank_h, bank, max_norm, F, T, D, k, out_scores, out_idx, base, s);
}

// ---- frame pre-processing -------------------------------------------------------------------------------
size_t vidil_preprocess_workspace_bytes(int32_t batch, int32_t in_h, int32_t in_w, int32_t out_size) {
    return preprocess_workspace_bytes(batch, in_h, in_w, out_size);
}

int32_t vidil_preprocess_frames(const uint8_t* frames_u8, int32_t batch, int32_t in_h, int32_t in_w, int32_t out_size,
                                const float* mean3, const float* std3, float* out, void* workspace, size_t workspace_bytes,
                                void* stream) {
    if (frames_u8 == nullptr || mean3 == nullptr || std3 == nullptr || out == nullptr || workspace == nullptr) {
        set_error("vidil_preprocess_frames: null argument");
        return 1;
    }
    if (gemm_num_sms() == 0) {
        if (get_error()[0] == 0) set_error("no sm_100 CUDA device available");
        return 1;
    }
    return preprocess_run(frames_u8, batch, in_h, in_w, out_size, mean3, std3, out, workspace, workspace_bytes,
                          reinterpret_cast<cudaStream_t>(stream));
}

size_t vidil_clip_preprocess_workspace_bytes(int32_t batch, int32_t in_h, int32_t in_w, int32_t out_size) {
    return clip_preprocess_workspace_bytes(batch, in_h, in_w, out_size);
}

int32_t vidil_clip_preprocess_frames(const uint8_t* frames_u8, int32_t batch, int32_t in_h, int32_t in_w, int32_t out_size,
                                     const float* mean3, const float* std3, float* out, void* workspace, size_t workspace_bytes,
                                     void* stream) {
    if (frames_u8 == nullptr || mean3 == nullptr || std3 == nullptr || out == nullptr || workspace == nullptr) {
        set_error("vidil_clip_preprocess_frames: null argument");
        return 1;
    }
    if (gemm_num_sms() == 0) {
        if (get_error()[0] == 0) set_error("no sm_100 CUDA device available");
        return 1;
    }
    return clip_preprocess_run(frames_u8, batch, in_h, in_w, out_size, mean3, std3, out, workspace, workspace_bytes,
                               reinterpret_cast<cudaStream_t>(stream));
}

void vidil_debug_set_attention_trace(void* dev_buf) { attention_set_trace(reinterpret_cast<long long*>(dev_buf)); }

// ---- operator-level entry points --------------------------------------------------------------------
size_t vidil_op_linear_workspace_bytes(int32_t M, int32_t N, int32_t K) {
    if (M <= 0 || N <= 0 || K <= 0) return 0;
    return align_up(static_cast<size_t>(M) * K * 2) + align_up(static_cast<size_t>(N) * K * 2) +
           align_up(static_cast<size_t>(M) * N * 2);
}

int32_t vidil_op_linear(const float* A, const float* W, const float* bias, float* out, int32_t M, int32_t N, int32_t K,
                        int32_t epi, int32_t dtype, int32_t cta_group, const float* pos, int32_t patches_per_frame,
                        void* workspace, size_t workspace_bytes, void* stream) {
    if (A == nullptr || W == nullptr || out == nullptr || workspace == nullptr) {
        set_error("vidil_op_linear: null argument");
        return 1;
    }
    if (workspace_bytes < vidil_op_linear_workspace_bytes(M, N, K) || (reinterpret_cast<uintptr_t>(workspace) & (ALIGN - 1))) {
        set_error("vidil_op_linear: workspace too small or misaligned");
        return 1;
    }
    if (epi < 0 || epi >= EPI_COUNT || (dtype != VIDIL_DTYPE_BF16 && dtype != VIDIL_DTYPE_FP16)) {
        set_error("vidil_op_linear: bad epilogue %d or dtype %d", epi, dtype);
        return 1;
    }
    if (gemm_num_sms() == 0) {
        if (get_error()[0] == 0) set_error("no sm_100 CUDA device available");
        return 1;
    }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const DType dt = (dtype == VIDIL_DTYPE_BF16) ? DT_BF16 : DT_FP16;
    uint8_t* base = reinterpret_cast<uint8_t*>(workspace);
    void* a_h = base;
    void* w_h = base + align_up(static_cast<size_t>(M) * K * 2);
    void* o_h = base + align_up(static_cast<size_t>(M) * K * 2) + align_up(static_cast<size_t>(N) * K * 2);
    if (cast_run(A, a_h, dt, M, K, K, s)) return 1;
    if (cast_run(W, w_h, dt, N, K, K, s)) return 1;
    const bool f32_out = (epi == EPI_RESID || epi == EPI_PATCH || epi == EPI_STORE_F32);
    GemmProblem g;
    g.dt = dt;
    g.epi = epi;
    g.cta_group = cta_group == 0 ? 2 : cta_group;
    g.M = M; g.N = N; g.K = K;
    g.A = a_h; g.lda = K;
    g.W = w_h; g.ldw = K;
    g.bias = bias;
    g.out = f32_out ? static_cast<void*>(out) : o_h;
    g.ldo = N;
    g.pos = pos;
    g.patches_per_frame = patches_per_frame;
    if (gemm_prepare(g)) return 1;
    if (gemm_run(g, s)) return 1;
    if (!f32_out) return uncast_run(o_h, out, dt, static_cast<int64_t>(M) * N, s);
    return 0;
}

int32_t vidil_op_layernorm(const float* in, const float* gamma, const float* beta, float* out, int32_t rows, int32_t D,
                           float eps, void* stream) {
    if (in == nullptr || gamma == nullptr || beta == nullptr || out == nullptr) {
        set_error("vidil_op_layernorm: null argument");
        return 1;
    }
    if (gemm_num_sms() == 0) {
        if (get_error()[0] == 0) set_error("no sm_100 CUDA device available");
        return 1;
    }
    return layernorm_run(in, D, gamma, beta, out, true, DT_BF16, rows, D, eps, reinterpret_cast<cudaStream_t>(stream));
}

size_t vidil_op_attention_workspace_bytes(int32_t B, int32_t N, int32_t H) {
    if (B <= 0 || N <= 0 || H <= 0) return 0;
    const size_t rows = static_cast<size_t>(B) * N;
    return align_up(rows * 3 * H * 64 * 2) + align_up(rows * H * 64 * 2);
}

static int op_attention_any(const float* qkv, float* out, int32_t B, int32_t N, int32_t H, float scale, int32_t dtype, bool causal,
                            void* workspace, size_t workspace_bytes, void* stream) {
    if (qkv == nullptr || out == nullptr || workspace == nullptr) {
        set_error("vidil_op_attention: null argument");
        return 1;
    }
    if (workspace_bytes < vidil_op_attention_workspace_bytes(B, N, H) ||
        (reinterpret_cast<uintptr_t>(workspace) & (ALIGN - 1))) {
        set_error("vidil_op_attention: workspace too small or misaligned");
        return 1;
    }
    if (gemm_num_sms() == 0) {
        if (get_error()[0] == 0) set_error("no sm_100 CUDA device available");
        return 1;
    }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const DType dt = (dtype == VIDIL_DTYPE_BF16) ? DT_BF16 : DT_FP16;
    const int64_t rows = static_cast<int64_t>(B) * N;
    uint8_t* base = reinterpret_cast<uint8_t*>(workspace);
    void* qkv_h = base;
    void* out_h = base + align_up(static_cast<size_t>(rows) * 3 * H * 64 * 2);
    if (cast_run(qkv, qkv_h, dt, rows, 3 * H * 64, 3 * H * 64, s)) return 1;
    if (causal && !attention_tc_supported(N)) {
        set_error("vidil_op_attention_causal: the causal mask is implemented for sequences of at most 208 tokens (N=%d)", N);
        return 1;
    }
    if (attention_tc_supported(N)) {
        AttentionMaps maps;
        if (attention_tc_prepare(maps, qkv_h, out_h, dt, B, N, H)) return 1;
        maps.causal = causal;
        if (attention_tc_run(maps, scale, s)) return 1;
    } else if (attention_tc257_supported(N)) {
        AttentionMaps maps;
        if (attention_tc257_prepare(maps, qkv_h, out_h, dt, B, N, H)) return 1;
        if (attention_tc257_run(maps, scale, s)) return 1;
    } else if (attention_tcl_supported(N)) {
        AttentionMaps maps;
        if (attention_tcl_prepare(maps, qkv_h, out_h, dt, B, N, H)) return 1;
        if (attention_tcl_run(maps, scale, s)) return 1;
    } else if (attention_run(qkv_h, out_h, dt, B, N, H, scale, s)) {
        return 1;
    }
    return uncast_run(out_h, out, dt, rows * H * 64, s);
}

int32_t vidil_op_attention(const float* qkv, float* out, int32_t B, int32_t N, int32_t H, float scale, int32_t dtype,
                           void* workspace, size_t workspace_bytes, void* stream) {
    return op_attention_any(qkv, out, B, N, H, scale, dtype, false, workspace, workspace_bytes, stream);
}

int32_t vidil_op_attention_causal(const float* qkv, float* out, int32_t B, int32_t N, int32_t H, float scale, int32_t dtype,
                                  void* workspace, size_t workspace_bytes, void* stream) {
    return op_attention_any(qkv, out, B, N, H, scale, dtype, true, workspace, workspace_bytes, stream);
}

}  // extern "C"

// =====================================================================================================================
// CLIP text tower (run_visual_tokenization.py:84-96 get_text_embeddings_clip; arithmetic in transformers'
// CLIPTextTransformer): token + position embedding -> depth x [LN -> q,k,v -> causal attention -> out -> +res -> LN ->
// fc1 -> quick_gelu -> fc2 -> +res] -> final LN on the EOS row -> bias-free projection -> L2 normalise.
// Same kernels as the image towers; the attention runs with a causal mask and one 128-row tile per (sequence, head).
// =====================================================================================================================
struct vidil_text_encoder {
    vidil_text_cfg cfg;
    DType dt = DT_BF16;
    DevBuf tok, pos, norm_w, norm_b, head_w;
    std::vector<std::unique_ptr<Layer>> layers;
};

namespace {

struct TextWs {
    size_t resid, xn, qkv, attn, hidden, pooled, pooled_ln, head_out, total;
};

TextWs text_ws(const vidil_text_encoder* e, int B, int L) {
    const size_t M = static_cast<size_t>(B) * L, D = e->cfg.embed_dim;
    TextWs w;
    size_t off = 0;
    w.resid = off;     off += align_up(M * D * 4);
    w.xn = off;        off += align_up(M * D * 2);
    w.qkv = off;       off += align_up(M * 3 * D * 2);
    w.attn = off;      off += align_up(M * D * 2);
    w.hidden = off;    off += align_up(M * e->cfg.mlp_dim * 2);
    w.pooled = off;    off += align_up(static_cast<size_t>(B) * D * 4);
    w.pooled_ln = off; off += align_up(static_cast<size_t>(B) * D * 2);
    w.head_out = off;  off += align_up(static_cast<size_t>(B) * e->cfg.proj_dim * 4);
    w.total = off;
    return w;
}

bool text_slot(vidil_text_encoder* e, const char* name, Slot& s) {
    const vidil_text_cfg& c = e->cfg;
    const int D = c.embed_dim, H = c.mlp_dim;
    auto vec = [&](DevBuf& b, int64_t n) { s = Slot{&b, n, false, 0, 0, 0}; return true; };
    auto mat = [&](DevBuf& b, int r, int k, int ld) { s = Slot{&b, static_cast<int64_t>(r) * k, true, r, k, ld}; return true; };
    if (!strcmp(name, "token_embedding")) return vec(e->tok, static_cast<int64_t>(c.vocab_size) * D);
    if (!strcmp(name, "position_embedding")) return vec(e->pos, static_cast<int64_t>(c.max_positions) * D);
    if (!strcmp(name, "norm.weight")) return vec(e->norm_w, D);
    if (!strcmp(name, "norm.bias")) return vec(e->norm_b, D);
    if (!strcmp(name, "head.proj.weight")) return mat(e->head_w, c.proj_dim, D, D);
    int idx = -1, consumed = 0;
    if (sscanf(name, "blocks.%d.%n", &idx, &consumed) == 1 && consumed > 0 && idx >= 0 && idx < c.depth) {
        const char* r = name + consumed;
        Layer& ly = *e->layers[idx];
        if (!strcmp(r, "norm1.weight")) return vec(ly.ln1_w, D);
        if (!strcmp(r, "norm1.bias")) return vec(ly.ln1_b, D);
        if (!strcmp(r, "attn.qkv.weight")) return mat(ly.qkv_w, 3 * D, D, D);
        if (!strcmp(r, "attn.qkv.bias")) return vec(ly.qkv_b, 3 * D);
        if (!strcmp(r, "attn.proj.weight")) return mat(ly.proj_w, D, D, D);
        if (!strcmp(r, "attn.proj.bias")) return vec(ly.proj_b, D);
        if (!strcmp(r, "norm2.weight")) return vec(ly.ln2_w, D);
        if (!strcmp(r, "norm2.bias")) return vec(ly.ln2_b, D);
        if (!strcmp(r, "mlp.fc1.weight")) return mat(ly.fc1_w, H, D, D);
        if (!strcmp(r, "mlp.fc1.bias")) return vec(ly.fc1_b, H);
        if (!strcmp(r, "mlp.fc2.weight")) return mat(ly.fc2_w, D, H, H);
        if (!strcmp(r, "mlp.fc2.bias")) return vec(ly.fc2_b, D);
    }
    return false;
}

void text_params(const vidil_text_encoder* e, std::vector<NamedBuf>& out) {
    out.push_back({"token_embedding", &e->tok});
    out.push_back({"position_embedding", &e->pos});
    for (int i = 0; i < e->cfg.depth; ++i) {
        const Layer& ly = *e->layers[i];
        const std::string p = "blocks." + std::to_string(i) + ".";
        out.push_back({p + "norm1.weight", &ly.ln1_w});
        out.push_back({p + "norm1.bias", &ly.ln1_b});
        out.push_back({p + "attn.qkv.weight", &ly.qkv_w});
        out.push_back({p + "attn.qkv.bias", &ly.qkv_b});
        out.push_back({p + "attn.proj.weight", &ly.proj_w});
        out.push_back({p + "attn.proj.bias", &ly.proj_b});
        out.push_back({p + "norm2.weight", &ly.ln2_w});
        out.push_back({p + "norm2.bias", &ly.ln2_b});
        out.push_back({p + "mlp.fc1.weight", &ly.fc1_w});
        out.push_back({p + "mlp.fc1.bias", &ly.fc1_b});
        out.push_back({p + "mlp.fc2.weight", &ly.fc2_w});
        out.push_back({p + "mlp.fc2.bias", &ly.fc2_b});
    }
    out.push_back({"norm.weight", &e->norm_w});
    out.push_back({"norm.bias", &e->norm_b});
    out.push_back({"head.proj.weight", &e->head_w});
}

}  // namespace

extern "C" {

int32_t vidil_text_encoder_create(const vidil_text_cfg* cfg, vidil_text_encoder** out) {
    if (cfg == nullptr || out == nullptr) {
        set_error("vidil_text_encoder_create: null argument");
        return 1;
    }
    *out = nullptr;
    vidil_text_cfg c = *cfg;
    if (c.cta_group == 0) c.cta_group = 2;
    if (c.vocab_size <= 0 || c.max_positions <= 0 || c.max_positions > 208) {
        set_error("unsupported text geometry: vocab_size=%d max_positions=%d (sequences of at most 208 tokens)", c.vocab_size,
                  c.max_positions);
        return 1;
    }
    if (c.embed_dim <= 0 || c.embed_dim % 128 != 0 || c.num_heads * 64 != c.embed_dim ||
        (c.embed_dim != 128 && c.embed_dim != 256 && c.embed_dim != 512 && c.embed_dim != 768 && c.embed_dim != 1024 &&
         c.embed_dim != 1280)) {
        set_error("unsupported width: embed_dim=%d num_heads=%d (head_dim must be 64; LayerNorm covers 128/256/512/768/1024/1280)",
                  c.embed_dim, c.num_heads);
        return 1;
    }
    if (c.depth <= 0 || c.mlp_dim <= 0 || c.mlp_dim % 64 != 0 || c.proj_dim <= 0 || c.proj_dim % 4 != 0) {
        set_error("unsupported depth=%d / mlp_dim=%d / proj_dim=%d", c.depth, c.mlp_dim, c.proj_dim);
        return 1;
    }
    if ((c.dtype != VIDIL_DTYPE_BF16 && c.dtype != VIDIL_DTYPE_FP16) || (c.act != VIDIL_ACT_GELU_ERF && c.act != VIDIL_ACT_QUICK_GELU) ||
        (c.cta_group != 1 && c.cta_group != 2)) {
        set_error("unsupported dtype %d / activation %d / cta_group %d", c.dtype, c.act, c.cta_group);
        return 1;
    }
    if (gemm_num_sms() == 0) {
        if (get_error()[0] == 0) set_error("no sm_100 CUDA device available");
        return 1;
    }
    std::unique_ptr<vidil_text_encoder> e(new vidil_text_encoder());
    e->cfg = c;
    e->dt = (c.dtype == VIDIL_DTYPE_BF16) ? DT_BF16 : DT_FP16;
    for (int i = 0; i < c.depth; ++i) e->layers.emplace_back(new Layer());
    std::vector<NamedBuf> names;
    text_params(e.get(), names);
    for (auto& nb : names) {
        Slot s;
        if (!text_slot(e.get(), nb.name.c_str(), s)) {
            set_error("internal: no slot for %s", nb.name.c_str());
            return 1;
        }
        const size_t bytes = s.matrix ? static_cast<size_t>(s.rows) * s.ld * 2 : static_cast<size_t>(s.numel) * 4;
        if (s.buf->alloc(bytes)) return 1;
    }
    *out = e.release();
    return 0;
}

void vidil_text_encoder_destroy(vidil_text_encoder* enc) { delete enc; }

int32_t vidil_text_encoder_load(vidil_text_encoder* enc, const char* name, const float* dev_ptr, int64_t numel, void* stream) {
    if (enc == nullptr || name == nullptr || dev_ptr == nullptr) {
        set_error("vidil_text_encoder_load: null argument");
        return 1;
    }
    Slot s;
    if (!text_slot(enc, name, s)) {
        set_error("vidil_text_encoder_load: unknown parameter '%s'", name);
        return 1;
    }
    if (numel != s.numel) {
        set_error("vidil_text_encoder_load: '%s' has %lld elements, expected %lld", name, (long long)numel, (long long)s.numel);
        return 1;
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (s.matrix) {
        if (cast_run(dev_ptr, s.buf->p, enc->dt, s.rows, s.cols, s.ld, st)) return 1;
    } else {
        VIDIL_CUDA_OK(cudaMemcpyAsync(s.buf->p, dev_ptr, static_cast<size_t>(numel) * 4, cudaMemcpyDeviceToDevice, st));
    }
    s.buf->loaded = true;
    return 0;
}

int32_t vidil_text_encoder_check_loaded(const vidil_text_encoder* enc) {
    if (enc == nullptr) {
        set_error("null encoder");
        return 1;
    }
    std::vector<NamedBuf> names;
    text_params(enc, names);
    for (auto& nb : names)
        if (!nb.buf->loaded) {
            set_error("parameter '%s' has not been loaded", nb.name.c_str());
            return 1;
        }
    return 0;
}

size_t vidil_text_encoder_workspace_bytes(const vidil_text_encoder* enc, int32_t batch, int32_t seq_len) {
    if (enc == nullptr || batch <= 0 || seq_len <= 0) return 0;
    return text_ws(enc, batch, seq_len).total;
}

int32_t vidil_clip_text_forward(vidil_text_encoder* enc, const int32_t* input_ids, const int32_t* eos_pos, int32_t batch,
                                int32_t seq_len, float* out_embeds, void* workspace, size_t workspace_bytes, void* stream) {
    if (enc == nullptr || input_ids == nullptr || eos_pos == nullptr || out_embeds == nullptr || workspace == nullptr) {
        set_error("vidil_clip_text_forward: null argument");
        return 1;
    }
    const vidil_text_cfg& c = enc->cfg;
    if (batch <= 0 || seq_len <= 0 || seq_len > c.max_positions) {
        set_error("vidil_clip_text_forward: batch=%d seq_len=%d (max_positions %d)", batch, seq_len, c.max_positions);
        return 1;
    }
    if (vidil_text_encoder_check_loaded(enc)) return 1;
    const TextWs L = text_ws(enc, batch, seq_len);
    if (workspace_bytes < L.total || (reinterpret_cast<uintptr_t>(workspace) & (ALIGN - 1))) {
        set_error("vidil_clip_text_forward: workspace too small (%zu < %zu) or not %zu-byte aligned", workspace_bytes, L.total, ALIGN);
        return 1;
    }
    if (gemm_num_sms() == 0) return 1;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    uint8_t* base = reinterpret_cast<uint8_t*>(workspace);
    float* resid = reinterpret_cast<float*>(base + L.resid);
    void* xn = base + L.xn;
    void* qkv = base + L.qkv;
    void* attn = base + L.attn;
    void* hidden = base + L.hidden;
    float* pooled = reinterpret_cast<float*>(base + L.pooled);
    void* pooled_ln = base + L.pooled_ln;
    float* head_out = reinterpret_cast<float*>(base + L.head_out);
    const int M = batch * seq_len, D = c.embed_dim;

    if (embed_tokens_run(input_ids, reinterpret_cast<const float*>(enc->tok.p), reinterpret_cast<const float*>(enc->pos.p), resid, M,
                         seq_len, D, c.vocab_size, s))
        return 1;
    AttentionMaps maps;
    if (attention_tc_prepare(maps, qkv, attn, enc->dt, batch, seq_len, c.num_heads)) return 1;
    maps.causal = true;
    GemmProblem g;
    g.dt = enc->dt;
    g.cta_group = c.cta_group;
    for (int i = 0; i < c.depth; ++i) {
        Layer& ly = *enc->layers[i];
        if (layernorm_run(resid, D, reinterpret_cast<const float*>(ly.ln1_w.p), reinterpret_cast<const float*>(ly.ln1_b.p), xn, false,
                          enc->dt, M, D, c.ln_eps, s))
            return 1;
        GemmProblem q = g;
        q.epi = EPI_STORE; q.M = M; q.N = 3 * D; q.K = D; q.A = xn; q.lda = D; q.W = ly.qkv_w.p; q.ldw = D;
        q.bias = reinterpret_cast<const float*>(ly.qkv_b.p); q.out = qkv; q.ldo = 3 * D;
        if (gemm_prepare(q) || gemm_run(q, s)) return 1;
        if (attention_tc_run(maps, 0.125f, s)) return 1;
        GemmProblem p = g;
        p.epi = EPI_RESID; p.M = M; p.N = D; p.K = D; p.A = attn; p.lda = D; p.W = ly.proj_w.p; p.ldw = D;
        p.bias = reinterpret_cast<const float*>(ly.proj_b.p); p.out = resid; p.ldo = D;
        if (gemm_prepare(p) || gemm_run(p, s)) return 1;
        if (layernorm_run(resid, D, reinterpret_cast<const float*>(ly.ln2_w.p), reinterpret_cast<const float*>(ly.ln2_b.p), xn, false,
                          enc->dt, M, D, c.ln_eps, s))
            return 1;
        GemmProblem f1 = g;
        f1.epi = (c.act == VIDIL_ACT_QUICK_GELU) ? EPI_QUICKGELU : EPI_GELU; f1.M = M; f1.N = c.mlp_dim; f1.K = D; f1.A = xn; f1.lda = D;
        f1.W = ly.fc1_w.p; f1.ldw = D; f1.bias = reinterpret_cast<const float*>(ly.fc1_b.p); f1.out = hidden; f1.ldo = c.mlp_dim;
        if (gemm_prepare(f1) || gemm_run(f1, s)) return 1;
        GemmProblem f2 = g;
        f2.epi = EPI_RESID; f2.M = M; f2.N = D; f2.K = c.mlp_dim; f2.A = hidden; f2.lda = c.mlp_dim; f2.W = ly.fc2_w.p; f2.ldw = c.mlp_dim;
        f2.bias = reinterpret_cast<const float*>(ly.fc2_b.p); f2.out = resid; f2.ldo = D;
        if (gemm_prepare(f2) || gemm_run(f2, s)) return 1;
    }
    // pooled_output = final_layer_norm(hidden)[b, eos_pos[b]]: LayerNorm is per row, so gather first
    if (gather_rows_run(resid, eos_pos, pooled, batch, seq_len, D, s)) return 1;
    if (layernorm_run(pooled, D, reinterpret_cast<const float*>(enc->norm_w.p), reinterpret_cast<const float*>(enc->norm_b.p), pooled_ln,
                      false, enc->dt, batch, D, c.ln_eps, s))
        return 1;
    GemmProblem h = g;
    h.epi = EPI_STORE_F32; h.M = batch; h.N = c.proj_dim; h.K = D; h.A = pooled_ln; h.lda = D; h.W = enc->head_w.p; h.ldw = D;
    h.bias = nullptr; h.out = head_out; h.ldo = c.proj_dim;
    if (gemm_prepare(h) || gemm_run(h, s)) return 1;
    return l2norm_run(head_out, out_embeds, batch, c.proj_dim, s);
}

}  // extern "C"
