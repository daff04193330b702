// Fused softmax attention on tcgen05 for sequences that fit one key tile (N <= 208 tokens, head_dim 64) — the
// ViT-L/16 @224 shape (N = 197) and the CLIP text tower (N = 77, causal).  Reference: models/vit.py:72-83
// (Attention.forward); the reference materialises [B,H,N,N] scores in HBM, here S and P never leave tensor memory.
//
// One persistent CTA per SM walks (frame, head) items, last frame first: the qkv GEMM wrote its rows in ascending order, so the
// end of the 310 MB qkv matrix is what the 126 MB L2 still holds when this kernel starts, and the projection GEMM that follows
// starts at the rows written last here.  Per item the queries form up to two 128-row tiles:
//   S = Q_t K^T          tcgen05.mma M=128, N=ceil16(keys), K=64;  fp32 S in TMEM columns [0, 208) of the lane's 256
//   P = exp2(c S - c max)   ONE thread per query row: pass 1 reads the row from TMEM and takes its maximum (3-input
//                           max), pass 2 reads it again, exponentiates, accumulates the row sum and writes the 16-bit
//                           probabilities straight back to TMEM (tcgen05.st), two per 32-bit column, over the S columns
//                           it has already consumed: P occupies columns [0, 104)
//   O = P V              tcgen05.mma M=128, N=64, K=keys with the A operand read FROM TMEM (no shared-memory round trip,
//                        no generic->async proxy fence); V is consumed straight from its TMA tile as an MN-major B
//                        operand (no transpose anywhere); fp32 O in columns [128, 192)
//   out = O / rowsum     -> 16-bit -> swizzled staging -> 3-D TMA store (rows past the frame's last token clipped)
// The CTA runs two independent "lanes" L = 0,1, each with its own MMA-issuing thread, 256 TMEM columns and 4 softmax
// warps (one per SM sub-partition: warp w may touch TMEM lanes 32 (w % 4) .. +31); lane L takes tile (L + item) & 1, so
// the 128-row and the 69-row tile alternate between the lanes and each lane's S -> softmax -> PV -> output chain overlaps
// the other lane's.  The exponential phase (the MUFU-bound part) is handed back and forth between the lanes like a token.
// Warp 0 is the TMA producer (Q/K are reloaded as soon as both S MMAs of an item retire, V as soon as both PV MMAs do).
// q/k/v are read in place from the fused-QKV GEMM output [B, N, 3, H, 64]; out is [B, N, H*64] (vit.py:83's
// transpose(1,2).reshape), ready to be the proj GEMM's A operand.
#include <math.h>
#include <stdlib.h>

#include <type_traits>

#include "attention_common.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace vidil {
namespace {

constexpr int HD = 64;
constexpr int QT = 128;          // query rows per tile (UMMA M)
constexpr int MAX_KEYS = 208;    // 13 x 16: S (fp32) fits the lane's TMEM columns below O, K/V tiles fit one TMA box
constexpr int O_COL = 128;       // O accumulator columns [128, 192) of the lane's 256: above P [0, 104), inside dead S
constexpr int ATC_THREADS = 384; // warp 0 producer, 1-2 MMA issuers, 3 TMEM allocator, 4-7 / 8-11 softmax warps of lane 0 / 1

constexpr int Q_BYTES = QT * HD * 2;               // 16 KB
constexpr int KV_BYTES = MAX_KEYS * HD * 2;        // 26 KB
// Q (two tiles), K and V are double-buffered: the TMA producer runs one (frame, head) item ahead, so the next item's tiles
// travel from HBM while the current item is computed (in a forward the qkv matrix does not fit L2; with single buffers the
// load latency of every item was exposed: 6 400 clocks per item cold against 4 700 warm)
constexpr int STAGES = 2;
constexpr int OFF_Q = 0;
constexpr int OFF_K = STAGES * 2 * Q_BYTES;
constexpr int OFF_V = OFF_K + STAGES * KV_BYTES;
constexpr int OFF_OUT = OFF_V + STAGES * KV_BYTES;  // 2 lanes x 4 quarters x [32 rows][128 B]
constexpr int OFF_BAR = OFF_OUT + 8 * 4096;
constexpr int ATC_SMEM = OFF_BAR + 256 + 1024;     // + alignment slack
static_assert(OFF_K % 1024 == 0 && OFF_V % 1024 == 0 && OFF_OUT % 1024 == 0, "tile alignment");
static_assert(ATC_SMEM <= 227 * 1024, "shared memory budget");
static_assert(MAX_KEYS / 2 <= O_COL && O_COL + HD <= 256 && MAX_KEYS <= 256, "TMEM column plan");

using namespace attn;

// Compile-time shape of a row for the specialised kernels.
template <int NCT>
struct RowShape {
    static constexpr int NK16 = (NCT + 15) & ~15;
    static constexpr int NPAIR = NK16 / 2;            // pairs of keys = packed P columns
    static constexpr int NCH = (NK16 + 31) / 32;      // 32-column chunks, the last one possibly 16 wide
    static constexpr bool SPLIT = NK16 > 128;         // PV issued in two parts: keys [0,128) while the rest is still exponentiated
    static constexpr int O_COL = SPLIT ? 64 : 128;    // O accumulator columns (see the column plan at the kernel)
    __host__ __device__ static constexpr int pairs_in(int ch) { return (NPAIR - 16 * ch) < 16 ? (NPAIR - 16 * ch) : 16; }
    // packed P column of key pair `pair`: keys >= 128 live above the O accumulator when the PV is split
    __host__ __device__ static constexpr int pcol(int pair) { return (SPLIT && pair >= 64) ? 128 + (pair - 64) : pair; }
};

// Pass 2 of a row for a compile-time sequence length, as ONE software-pipelined stream over the row's key pairs instead of a
// loop over chunks: pair p is exponentiated (MUFU or polynomial) LAGP pairs ahead of the point where its results are summed
// and packed, ACROSS chunk boundaries; the next chunk's S columns are waited for, turned into exponent arguments and the
// chunk after that requested from TMEM in the middle of the current chunk's exponentials.  The MUFU pipe (one warp-wide ex2
// per 8 clocks) then never waits for a chunk's head (TMEM wait, 16 FFMA2) or tail (adds, packs, TMEM store): measured per
// 32-column chunk, one warp per sub-partition: 338 clocks chunk by chunk, 262 with 6 of 16 pairs on the polynomial, against
// 256 / 160 of pure MUFU time.
template <typename T, int NCT, int POLY>
__device__ __forceinline__ void softmax_stream(uint32_t taddr, float c, float neg_mxs, float (&sum)[4], uint64_t* tok_bar,
                                               uint64_t* p_half_bar, bool lane0) {
    using RS = RowShape<NCT>;
    constexpr int LAGP = 6;
    float a[2][32];
    uint32_t sreg[32];
    uint32_t pk[16];
    auto load = [&](int ch) {
        if (32 * ch + 32 <= RS::NK16) {
            ptx::tmem_ld_32x32b_x32(taddr + 32 * ch, sreg);
        } else {
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(sreg[0]), "=r"(sreg[1]), "=r"(sreg[2]), "=r"(sreg[3]), "=r"(sreg[4]), "=r"(sreg[5]), "=r"(sreg[6]),
                  "=r"(sreg[7]), "=r"(sreg[8]), "=r"(sreg[9]), "=r"(sreg[10]), "=r"(sreg[11]), "=r"(sreg[12]), "=r"(sreg[13]),
                  "=r"(sreg[14]), "=r"(sreg[15])
                : "r"(taddr + 32 * ch)
                : "memory");
        }
    };
    auto args = [&](int ch) {
#pragma unroll
        for (int q = 0; q < 16; ++q)
            if (q < RS::pairs_in(ch) && 32 * ch + 2 * q < NCT)
                ptx::fma2(a[ch & 1][2 * q], a[ch & 1][2 * q + 1], __uint_as_float(sreg[2 * q]), __uint_as_float(sreg[2 * q + 1]), c,
                          neg_mxs);
    };
    load(0);
    ptx::tmem_ld_wait();
    args(0);
    if (RS::NCH > 1) load(1);
#pragma unroll
    for (int p = 0; p < RS::NPAIR + LAGP; ++p) {
        if (p < RS::NPAIR) {
            const int ch = p >> 4, q = p & 15;
            if (q == RS::pairs_in(ch) / 2 && ch + 1 < RS::NCH) {
                ptx::tmem_ld_wait();
                args(ch + 1);
                if (ch + 2 < RS::NCH) load(ch + 2);
            }
            if (q == 0 && ch == RS::NCH - 1 && lane0) ptx::mbar_arrive(tok_bar);  // entering the last chunk: hand the MUFU phase over
            if (2 * p < NCT) {
                if (poly_pair<POLY>(q)) {
                    exp2_poly2(a[ch & 1][2 * q], a[ch & 1][2 * q + 1]);
                } else {
                    a[ch & 1][2 * q] = ptx::ex2_approx(a[ch & 1][2 * q]);
                    a[ch & 1][2 * q + 1] = ptx::ex2_approx(a[ch & 1][2 * q + 1]);
                }
            }
        }
        if (p >= LAGP) {
            const int j = p - LAGP, cj = j >> 4, qj = j & 15;
            float x0 = 0.f, x1 = 0.f;  // keys >= N are padding: P = 0
            if (2 * j < NCT) x0 = a[cj & 1][2 * qj];
            if (2 * j + 1 < NCT) x1 = a[cj & 1][2 * qj + 1];
            if (2 * j < NCT) ptx::add2(sum[2 * (j & 1)], sum[2 * (j & 1) + 1], x0, x1);
            pk[qj] = pack2<T>(x0, x1);
            if (qj == RS::pairs_in(cj) - 1) {  // chunk cj is complete
                if (RS::SPLIT && cj == 4) {
                    // keys [0,128) = chunks 0..3 were stored a whole chunk ago: the wait returns at once, and the MMA issuer may
                    // start the first eight PV k-steps while this warp exponentiates the rest of the row
                    ptx::tmem_st_wait();
                    ptx::tcgen05_fence_before();
                    __syncwarp();
                    if (lane0) ptx::mbar_arrive(p_half_bar);
                }
                if (RS::pairs_in(cj) == 16) {
                    ptx::tmem_st_32x32b_x16(taddr + RS::pcol(16 * cj), pk);
                } else {
                    uint32_t pk8[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) pk8[i] = pk[i];
                    ptx::tmem_st_32x32b_x8(taddr + RS::pcol(16 * cj), pk8);
                }
            }
        }
    }
}

// NCT > 0: the sequence length is a compile-time constant (the headline shapes).  Every chunk loop then unrolls completely:
// no loop control between the chunks of a row (measured on the rolled version: 52 M executed instructions of which 19 M were
// softmax arithmetic, ~600 clocks per 32-column chunk against 338 for the bare chunk), the padding mask of the last chunk
// folds away, and the MMA descriptors of all k-steps are constants.  NCT == 0: any N <= 208 at run time, causal or not.
template <typename T, int NCT, int POLY>
__global__ void __launch_bounds__(ATC_THREADS, 1)
    attention_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv,
                        const __grid_constant__ CUtensorMap map_out, int n_items, int N_rt, int H, float scale_log2e,
                        int causal_rt, long long* trace, int flags) {
    const int N = NCT > 0 ? NCT : N_rt;
    // TMEM columns of a lane (256): generic kernel: S [0,208) -> P [0,104), O [128,192).  Specialised kernel with more than
    // 128 keys: P of keys [0,128) in [0,64), O in [64,128) (S columns dead once chunks 0..3 are consumed), P of keys >= 128 in
    // [128,168) — the first eight PV k-steps run under the second half of the row's exponentials.
    constexpr bool kSplit = NCT > 0 && RowShape<NCT>::SPLIT;
    constexpr int kOCol = NCT > 0 ? RowShape<NCT>::O_COL : O_COL;
    const int causal = NCT > 0 ? 0 : causal_rt;
    // Developer timeline (vidil_debug_set_attention_trace): CTA 0 stamps clock64 at its synchronisation points for 16 items.
#define ATC_TRACE(role, ev)                                                                                   \
    do {                                                                                                      \
        if (trace != nullptr && blockIdx.x == 0 && it < 16) trace[((role) * 16 + it) * 8 + (ev)] = clock64(); \
    } while (0)
#define ATC_TRACE_S(ev)                                    \
    do {                                                   \
        if (quarter == 0 && lane == 0) ATC_TRACE(3 + L, ev); \
    } while (0)
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    const uint32_t sbase = ptx::smem_u32(smem);

    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* qk_full = bars + 0;   // [2] per stage
    uint64_t* v_full = bars + 2;    // [2]
    uint64_t* qk_empty = bars + 4;  // [2]
    uint64_t* v_empty = bars + 6;   // [2]
    uint64_t* s_full = bars + 8;    // [2] per lane
    uint64_t* p_full = bars + 10;   // [2]
    uint64_t* o_full = bars + 12;   // [2]
    uint64_t* s_free = bars + 14;   // [2]
    uint64_t* tok = bars + 16;      // [2] "this lane is in the last chunk of its exponentials"
    uint64_t* p_half = bars + 18;   // [2] keys [0,128) of P are in TMEM (specialised kernels with more than 128 keys)
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 20);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int nk16 = (N + 15) & ~15;          // keys padded to the UMMA N / K granularity
    const int nfull = nk16 >> 5;              // 32-column chunks of an S row, plus one 16-column tail chunk
    const bool tail16 = (nk16 & 16) != 0;
    const int n_tiles = (N + QT - 1) / QT;    // 1 or 2
    const int D = H * HD;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&map_q);
        ptx::prefetch_tensormap(&map_kv);
        ptx::prefetch_tensormap(&map_out);
        for (int st = 0; st < STAGES; ++st) {
            ptx::mbar_init(&qk_full[st], 1);
            ptx::mbar_init(&v_full[st], 1);
            ptx::mbar_init(&qk_empty[st], n_tiles);  // one commit per lane that issued an S MMA for the item
            ptx::mbar_init(&v_empty[st], n_tiles);
        }
        for (int l = 0; l < 2; ++l) {
            ptx::mbar_init(&s_full[l], 1);
            ptx::mbar_init(&p_full[l], 4);  // lane 0 of each of the lane's 4 softmax warps
            ptx::mbar_init(&o_full[l], 1);
            ptx::mbar_init(&s_free[l], 4);  // each softmax warp, once its O rows are in registers
            ptx::mbar_init(&tok[l], 4);     // each softmax warp, when it enters its last chunk
            ptx::mbar_init(&p_half[l], 4);
        }
        ptx::fence_mbar_init();
    }
    if (warp == 3) ptx::tmem_alloc<1>(tmem_ptr_smem, 512);
    ptx::tcgen05_fence_before();
    __syncthreads();
    ptx::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    // PDL: set-up done; let the next kernel start its own, then wait for the predecessor's qkv before any global access
    ptx::griddep_launch();
    ptx::griddep_wait();

    if (warp == 0) {
        if (lane == 0) {
            // ===================== TMA producer =====================
            int it = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const int st = it & 1;
                const uint32_t ph = (it >> 1) & 1;
                const int ritem = (flags & 64) ? item : n_items - 1 - item;  // last frames first: see the header (L2 residency of the freshly written qkv rows)
                const int b = ritem / H, h = ritem - b * H;
                const int row0 = b * N;
                uint8_t* sq = smem + OFF_Q + st * 2 * Q_BYTES;
                ATC_TRACE(0, 0);
                ptx::mbar_wait(&qk_empty[st], ph ^ 1);  // the S MMAs of the item two back have consumed this stage's Q and K
                ATC_TRACE(0, 1);
                ptx::mbar_arrive_expect_tx(&qk_full[st], n_tiles * Q_BYTES + nk16 * HD * 2);
                ptx::tma_load_2d(&map_q, &qk_full[st], sq, h * HD, row0);
                if (n_tiles == 2) ptx::tma_load_2d(&map_q, &qk_full[st], sq + Q_BYTES, h * HD, row0 + QT);
                ptx::tma_load_2d(&map_kv, &qk_full[st], smem + OFF_K + st * KV_BYTES, D + h * HD, row0);
                ptx::mbar_wait(&v_empty[st], ph ^ 1);  // the PV MMAs of the item two back have consumed this stage's V
                ATC_TRACE(0, 2);
                ptx::mbar_arrive_expect_tx(&v_full[st], nk16 * HD * 2);
                ptx::tma_load_2d(&map_kv, &v_full[st], smem + OFF_V + st * KV_BYTES, 2 * D + h * HD, row0);
            }
        }
        __syncwarp();
    } else if (warp == 1 || warp == 2) {
        if (lane == 0) {
            // ===================== MMA issuer of lane L =====================
            const int L = warp - 1;
            constexpr bool kIsBf16 = std::is_same<T, __nv_bfloat16>::value;
            const uint32_t idesc_s = ptx::make_idesc_f16(kIsBf16, QT, nk16);
            const uint32_t idesc_o = ptx::make_idesc_f16_bmn(kIsBf16, QT, HD);
            const int ksteps = nk16 / 16;
            const uint32_t tmem_s = tmem_base + L * 256;  // S, then P packed over its first columns
            const uint32_t tmem_o = tmem_s + kOCol;
            int it = 0, n = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const int t = (n_tiles == 2) ? ((L + it) & 1) : L;  // one tile: lane 0 does every item, lane 1 idles
                if (t >= n_tiles) continue;
                const int st = it & 1;
                const uint32_t ph = (it >> 1) & 1, par = n & 1;
                ++n;
                ATC_TRACE(1 + L, 0);
                ptx::mbar_wait(&qk_full[st], ph);
                ATC_TRACE(1 + L, 1);
                ptx::mbar_wait(&s_free[L], par ^ 1);  // this lane's previous O (inside the S columns) has been read out
                ATC_TRACE(1 + L, 2);
                ptx::tcgen05_fence_after();
                const uint64_t dk = ptx::make_kmajor_sw128_desc(sbase + OFF_K + st * KV_BYTES);
                const uint64_t dq = ptx::make_kmajor_sw128_desc(sbase + OFF_Q + (st * 2 + t) * Q_BYTES);
                const uint32_t sv = sbase + OFF_V + st * KV_BYTES;
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) ptx::umma_f16<1>(tmem_s, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
                ptx::umma_commit<1>(&s_full[L]);
                ptx::umma_commit<1>(&qk_empty[st]);
                ATC_TRACE(1 + L, 3);
                ptx::mbar_wait(&v_full[st], ph);
                ATC_TRACE(1 + L, 4);
                if constexpr (kSplit) {
                    ptx::mbar_wait(&p_half[L], par);  // P of keys [0,128) is in TMEM columns [0,64); S columns [0,128) are dead
                    ptx::tcgen05_fence_after();
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint64_t db = ptx::make_smem_desc(sv + j * 2048, 0, 1024, 2);
                        ptx::umma_f16_tmem_a(tmem_o, tmem_s + j * 8, db, idesc_o, j != 0);
                    }
                }
                ptx::mbar_wait(&p_full[L], par);  // all of P is in TMEM, S has been consumed
                ATC_TRACE(1 + L, 5);
                ptx::tcgen05_fence_after();
#pragma unroll
                for (int j = kSplit ? 8 : 0; j < ksteps; ++j) {
                    // A: P k-step j = 16 keys = 8 packed TMEM columns
                    // B: V rows 16j..16j+15 (two 8-row 1024-byte swizzle groups), 64 contiguous channels per row
                    const uint64_t db = ptx::make_smem_desc(sv + j * 2048, 0, 1024, 2);
                    ptx::umma_f16_tmem_a(tmem_o, tmem_s + (kSplit ? 128 + (j - 8) * 8 : j * 8), db, idesc_o, j != 0);
                }
                ptx::umma_commit<1>(&o_full[L]);
                ptx::umma_commit<1>(&v_empty[st]);
                ATC_TRACE(1 + L, 6);
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ===================== softmax + output warps of lane L =====================
        const int L = (warp - 4) >> 2;
        const int quarter = warp & 3;            // TMEM lanes this warp may touch: 32 * (warp % 4) ...
        const uint32_t group_bar = 1 + L;        // named barrier of the lane's 4 softmax warps
        const bool poller = quarter == 0;
        // A warp spinning on an mbarrier issues a few instructions every ~20 clocks, taking issue slots from the warp of the
        // other lane on the same sub-partition: one warp per lane polls, the other three sleep in a hardware named barrier.
        auto group_wait = [&](uint64_t* bar, uint32_t parity) {
            if (poller) ptx::mbar_wait(bar, parity);
            ptx::named_bar_sync(group_bar, 128);
        };
        const int row_in_tile = quarter * 32 + lane;
        const uint32_t stage = sbase + OFF_OUT + (L * 4 + quarter) * 4096;
        const void* stage_ptr = smem + OFF_OUT + (L * 4 + quarter) * 4096;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + L * 256;
        int it = 0, n = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const int t = (n_tiles == 2) ? ((L + it) & 1) : L;  // one tile: lane 0 does every item, lane 1 idles
            if (t >= n_tiles) continue;
            const uint32_t par = n & 1;
            ++n;
            const int ritem = (flags & 64) ? item : n_items - 1 - item;
            const int b = ritem / H, h = ritem - b * H;
            const bool warp_valid = (t * QT + quarter * 32) < N;
            // keys this query row may see: all N, or 0..row under a causal mask (padding rows see nothing that is kept)
            const int NV = causal ? min(N, t * QT + row_in_tile + 1) : N;
            float sum[4] = {0.f, 0.f, 0.f, 0.f};

            ATC_TRACE_S(0);
            group_wait(&s_full[L], par);
            ATC_TRACE_S(1);
            ptx::tcgen05_fence_after();
            float neg_mxs = 0.f;
            if (warp_valid) {
                // ---- pass 1: row maximum; two TMEM loads in flight per wait ----
                float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
                uint32_t ra[32], rb[32];
#pragma unroll
                for (int c = 0; c < nfull; c += 2) {
                    ptx::tmem_ld_32x32b_x32(taddr + 32 * c, ra);
                    if (c + 1 < nfull) ptx::tmem_ld_32x32b_x32(taddr + 32 * c + 32, rb);
                    ptx::tmem_ld_wait();
                    chunk_max<32>(ra, 32 * c, NV, mx);
                    if (c + 1 < nfull) chunk_max<32>(rb, 32 * c + 32, NV, mx);
                }
                if (tail16) {
                    uint32_t rt[16];
                    ptx::tmem_ld_32x32b_x16(taddr + 32 * nfull, rt);
                    ptx::tmem_ld_wait();
                    chunk_max<16>(rt, 32 * nfull, NV, mx);
                }
                // key 0 is visible to every row, so the maximum is finite
                neg_mxs = -fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * scale_log2e;
            }
            // Pass the exponential phase back and forth between the lanes, so that one lane's softmax runs under the other
            // lane's MMAs / barriers / output instead of both lanes doing the same phase at the same time (they fall into
            // lockstep otherwise, because they share the Q/K/V buffers).
            ATC_TRACE_S(2);
            if (n_tiles == 2 && !(flags & 1)) group_wait(&tok[L ^ 1], L == 0 ? (par ^ 1) : par);
            ATC_TRACE_S(3);
            if (NCT > 0 && warp_valid) {
                softmax_stream<T, NCT, POLY>(taddr, scale_log2e, neg_mxs, sum, &tok[L], &p_half[L], lane == 0);
                ptx::tmem_st_wait();
            } else if (warp_valid) {
                // ---- pass 2: exponentials, row sum, P -> TMEM; the next chunk's TMEM load is in flight under the math.
                //      Entering its last chunk, the warp hands the exponential phase over to the other lane. ----
                uint32_t ra[32], rb[32], rt[16];
                auto prefetch = [&](int c, uint32_t (&buf)[32]) {  // chunk c (or the 16-column tail, or nothing) while chunk c-1 computes
                    if (c < nfull)
                        ptx::tmem_ld_32x32b_x32(taddr + 32 * c, buf);
                    else if (tail16)
                        ptx::tmem_ld_32x32b_x16(taddr + 32 * nfull, rt);
                    if (c >= nfull + (tail16 ? 1 : 0) && lane == 0) ptx::mbar_arrive(&tok[L]);  // chunk c-1 is the last one
                };
                prefetch(0, ra);
#pragma unroll
                for (int c = 0; c < nfull; c += 2) {
                    ptx::tmem_ld_wait();
                    prefetch(c + 1, rb);
                    chunk_exp<T, 32, POLY>(ra, 32 * c, NV, scale_log2e, neg_mxs, taddr, sum);
                    if (c + 1 < nfull) {
                        ptx::tmem_ld_wait();
                        prefetch(c + 2, ra);
                        chunk_exp<T, 32, POLY>(rb, 32 * c + 32, NV, scale_log2e, neg_mxs, taddr, sum);
                    }
                }
                if (tail16) {
                    ptx::tmem_ld_wait();
                    if (lane == 0) ptx::mbar_arrive(&tok[L]);
                    chunk_exp<T, 16, POLY>(rt, 32 * nfull, NV, scale_log2e, neg_mxs, taddr, sum);
                }
                ptx::tmem_st_wait();
            } else if (lane == 0) {
                ptx::mbar_arrive(&tok[L]);  // rows past the sequence: nothing to do, pass the turn on
                if (kSplit) ptx::mbar_arrive(&p_half[L]);
            }
            ptx::tcgen05_fence_before();  // P (tcgen05.st) and the S reads, before the MMA issuer's PV
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&p_full[L]);
            ATC_TRACE_S(4);

            group_wait(&o_full[L], par);
            ATC_TRACE_S(5);
            ptx::tcgen05_fence_after();
            const uint32_t srow = stage + lane * 128;
            const uint32_t swz = static_cast<uint32_t>(lane & 7);
            const float inv_sum = 1.0f / ((sum[0] + sum[1]) + (sum[2] + sum[3]));
            // O leaves TMEM in two 32-column halves (64 live registers would spill next to the softmax buffers)
            auto drain_half = [&](int half) {
                uint32_t o[32];
                ptx::tmem_ld_32x32b_x32(taddr + kOCol + 32 * half, o);
                // the staging tile is about to be reused: its previous TMA store must have read it
                if (half == 0 && lane == 0) ptx::bulk_wait_group_read<0>();
                ptx::tmem_ld_wait();
                if (half == 0) __syncwarp();
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t u[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float a0, a1;
                        ptx::mul2(a0, a1, __uint_as_float(o[8 * c + 2 * i]), __uint_as_float(o[8 * c + 2 * i + 1]), inv_sum);
                        u[i] = pack2<T>(a0, a1);
                    }
                    ptx::st_shared_v4(srow + ((static_cast<uint32_t>(4 * half + c) ^ swz) << 4), u[0], u[1], u[2], u[3]);
                }
            };
            if (warp_valid) {
                drain_half(0);
                drain_half(1);
            }
            ptx::tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&s_free[L]);  // the lane's TMEM columns may take the next S
            ATC_TRACE_S(6);
            if (warp_valid) {
                ptx::fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    ptx::tma_store_3d(&map_out, stage_ptr, h * HD, t * QT + quarter * 32, b);
                    ptx::bulk_commit_group();
                }
            }
        }
        if (lane == 0) ptx::bulk_wait_group<0>();
    }

    ptx::tcgen05_fence_before();
    __syncthreads();
    if (warp == 3) ptx::tmem_dealloc<1>(tmem_base, 512);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

long long* g_trace = nullptr;  // developer hook, see attention_set_trace
// developer A/B switches (VIDIL_ATC_FLAGS): 1 = no token between the lanes, 16 = generic kernel for N = 197, 32 = all-MUFU exponentials,
// 64 = items in ascending order
const int g_flags = [] { const char* e = getenv("VIDIL_ATC_FLAGS"); return e ? atoi(e) : 0; }();

template <typename T, int NCT, int POLY>
int launch_tc_n(const AttentionMaps& m, int B, int N, int H, float scale_log2e, cudaStream_t stream) {
    auto kern = attention_tc_kernel<T, NCT, POLY>;
    // per launch: the attribute belongs to the current device's copy of the function
    VIDIL_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ATC_SMEM));
    const int n_items = B * H;
    int grid = gemm_num_sms();
    if (grid > n_items) grid = n_items;
    if (grid < 1) return 1;
    VIDIL_CUDA_OK(launch_pdl(kern, dim3(grid), dim3(ATC_THREADS), ATC_SMEM, stream, m.q, m.kv, m.out, n_items, N, H, scale_log2e,
                             m.causal ? 1 : 0, g_trace, g_flags));
    count_launches(1);
    return 0;
}

template <typename T>
int launch_tc(const AttentionMaps& m, int B, int N, int H, float scale_log2e, cudaStream_t stream) {
    if (N == 197 && !m.causal && !(g_flags & 16)) {  // ViT-*/16 @224
        // 6 of every 16 pairs of exponentials on the FMA pipe (measured per layer, isolated: 0 -> 105.1, 4 -> 106.9,
        // 6 -> 102.4, 8 -> 106.9 us; same max / mean error against fp64 as the all-MUFU version)
        if (g_flags & 32) return launch_tc_n<T, 197, 0>(m, B, N, H, scale_log2e, stream);  // developer A/B
        return launch_tc_n<T, 197, 6>(m, B, N, H, scale_log2e, stream);
    }
    return launch_tc_n<T, 0, 0>(m, B, N, H, scale_log2e, stream);
}

}  // namespace

void attention_tc257_set_trace(long long* dev_buf);
void attention_set_trace(long long* dev_buf) {
    g_trace = dev_buf;
    attention_tc257_set_trace(dev_buf);
}

bool attention_tc_supported(int N) { return N >= 1 && N <= MAX_KEYS; }

int attention_tc_prepare(AttentionMaps& m, const void* qkv, void* out, DType dt, int B, int N, int H) {
    EncodeTiledFn fn = encode_fn();
    if (fn == nullptr) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return 1;
    }
    if ((reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) {
        set_error("attention: qkv and out must be 16-byte aligned");
        return 1;
    }
    const CUtensorMapDataType cdt = (dt == DT_BF16) ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    const int nk16 = (N + 15) & ~15;
    const cuuint32_t estr3[3] = {1, 1, 1};
    {
        const cuuint64_t dims[2] = {static_cast<cuuint64_t>(3) * H * HD, static_cast<cuuint64_t>(B) * N};
        const cuuint64_t strides[1] = {static_cast<cuuint64_t>(3) * H * HD * 2};
        const cuuint32_t box_q[2] = {HD, QT};
        const cuuint32_t box_kv[2] = {HD, static_cast<cuuint32_t>(nk16)};
        CUresult r = fn(&m.q, cdt, 2, const_cast<void*>(qkv), dims, strides, box_q, estr3, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r == CUDA_SUCCESS)
            r = fn(&m.kv, cdt, 2, const_cast<void*>(qkv), dims, strides, box_kv, estr3, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("attention: cuTensorMapEncodeTiled(qkv) failed with CUresult %d", static_cast<int>(r));
            return 1;
        }
    }
    {
        const cuuint64_t dims[3] = {static_cast<cuuint64_t>(H) * HD, static_cast<cuuint64_t>(N), static_cast<cuuint64_t>(B)};
        const cuuint64_t strides[2] = {static_cast<cuuint64_t>(H) * HD * 2, static_cast<cuuint64_t>(N) * H * HD * 2};
        const cuuint32_t box[3] = {HD, 32, 1};
        CUresult r = fn(&m.out, cdt, 3, out, dims, strides, box, estr3, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("attention: cuTensorMapEncodeTiled(out) failed with CUresult %d", static_cast<int>(r));
            return 1;
        }
    }
    m.qkv = qkv;
    m.out_ptr = out;
    m.B = B;
    m.N = N;
    m.H = H;
    m.dt = dt;
    return 0;
}

int attention_tc_run(const AttentionMaps& m, float scale, cudaStream_t stream) {
    if (gemm_num_sms() == 0) return 1;
    const float sl2 = scale * 1.4426950408889634f;
    if (m.dt == DT_BF16) return launch_tc<__nv_bfloat16>(m, m.B, m.N, m.H, sl2, stream);
    return launch_tc<__half>(m, m.B, m.N, m.H, sl2, stream);
}

}  // namespace vidil
