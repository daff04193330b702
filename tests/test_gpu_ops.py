"""Operator-level parity on a B200, through the C ABI (vidil_op_*, vidil_sim_topk): each hand-written kernel against
plain PyTorch fp32 on the same operands (rounded to the tensor-core operand type where the kernel rounds)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import tokenization_oracle, weights as W
from vidil_b200 import _lib, ops

pytestmark = pytest.mark.gpu

TD = {"bf16": torch.bfloat16, "fp16": torch.float16}
# one ulp of the 16-bit output type at |x| ~ 8 plus accumulation-order noise
OUT_TOL = {"bf16": 4e-2, "fp16": 6e-3}


def _gemm_case(M, N, K, epi, dtype, cg, dev):
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K + epi)
    a = torch.randn(M, K, generator=g).to(dev)
    w = (torch.randn(N, K, generator=g) * K ** -0.5).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    td = TD[dtype]
    ref = a.to(td).double() @ w.to(td).double().t() + bias.double()
    out, kw = None, {}
    if epi == _lib.EPI_GELU:
        ref = F.gelu(ref)
    elif epi == _lib.EPI_QUICKGELU:
        ref = ref * torch.sigmoid(1.702 * ref)
    elif epi == _lib.EPI_RESID:
        out = torch.randn(M, N, generator=g).to(dev)
        ref = ref + out.double()
    elif epi == _lib.EPI_PATCH:
        P = 4
        frames = M // P
        pos = torch.randn(P + 1, N, generator=g).to(dev)
        out = torch.full((frames * (P + 1), N), 7.0, device=dev)
        full = torch.full((frames, P + 1, N), 7.0, device=dev, dtype=torch.float64)   # row 0 of each frame untouched
        full[:, 1:] = ref.view(frames, P, N) + pos[1:].double()
        ref = full.view(-1, N)
        kw = dict(pos=pos, patches_per_frame=P)
    got = ops.linear(a, w, bias, epilogue=epi, dtype=dtype, cta_group=cg, out=out, **kw)
    return got.double(), ref


GEMM_SHAPES = [
    (128, 256, 64, _lib.EPI_STORE_F32, "bf16"),      # one tile, one k-block
    (1, 8, 64, _lib.EPI_STORE_F32, "fp16"),          # degenerate: a single row
    (300, 768, 1024, _lib.EPI_STORE, "bf16"),        # ragged M
    (1000, 1000, 768, _lib.EPI_STORE_F32, "fp16"),   # ragged M and N (N tail < 32-wide chunk)
    (788, 4096, 1024, _lib.EPI_GELU, "bf16"),        # fc1 of ViT-L, 4 frames
    (788, 4096, 1024, _lib.EPI_QUICKGELU, "fp16"),   # CLIP fc1
    (788, 1024, 4096, _lib.EPI_RESID, "bf16"),       # fc2 + residual
    (784, 1024, 768, _lib.EPI_PATCH, "bf16"),        # patch embed scatter (+pos, CLS rows untouched)
    (6304, 3072, 1024, _lib.EPI_STORE, "fp16"),      # qkv of 32 frames: several persistent tiles per CTA
]


@pytest.mark.parametrize("cg", [1, 2])
@pytest.mark.parametrize("M,N,K,epi,dtype", GEMM_SHAPES)
def test_gemm_epilogues(cuda, M, N, K, epi, dtype, cg):
    got, ref = _gemm_case(M, N, K, epi, dtype, cg, cuda)
    assert not torch.isnan(got).any()
    err = (got - ref).abs().max().item()
    f32_out = epi in (_lib.EPI_RESID, _lib.EPI_PATCH, _lib.EPI_STORE_F32)
    tol = 2e-4 * max(1.0, K / 1024) if f32_out else OUT_TOL[dtype] * max(1.0, ref.abs().max().item() / 8)
    assert err <= tol, f"max err {err} > {tol}"


def test_gemm_rejects_bad_k(cuda):
    with pytest.raises(RuntimeError, match="multiple of 64"):
        ops.linear(torch.zeros(8, 100, device=cuda), torch.zeros(8, 100, device=cuda))


@pytest.mark.parametrize("rows,D", [(1, 128), (1000, 1024), (577, 768), (4099, 1280), (9, 256), (50432, 1024)])
def test_layernorm(cuda, rows, D):
    g = torch.Generator().manual_seed(rows + D)
    x = (torch.randn(rows, D, generator=g) * 3 + 0.5).to(cuda)
    gamma, beta = torch.randn(D, generator=g).to(cuda), torch.randn(D, generator=g).to(cuda)
    for eps in (1e-6, 1e-5):
        got = ops.layernorm(x, gamma, beta, eps)
        ref = F.layer_norm(x.double(), (D,), gamma.double(), beta.double(), eps)
        assert (got.double() - ref).abs().max().item() < 2e-5


def test_layernorm_constant_row_is_finite(cuda):
    x = torch.full((4, 1024), 3.25, device=cuda)
    got = ops.layernorm(x, torch.ones(1024, device=cuda), torch.zeros(1024, device=cuda), 1e-6)
    assert torch.isfinite(got).all() and got.abs().max().item() < 1e-3


def test_layernorm_rejects_unsupported_width(cuda):
    with pytest.raises(RuntimeError, match="unsupported width"):
        ops.layernorm(torch.zeros(2, 384, device=cuda), torch.ones(384, device=cuda), torch.zeros(384, device=cuda), 1e-6)


@pytest.mark.parametrize("B,N,H,dtype", [(1, 1, 1, "fp16"), (1, 64, 1, "bf16"), (2, 197, 16, "bf16"), (2, 197, 16, "fp16"),
                                         (1, 257, 16, "fp16"), (1, 577, 12, "bf16"), (3, 5, 2, "fp16"), (64, 197, 16, "bf16"),
                                         (2, 130, 2, "bf16"), (2, 144, 1, "fp16"), (3, 161, 2, "bf16"),
                                         (2, 209, 2, "bf16"), (3, 256, 4, "fp16"), (2, 272, 3, "bf16"), (2, 300, 2, "fp16"),
                                         (2, 258, 2, "fp16"), (3, 260, 3, "bf16"), (2, 261, 2, "fp16"), (2, 385, 2, "bf16"),
                                         (40, 257, 16, "bf16"), (1, 768, 2, "fp16"), (1, 900, 2, "bf16")])
def test_attention(cuda, B, N, H, dtype):
    g = torch.Generator().manual_seed(B * 100 + N + H)
    qkv = torch.randn(B, N, 3 * H * 64, generator=g).to(cuda)
    got = ops.attention(qkv, H, dtype=dtype)
    q, k, v = qkv.to(TD[dtype]).double().view(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
    att = ((q @ k.transpose(-2, -1)) * 64 ** -0.5).softmax(-1)
    ref = (att @ v).transpose(1, 2).reshape(B, N, H * 64)
    err = (got.double() - ref).abs()
    assert not torch.isnan(got).any()
    assert err.max().item() < (3e-2 if dtype == "bf16" else 4e-3)
    assert err.mean().item() < (2e-3 if dtype == "bf16" else 3e-4)


@pytest.mark.parametrize("B,N,H,dtype", [(3, 77, 12, "bf16"), (2, 5, 2, "fp16"), (2, 197, 4, "bf16"), (1, 130, 2, "fp16"),
                                         (65, 77, 12, "fp16"), (2, 33, 1, "bf16")])
def test_attention_causal(cuda, B, N, H, dtype):
    """The CLIP text tower's mask (query i attends to keys 0..i): per-lane different key counts inside one warp, one and two
    query tiles."""
    torch.manual_seed(N)
    qkv = torch.randn(B, N, 3 * H * 64, device=cuda)
    got = ops.attention(qkv, H, dtype=dtype, causal=True)
    q, k, v = qkv.view(B, N, 3, H, 64).permute(2, 0, 3, 1, 4).double()
    s = q @ k.transpose(-1, -2) * 0.125
    s = s.masked_fill(torch.triu(torch.ones(N, N, dtype=torch.bool, device=cuda), 1), float("-inf"))
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B, N, H * 64)
    err = (got.double() - ref).abs().max().item()
    assert err < (3e-3 if dtype == "fp16" else 2.5e-2), err


def test_attention_large_logits_are_stable(cuda):
    qkv = torch.randn(1, 197, 3 * 64, generator=torch.Generator().manual_seed(1)).to(cuda) * 30
    got = ops.attention(qkv, 1, dtype="bf16")
    assert torch.isfinite(got).all()


# ---- similarity + top-k: index work, bit-exact against the oracle (np.argsort(...)[::-1][:k]) ----------------------
@pytest.mark.parametrize("Fr,T,D,k", [(64, 1000, 768, 5), (2048, 10000, 768, 5), (1, 5, 64, 5), (7, 19965, 768, 5),
                                      (33, 365, 768, 5), (5, 16, 128, 1), (512, 10000, 768, 12)])
def test_sim_topk_indices_equal_oracle(cuda, Fr, T, D, k):
    img, bank = W.unit_rows(Fr, D, seed=Fr), W.unit_rows(T, D, seed=T + 1)
    ref_scores, ref_idx = tokenization_oracle.sim_topk(img.numpy(), bank.numpy(), k)
    scores, idx = ops.sim_topk(img.to(cuda), bank.to(cuda), k)
    assert idx.dtype == torch.int32 and tuple(idx.shape) == (Fr, k)
    assert np.array_equal(idx.cpu().numpy().astype(np.int64), ref_idx)
    assert np.abs(scores.cpu().numpy() - ref_scores).max() < 1e-5


def test_sim_topk_golden_fixture(cuda, golden_dir):
    import json
    import os
    g = json.load(open(os.path.join(golden_dir, "tokenization.json")))
    img, bank = W.unit_rows(g["F"], g["D"], seed=g["img_seed"]), W.unit_rows(g["T"], g["D"], seed=g["bank_seed"])
    _, idx = ops.sim_topk(img.to(cuda), bank.to(cuda), g["k"])
    assert idx.view(g["F"] // 8, 8, g["k"]).cpu().tolist() == g["topk_indices"]


def test_sim_topk_ties_and_duplicates(cuda):
    """Duplicate phrases score identically; argsort's tie order is unspecified, so compare as sets, and require the
    duplicates to be reported with exactly equal scores."""
    img, bank = W.unit_rows(16, 256, seed=3), W.unit_rows(300, 256, seed=4)
    bank[100] = bank[7]
    bank[200] = bank[7]
    img[0] = bank[7]
    scores, idx = ops.sim_topk(img.to(cuda), bank.to(cuda), 5)
    assert set(idx[0, :3].cpu().tolist()) == {7, 100, 200}
    assert scores[0, 0] == scores[0, 1] == scores[0, 2]
    ref_scores, ref_idx = tokenization_oracle.sim_topk(img.numpy(), bank.numpy(), 5)
    for f in range(16):
        assert set(idx[f].cpu().tolist()) == set(ref_idx[f].tolist())


def test_sim_topk_near_ties_below_fp16_resolution(cuda):
    """Phrases ~1e-5 apart in score: inside the noise of the fp16 tensor-core pass (~1e-4), well above fp32 summation-order
    noise (~1e-7, which no two fp32 matmul implementations agree on) — the fp32 re-rank must order them like the oracle."""
    D = 768
    base = W.unit_rows(1, D, seed=9)[0]
    bank = W.unit_rows(2000, D, seed=10)
    noise = W.unit_rows(8, D, seed=11)
    for j in range(8):
        v = base + (3e-3 * (j + 1)) * noise[j]
        bank[50 + 37 * j] = v / v.norm()
    img = base[None].repeat(4, 1)
    ref_scores, ref_idx = tokenization_oracle.sim_topk(img.numpy(), bank.numpy(), 8)
    gaps = -np.diff(ref_scores[0])
    assert gaps.min() > 5e-6 and (ref_scores[0][0] - ref_scores[0][-1]) < 1e-3
    _, idx = ops.sim_topk(img.to(cuda), bank.to(cuda), 8)
    assert np.array_equal(idx.cpu().numpy().astype(np.int64), ref_idx)


def test_sim_topk_combined_vg_bank_and_beyond_the_old_row_limit(cuda):
    """The four `vg` banks as ONE 44 437-phrase bank (19 965 + 16 693 + 365 + 7 414, run_visual_tokenization.py:371-383), and a
    bank larger than the 51 200-row shared-memory limit the first-generation kernel had: indices equal the fp32 argsort."""
    for Fr, T, D, k in [(256, 44437, 768, 5), (40, 70001, 64, 7)]:
        img, bank = W.unit_rows(Fr, D, seed=Fr + 1), W.unit_rows(T, D, seed=T)
        ref_scores, ref_idx = tokenization_oracle.sim_topk(img.numpy(), bank.numpy(), k)
        scores, idx = ops.sim_topk(img.to(cuda), bank.to(cuda), k)
        assert np.array_equal(idx.cpu().numpy().astype(np.int64), ref_idx), (Fr, T)
        assert np.abs(scores.cpu().numpy() - ref_scores).max() < 1e-5


def test_sim_topk_clusters_of_near_synonyms(cuda):
    """Real ontology banks contain clusters of near-duplicate phrases.  40 phrases within ~1e-4 of one another in score — far
    more than any fixed candidate count, and below the resolution of the fp16 tensor-core pass — all inside ONE 32-phrase
    group or spread over many: the selection keeps re-scoring until the fp32 ranking is certain (the margin guard)."""
    D, k = 768, 10
    base = W.unit_rows(1, D, seed=21)[0]
    noise = W.unit_rows(64, D, seed=22)
    for layout in ("one_group", "spread"):
        bank = W.unit_rows(4000, D, seed=23)
        for j in range(40):
            v = base + (2e-3 * (j + 1)) * noise[j]                      # consecutive scores ~1e-5 apart: above fp32 summation noise
            row = 1024 + j if layout == "one_group" else 37 + 97 * j      # one_group: columns 1024..1063 = groups 32 and 33
            bank[row] = v / v.norm()
        img = torch.stack([base, base + 1e-3 * noise[50], W.unit_rows(1, D, seed=24)[0]])
        img = img / img.norm(dim=1, keepdim=True)
        ref_scores, ref_idx = tokenization_oracle.sim_topk(img.numpy(), bank.numpy(), k)
        assert (ref_scores[0][0] - ref_scores[0][-1]) < 1e-3 and (-np.diff(ref_scores[0])).min() > 2e-6   # a cluster, but rankable
        scores, idx = ops.sim_topk(img.to(cuda), bank.to(cuda), k)
        assert np.array_equal(idx.cpu().numpy().astype(np.int64), ref_idx), layout


def test_sim_topk_best_entries_owned_by_one_lane(cuda):
    """The selection kernel finds its threshold from four sorted heads per lane and falls back to a destructive extraction
    when one lane owns more of the k best than that (csrc/topk.cu, step 1).  A lane's share of the tagged-score pool is
    entry (lane + slice) mod 32 of every 32-entry slice, a group's best score is pool entry 4 * group: the best phrases of
    groups 0, 33, 66, 99, 132 and 165 all belong to lane 0 — the fallback must give the oracle's ranking — and with the best
    phrases in groups 0..5 (six different lanes) the fast path must."""
    D, k = 768, 8
    for groups in ([0, 33, 66, 99, 132, 165], [0, 1, 2, 3, 4, 5]):
        bank = W.unit_rows(6000, D, seed=31)
        cols = [g * 32 + 7 for g in groups]
        img = sum((1.0 + 0.05 * j) * bank[c] for j, c in enumerate(cols)) + 0.5 * W.unit_rows(1, D, seed=32)[0]
        img = torch.stack([img / img.norm(), W.unit_rows(1, D, seed=33)[0]])
        ref_scores, ref_idx = tokenization_oracle.sim_topk(img.numpy(), bank.numpy(), k)
        assert set(ref_idx[0][:6].tolist()) == set(cols)
        scores, idx = ops.sim_topk(img.to(cuda), bank.to(cuda), k)
        assert np.array_equal(idx.cpu().numpy().astype(np.int64), ref_idx), groups
        assert np.abs(scores.cpu().numpy() - ref_scores).max() < 1e-5


def test_sim_topk_cached_bank_equals_one_shot_and_follows_updates(cuda):
    """ops.sim_topk prepares a bank once per tensor (SimBank); the one-shot ABI call converts it inside the call.  Same
    result; an in-place update of the bank tensor is seen (the cache keys on the tensor's version)."""
    img, bank = W.unit_rows(33, 256, seed=1).to(cuda), W.unit_rows(777, 256, seed=2).to(cuda)
    a = ops.sim_topk(img, bank, 5)
    b = ops.sim_topk(img, bank, 5, cache_bank=False)
    c = ops.sim_topk(img, bank, 5)
    assert torch.equal(a[1], b[1]) and torch.equal(a[0], b[0]) and torch.equal(a[1], c[1])
    bank[5] = img[0]
    d = ops.sim_topk(img, bank, 5)
    assert int(d[1][0, 0]) == 5 and abs(float(d[0][0, 0]) - 1.0) < 1e-5
    # non-unit rows: the error bound scales with the norms
    img2, bank2 = img * 7.5, bank * 0.01
    ref_scores, ref_idx = tokenization_oracle.sim_topk(img2.cpu().numpy(), bank2.cpu().numpy(), 5)
    _, idx = ops.sim_topk(img2, bank2, 5)
    assert np.array_equal(idx.cpu().numpy().astype(np.int64), ref_idx)


def test_sim_topk_errors(cuda):
    img, bank = W.unit_rows(4, 64, seed=0).to(cuda), W.unit_rows(10, 64, seed=1).to(cuda)
    with pytest.raises(RuntimeError):
        ops.sim_topk(img, bank, 13)           # k > 12
    with pytest.raises(RuntimeError):
        ops.sim_topk(img, bank[:3], 5)        # k > T
    s, i = ops.sim_topk(img[:0], bank, 5)     # empty frame list
    assert tuple(s.shape) == (0, 5) and tuple(i.shape) == (0, 5)
