"""The med.py text stack on a B200, through the C ABI (vidil_med_forward / vidil_med_generate / vidil_op_beam_search)
behind the drop-in modules, against the CPU oracle and the fixtures made by the reference's own med.py."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import med_oracle, weights as W
from vidil_b200 import _lib
from vidil_b200.med import BertConfig, BertLMHeadModel, BertModel

pytestmark = pytest.mark.gpu

# Tolerances on logits, as a fraction of the logits' standard deviation (1.2 for the tiny model, 2.8 for BERT-base; |logit| up
# to ~12): max / mean absolute error.  Measured on B200: fp16 9.5e-3 / 2.3e-3, bf16 6.7e-2 / 1.3e-2 at BERT-base depth — the
# post-LayerNorm stack re-normalises after each of its 36 sub-layers, so operand rounding is not diluted by a growing residual
# as in the pre-LN ViT; 16-bit operand rounding (2^-11 fp16, 2^-8 bf16) is the whole error, accumulation is fp32.
LOGIT_MAX = {"fp16": 2e-2, "bf16": 1.2e-1}
LOGIT_MEAN = {"fp16": 5e-3, "bf16": 2.5e-2}
LOGIT_TOL = {k: v * 1.2 for k, v in LOGIT_MAX.items()}   # tiny model, absolute


def _cfg(name):
    return BertConfig(**W.MED_CONFIGS[name])


def _decoder(name, dtype, dev, seed=0):
    sd = W.med_state_dict(name, "decoder", seed=seed)
    m = BertLMHeadModel(_cfg(name), compute_dtype=dtype)
    missing, unexpected = m.load_state_dict({k[len("text_decoder."):]: v for k, v in sd.items()}, strict=False)
    assert not unexpected and all("position_ids" in k for k in missing)
    return m.to(dev).eval(), sd


def _itm(name, dtype, dev):
    sd = W.med_state_dict(name, "itm", seed=0)
    m = BertModel(_cfg(name), compute_dtype=dtype)
    missing, unexpected = m.load_state_dict({k[len("text_encoder."):]: v for k, v in sd.items() if k.startswith("text_encoder.")},
                                            strict=False)
    assert not unexpected and all("position_ids" in k for k in missing)
    head = torch.nn.Linear(W.MED_CONFIGS[name]["hidden_size"], 2)
    head.load_state_dict({"weight": sd["itm_head.weight"], "bias": sd["itm_head.bias"]})
    m, head = m.to(dev).eval(), head.to(dev)
    m.attach_cls_head(head)
    return m, sd


# ---- beam search bookkeeping alone: bit-exact tokens against the restated v4.15 rules ----------------------------------
def _op_beam_search(dev, L, F, K, V, prompt, max_length, min_length, eos, pad=0, lp=1.0):
    lib = _lib.load()
    S = len(L)
    logits = torch.from_numpy(np.stack(L)).to(dev).contiguous()
    need = lib.vidil_op_beam_search_workspace_bytes(F, K, max_length)
    buf = torch.empty(need + 1024, dtype=torch.uint8, device=dev)
    off = (-buf.data_ptr()) % 1024
    toks = torch.empty(F, max_length, dtype=torch.int32, device=dev)
    lens = torch.empty(F, dtype=torch.int32, device=dev)
    scores = torch.empty(F, dtype=torch.float32, device=dev)
    p = (ctypes.c_int32 * len(prompt))(*prompt)
    st = lib.vidil_op_beam_search(logits.data_ptr(), S, F, K, V, ctypes.cast(p, ctypes.c_void_p), len(prompt), max_length, min_length,
                                  eos, pad, lp, toks.data_ptr(), lens.data_ptr(), scores.data_ptr(), buf.data_ptr() + off, need,
                                  torch.cuda.current_stream().cuda_stream)
    _lib.check(st, "vidil_op_beam_search")
    torch.cuda.synchronize()
    return [toks[b, :int(lens[b])].tolist() for b in range(F)], scores.cpu().numpy()


@pytest.mark.parametrize("K,V,eos_boost,quantise", [(3, 200, 3.0, False), (3, 30524, 6.0, False), (1, 64, 2.0, False),
                                                    (4, 400, 4.0, False), (2, 120, 1.0, False), (3, 200, 3.0, True)])
def test_beam_search_op_matches_oracle(cuda, K, V, eos_boost, quantise):
    F, prompt, max_length, min_length, eos = 7, [V - 2, 5, 9, 11], 14, 6, 2
    rng = np.random.default_rng(K * 1000 + V)
    S = max_length - len(prompt)
    L = []
    for s in range(S):
        x = rng.standard_normal((F * K, V)).astype(np.float32) * 2.0
        x[:, eos] += eos_boost * rng.random((F * K,)).astype(np.float32) * 2
        if quantise:
            x = np.round(x * 2) / 2        # many exact ties inside a row: lowest token index must win
        L.append(x)
    it = iter(L)
    ref_toks, ref_scores, trace = med_oracle.beam_search_from_logits(lambda ids, bi: next(it), F, prompt, K, max_length, min_length,
                                                                     eos, 0)
    toks, scores = _op_beam_search(cuda, L, F, K, V, prompt, max_length, min_length, eos)
    if eos_boost >= 2.0:
        assert any(t[-1] == eos for t in ref_toks), "the case should exercise finished hypotheses"
    if quantise:   # equal-score alternatives across beams may legitimately resolve differently by one ulp of the lse
        assert all(t == r or abs(s - rs) < 1e-5 for t, r, s, rs in zip(toks, ref_toks, scores, ref_scores))
    else:
        assert toks == ref_toks
    assert np.allclose(scores, np.asarray(ref_scores, dtype=np.float32), atol=2e-5)


def test_beam_search_op_known_answer(cuda):
    """The hand-worked case (b) of tests/test_med_oracle.py."""
    K, V, eos = 2, 8, 1
    row = np.full(V, -30.0, dtype=np.float32)
    row[:6] = np.log(np.array([0.01, 0.60, 0.30, 0.05, 0.02, 0.02], dtype=np.float32))
    L = [np.tile(row, (K, 1)) for _ in range(7)]
    toks, scores = _op_beam_search(cuda, L, 1, K, V, [5], 8, 3, eos)
    assert toks[0] == [5, 2, 2, eos]
    assert abs(scores[0] - (2 * np.log(0.3) + np.log(0.6)) / 3) < 1e-4


# ---- nucleus sampling: the processors, warpers and the draw alone, on given logits and given uniform numbers ---------------
def _op_sample(dev, L, F, V, prompt, uniforms, max_length, min_length, eos, pad=0, top_k=50, top_p=0.9, rep=1.1):
    lib = _lib.load()
    S = len(L)
    logits = torch.from_numpy(np.stack(L)).to(dev).contiguous()
    uni = torch.from_numpy(np.ascontiguousarray(uniforms, dtype=np.float32)).to(dev)
    need = lib.vidil_op_beam_search_workspace_bytes(F, 1, max_length)
    buf = torch.empty(need + 1024, dtype=torch.uint8, device=dev)
    off = (-buf.data_ptr()) % 1024
    toks = torch.empty(F, max_length, dtype=torch.int32, device=dev)
    lens = torch.empty(F, dtype=torch.int32, device=dev)
    scores = torch.empty(F, dtype=torch.float32, device=dev)
    p = (ctypes.c_int32 * len(prompt))(*prompt)
    st = lib.vidil_op_sample(logits.data_ptr(), S, F, V, ctypes.cast(p, ctypes.c_void_p), len(prompt), max_length, min_length, eos, pad,
                             top_k, top_p, rep, uni.data_ptr(), toks.data_ptr(), lens.data_ptr(), scores.data_ptr(), buf.data_ptr() + off,
                             need, torch.cuda.current_stream().cuda_stream)
    _lib.check(st, "vidil_op_sample")
    torch.cuda.synchronize()
    return [toks[b, :int(lens[b])].tolist() for b in range(F)], scores.cpu().numpy(), toks.cpu().numpy()


@pytest.mark.parametrize("V,top_k,top_p,scale,eos_boost,quantise", [(200, 50, 0.9, 2.0, 3.0, False), (30524, 50, 0.9, 1.0, 6.0, False),
                                                                    (30524, 50, 0.9, 4.0, 9.0, False), (64, 5, 0.5, 2.0, 2.0, False),
                                                                    (400, 50, 0.99, 0.5, 1.0, False), (200, 50, 0.9, 2.0, 3.0, True),
                                                                    (2048, 1000, 0.999, 0.01, 0.0, False)])
def test_sample_op_matches_oracle(cuda, V, top_k, top_p, scale, eos_boost, quantise):
    """Same logits, same uniform numbers: the drawn tokens are those of the oracle (med_oracle.sample_from_logits, whose
    processors are pinned against transformers' own classes).  Covers the real vocabulary, a flat row with more than 256
    survivors, quantised rows full of exact ties, an early and a late eos."""
    F, prompt, max_length, min_length, eos, pad = 9, [V - 2, 5, 9, 11], 16, 7, 2, 0
    rng = np.random.default_rng(V + top_k)
    S = max_length - len(prompt)
    L = []
    for s in range(S):
        x = rng.standard_normal((F, V)).astype(np.float32) * scale
        x[:, eos] += eos_boost * rng.random((F,)).astype(np.float32) * 2
        if quantise:
            x = np.round(x * 2) / 2
        L.append(x)
    uniforms = rng.random((S, F)).astype(np.float32)
    it = iter(L)
    ref, ref_logp = med_oracle.sample_from_logits(lambda ids, _: next(it), F, prompt, uniforms, max_length, min_length, eos, pad, top_k,
                                                  top_p, 1.1)
    toks, logp, raw = _op_sample(cuda, L, F, V, prompt, uniforms, max_length, min_length, eos, pad, top_k, top_p, 1.1)
    same = sum(t == r for t, r in zip(toks, ref))
    print(f"sample op V={V} top_k={top_k}: {same}/{F} sequences identical")
    # one draw in ~10^5 can land within fp32 rounding of a boundary of the cumulative distribution (expf vs numpy's exp)
    assert same >= F - 1
    assert all(eos not in t[len(prompt):min_length] for t in toks)
    for b, t in enumerate(toks):
        assert (raw[b, len(t):] == pad).all()
        if t == ref[b]:
            assert abs(logp[b] - ref_logp[b]) < 1e-3
    if eos_boost >= 3.0:
        assert any(t[-1] == eos and len(t) < max_length for t in toks), "the case should exercise finished frames"


def test_sample_op_zero_uniform_is_greedy_and_draws_follow_the_distribution(cuda):
    """u = 0 takes the most probable surviving token at every step; and over many frames with the SAME logits the drawn
    tokens follow the renormalised top-k / top-p distribution (chi-square against the oracle's probabilities)."""
    V, eos, pad, prompt = 200, 2, 0, [7, 9]
    rng = np.random.default_rng(11)
    row = (rng.standard_normal(V) * 1.5).astype(np.float32)
    n = 4096
    L = [np.tile(row, (n, 1))]
    toks, _, _ = _op_sample(cuda, L, n, V, prompt, np.zeros((1, n), np.float32), 3, 0, eos, pad)
    kept, v = med_oracle.process_sampling_scores(row, np.array(prompt), 2, 0, eos, 50, 0.9, 1.1)
    assert all(t[2] == kept[0] for t in toks)
    toks, _, _ = _op_sample(cuda, L, n, V, prompt, rng.random((1, n)).astype(np.float32), 3, 0, eos, pad)
    drawn = np.array([t[2] for t in toks])
    assert set(drawn.tolist()) <= set(kept.tolist())
    p = np.exp(v - v[0])
    p /= p.sum()
    counts = np.array([(drawn == k).sum() for k in kept])
    assert (((counts - n * p) ** 2) / (n * p)).sum() < 3 * len(kept) + 20


# ---- network arithmetic ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", ["fp16", "bf16"])
def test_decoder_logits_tiny_vs_fixture_and_oracle(cuda, golden_dir, dtype):
    name, batch, T, n_img = "tiny", 3, 9, 5
    g = np.load(os.path.join(golden_dir, "med_tiny.npz"))
    m, sd = _decoder(name, dtype, cuda)
    c, sp = W.MED_CONFIGS[name], W.MED_SPECIAL[name]
    enc = W.image_tokens(batch, n_img, c["encoder_width"], seed=0)
    ids, _ = W.caption_ids(name, batch, T, seed=0, min_words=T - 2)
    ids[:, 0] = sp["bos"]
    out = m(ids, encoder_hidden_states=enc.to(cuda)).logits.cpu().numpy()
    err = np.abs(out[:, :, g["vocab"]] - g["logits"]).max()
    print(f"med tiny {dtype}: logits max-abs err {err:.3e} (std {float(g['logits_std']):.2f})")
    assert err < LOGIT_TOL[dtype]
    with torch.no_grad():
        ref, _ = med_oracle.decoder_logits(sd, "text_decoder.", ids, enc, c["num_attention_heads"], c["num_hidden_layers"])
    assert np.abs(out - ref.numpy()).max() < LOGIT_TOL[dtype]


@pytest.mark.parametrize("dtype", ["fp16", "bf16"])
def test_decoder_logits_base_vs_reference_fixture(cuda, golden_dir, dtype):
    """BERT-base decoder on ViT-L tokens (197 x 1024) against the reference's med.py output."""
    name, batch, T, n_img = "base_l", 2, 8, 197
    g = np.load(os.path.join(golden_dir, "med_base_l.npz"))
    m, _ = _decoder(name, dtype, cuda)
    c, sp = W.MED_CONFIGS[name], W.MED_SPECIAL[name]
    enc = W.image_tokens(batch, n_img, c["encoder_width"], seed=0)
    ids, _ = W.caption_ids(name, batch, T, seed=0, min_words=T - 2)
    ids[:, 0] = sp["bos"]
    out = m(ids, encoder_hidden_states=enc.to(cuda)).logits.cpu().numpy()
    err = np.abs(out[:, :, g["vocab"]] - g["logits"])
    print(f"med base_l {dtype}: logits max-abs err {err.max():.3e} mean {err.mean():.3e} (std {float(g['logits_std']):.2f})")
    std = float(g["logits_std"])
    assert err.max() < LOGIT_MAX[dtype] * std and err.mean() < LOGIT_MEAN[dtype] * std
    # the top token of every position agrees with the reference wherever the reference's margin is clear
    ref = g["logits"]
    top2 = np.sort(ref, axis=-1)[..., -2:]
    clear = (top2[..., 1] - top2[..., 0]) > 2 * LOGIT_MAX[dtype] * std
    assert (out[:, :, g["vocab"]].argmax(-1) == ref.argmax(-1))[clear].all()


@pytest.mark.parametrize("name,n_img,dtype", [("tiny", 5, "fp16"), ("tiny", 5, "bf16"), ("base_l", 197, "bf16")])
def test_itm_logits_vs_fixture(cuda, golden_dir, name, n_img, dtype):
    batch, T = (3, 9) if name == "tiny" else (2, 8)
    g = np.load(os.path.join(golden_dir, f"med_{name}.npz"))
    m, sd = _itm(name, dtype, cuda)
    c, sp = W.MED_CONFIGS[name], W.MED_SPECIAL[name]
    enc = W.image_tokens(batch, n_img, c["encoder_width"], seed=0)
    cap, mask = W.caption_ids(name, batch, T, seed=1)
    cap[:, 0] = sp["enc"]
    hidden, _, cls = m.run(cap, mask, enc.to(cuda), causal=False, want_hidden=True, want_cls=True)
    err_h = np.abs(hidden[:, 0].cpu().numpy() - g["itm_hidden_cls"]).max()
    err = np.abs(cls.cpu().numpy() - g["itm_logits"]).max()
    print(f"itm {name} {dtype}: cls hidden err {err_h:.3e}, itm logits err {err:.3e}")
    assert err_h < (1e-2 if dtype == "fp16" else 1e-1)      # max over the 768 unit-scale components of the [ENC] row
    assert err < (1e-2 if dtype == "fp16" else 8e-2)
    # module call surface of blip_itm.py:49-56
    out = m(cap, attention_mask=mask, encoder_hidden_states=enc.to(cuda), return_dict=True)
    assert torch.equal(out.last_hidden_state, hidden)


def test_itm_pairs_share_frames(cuda):
    """frame_of_seq: every (caption, frame) pair of a video in one call equals one call per caption (run_video_CapFilt.py:108-112)."""
    name = "tiny"
    m, sd = _itm(name, "bf16", cuda)
    c, sp = W.MED_CONFIGS[name], W.MED_SPECIAL[name]
    n_frames, n_caps, T = 4, 3, 12
    enc = W.image_tokens(n_frames, 7, c["encoder_width"], seed=5).to(cuda)
    cap, mask = W.caption_ids(name, n_caps, T, seed=2)
    cap[:, 0] = sp["enc"]
    pair_ids = cap.repeat_interleave(n_frames, 0)
    pair_mask = mask.repeat_interleave(n_frames, 0)
    frame_of = torch.arange(n_frames).repeat(n_caps)
    _, _, all_pairs = m.run(pair_ids, pair_mask, enc, frame_of_seq=frame_of, want_hidden=False, want_cls=True)
    for i in range(n_caps):
        _, _, one = m.run(cap[i:i + 1].repeat(n_frames, 1), mask[i:i + 1].repeat(n_frames, 1), enc, want_hidden=False, want_cls=True)
        assert torch.equal(one, all_pairs[i * n_frames:(i + 1) * n_frames])
    with torch.no_grad():
        ref = med_oracle.itm_logits(sd, enc.cpu()[frame_of], pair_ids, pair_mask, c["num_attention_heads"], c["num_hidden_layers"])
    assert (all_pairs.cpu() - ref).abs().max() < 8e-2
    # frame-major pairs with seqs_per_frame: the captions of a frame share one cross-attention query group (what
    # vidil_b200.capfilt.filter_captions sends); same numbers, re-ordered
    _, _, grouped = m.run(cap.repeat(n_frames, 1), mask.repeat(n_frames, 1), enc, want_hidden=False, want_cls=True, seqs_per_frame=n_caps)
    assert torch.equal(grouped.view(n_frames, n_caps, 2).transpose(0, 1).reshape(-1, 2), all_pairs)


# ---- generation ----------------------------------------------------------------------------------------------------------
_TORCH_DT = {"fp16": torch.float16, "bf16": torch.bfloat16}


def _generate_case(cuda, name, dtype, F, n_img, max_length, min_length, seed, emulate=False, margins=None):
    """emulate: compare with the oracle that rounds 16-bit operands / stored tensors where the kernels do (med_oracle.emulate)
    instead of the plain fp32 restatement.  margins: list filled with the oracle's per-frame decision margin."""
    m, sd = _decoder(name, dtype, cuda)
    c, sp = W.MED_CONFIGS[name], W.MED_SPECIAL[name]
    enc = W.image_tokens(F, n_img, c["encoder_width"], seed=seed)
    prompt = torch.tensor([sp["prompt"]], dtype=torch.long).repeat(F, 1)
    out, scores, lens = m.generate(input_ids=prompt, max_length=max_length, min_length=min_length, num_beams=3,
                                   eos_token_id=sp["eos"], pad_token_id=sp["pad"], encoder_hidden_states=enc.to(cuda),
                                   return_scores=True)
    ref_toks, ref_scores, _ = med_oracle.generate(sd, enc, sp["prompt"], c["num_attention_heads"], c["num_hidden_layers"],
                                                  num_beams=3, max_length=max_length, min_length=min_length, eos=sp["eos"],
                                                  pad=sp["pad"], operand_dtype=_TORCH_DT[dtype] if emulate else None,
                                                  margins=margins)
    got = [out[b, :int(lens[b])].tolist() for b in range(F)]
    return got, scores.cpu().numpy(), ref_toks, np.asarray(ref_scores, dtype=np.float32), out, sp


@pytest.mark.parametrize("dtype", ["fp16", "bf16"])
def test_generate_tiny_vs_oracle(cuda, dtype):
    got, scores, ref, ref_scores, out, sp = _generate_case(cuda, "tiny", dtype, F=24, n_img=5, max_length=14, min_length=5, seed=3)
    same = sum(g == r for g, r in zip(got, ref))
    print(f"generate tiny {dtype}: {same}/24 captions identical; score err {np.abs(scores - ref_scores).max():.3e}")
    # 16-bit operands move logits by ~1e-2: a near-tie between two beams can legitimately pick the other branch.  Where the
    # caption differs its score must still be as good as the oracle's to within that noise.
    assert same >= (22 if dtype == "fp16" else 18)
    tol = 0.02 if dtype == "fp16" else 0.08
    assert all(g == r or s > rs - tol for g, r, s, rs in zip(got, ref, scores, ref_scores))
    assert np.abs(scores - ref_scores)[[g == r for g, r in zip(got, ref)]].max() < tol
    assert all(g[:4] == sp["prompt"] for g in got)
    assert out.dtype == torch.int64 and out.shape[1] == max(len(g) for g in got)


# Token equality against the operand-emulating oracle.  The emulation rounds operands where the kernels do, but it cannot make
# the two paths bit-identical: their fp32 accumulation orders differ by ~1e-7, which now and then rounds an intermediate to the
# neighbouring 16-bit value (2^-8 relative in bf16, 2^-11 in fp16); twelve layers later a candidate's log-probability has moved
# by ~1e-2 (bf16) — mostly in common with its neighbours, which is why captions agree far more often than that figure suggests.
# Random-initialised decoders have near-uniform next-token distributions, so some searches are decided by less than what is left
# between neighbours.  The oracle therefore reports, per frame, the smallest gap by which any comparison that shaped the result
# was decided (`margins`, med_oracle.beam_search_from_logits), and the assertion is hard wherever it can be: a frame decided by
# more than _DECIDED MUST come out token-identical; a frame decided by less may follow the other branch and must then score
# within the operand noise of the oracle's choice.  _DECIDED is 1.5x the largest margin at which a caption was ever seen to
# differ on the B200 (fp16 3.3e-4, bf16 2.6e-3 over the 66 frame-cases below; the per-frame table is printed with -s).
_DECIDED = {"fp16": 1.0e-3, "bf16": 4.0e-3}        # sum-of-log-probability units
_SAME_SCORE = {("tiny", "fp16"): 1e-3, ("tiny", "bf16"): 3e-3, ("base", "fp16"): 1e-2, ("base", "bf16"): 0.1}   # per token
_OTHER_BRANCH = {"fp16": 0.02, "bf16": 0.1}        # per token, as in the fp32-oracle tests


def _assert_equal_where_decided(tag, dtype, got, scores, ref, ref_scores, margins, min_decided):
    decided = [mg > _DECIDED[dtype] for mg in margins]
    for b, (g, r) in enumerate(zip(got, ref)):
        print(f"  {tag} {dtype} frame {b}: margin {margins[b]:.2e} {'decided ' if decided[b] else 'near-tie'} "
              f"{'identical' if g == r else 'DIFFERENT'} score {scores[b]:.5f} vs {ref_scores[b]:.5f}")
    n_same = sum(g == r for g, r in zip(got, ref))
    print(f"generate {tag} {dtype} vs emulating oracle: {n_same}/{len(ref)} identical, {sum(decided)} decided by more than "
          f"{_DECIDED[dtype]:.0e} — all of those identical")
    same_tol = _SAME_SCORE[("tiny" if tag == "tiny" else "base", dtype)]
    for b, (g, r) in enumerate(zip(got, ref)):
        if decided[b]:
            assert g == r, f"frame {b}: decided by {margins[b]:.3e} but the tokens differ"
        if g == r:
            assert abs(scores[b] - ref_scores[b]) < same_tol, f"frame {b}: same tokens, score off by {abs(scores[b] - ref_scores[b]):.3e}"
        else:
            assert scores[b] > ref_scores[b] - _OTHER_BRANCH[dtype], f"frame {b}: follows a branch that scores worse than the noise allows"
    assert sum(decided) >= min_decided
    assert n_same >= len(ref) - max(1, len(ref) // 8)


@pytest.mark.parametrize("dtype,min_decided", [("fp16", 20), ("bf16", 12)])
def test_generate_tiny_tokens_equal_the_operand_emulating_oracle(cuda, dtype, min_decided):
    margins = []
    got, scores, ref, ref_scores, out, sp = _generate_case(cuda, "tiny", dtype, F=24, n_img=5, max_length=14, min_length=5, seed=3,
                                                           emulate=True, margins=margins)
    _assert_equal_where_decided("tiny", dtype, got, scores, ref, ref_scores, margins, min_decided)


@pytest.mark.parametrize("name,frames,n_img,dtype,min_decided", [("base_l", 6, 197, "bf16", 2), ("base_b", 3, 577, "bf16", 1),
                                                                 ("base_l", 6, 197, "fp16", 3), ("base_b", 3, 577, "fp16", 1)])
def test_generate_base_tokens_equal_the_operand_emulating_oracle(cuda, name, frames, n_img, dtype, min_decided):
    """BLIP's real decoder shapes (BERT-base, 30 524-token vocabulary; 197 ViT-L/16@224 tokens or 577 ViT-B/16@384 tokens per
    frame), reference call-site arguments (run_video_CapFilt.py:102: beams 3, max_length 20, min_length 5).  Measured: fp16
    6/6 and 3/3 identical, bf16 5/6 and 2/3 (the two that differ were decided by 1.6e-3 and 2.6e-3)."""
    margins = []
    got, scores, ref, ref_scores, out, sp = _generate_case(cuda, name, dtype, F=frames, n_img=n_img, max_length=20, min_length=5,
                                                           seed=1, emulate=True, margins=margins)
    _assert_equal_where_decided(name, dtype, got, scores, ref, ref_scores, margins, min_decided)


def test_generate_base_vs_fp32_oracle_statistic(cuda):
    """The same against the plain fp32 restatement (the reported statistic): bf16 operands move BERT-base logits by up to ~0.2
    (test_decoder_logits_base_vs_reference_fixture), so where two continuations are closer than that the search may follow the
    other one; a differing caption must still score within that noise of the fp32 oracle's."""
    got, scores, ref, ref_scores, out, sp = _generate_case(cuda, "base_l", "bf16", F=6, n_img=197, max_length=20, min_length=5, seed=1)
    same = sum(g == r for g, r in zip(got, ref))
    print(f"generate base_l bf16 vs fp32 oracle: {same}/6 captions identical; scores {scores} vs {ref_scores}")
    assert same >= 3
    assert all(g == r or s > rs - 0.1 for g, r, s, rs in zip(got, ref, scores, ref_scores))


def test_generate_is_batch_invariant_and_accepts_expanded_tokens(cuda):
    """A frame's caption does not depend on what else is in the batch, and the reference's repeat_interleaved image tokens
    (blip.py:130) are accepted."""
    name = "tiny"
    m, _ = _decoder(name, "bf16", cuda)
    c, sp = W.MED_CONFIGS[name], W.MED_SPECIAL[name]
    enc = W.image_tokens(9, 6, c["encoder_width"], seed=7).to(cuda)
    kw = dict(max_length=12, min_length=5, num_beams=3, eos_token_id=sp["eos"], pad_token_id=sp["pad"], return_scores=True)
    p = lambda n: torch.tensor([sp["prompt"]], dtype=torch.long).repeat(n, 1)  # noqa: E731
    full, fs, fl = m.generate(input_ids=p(9), encoder_hidden_states=enc, **kw)
    part, ps, pl = m.generate(input_ids=p(4), encoder_hidden_states=enc[2:6], **kw)
    for i in range(4):
        assert full[2 + i, :int(fl[2 + i])].tolist() == part[i, :int(pl[i])].tolist()
    exp, es, el = m.generate(input_ids=p(9), encoder_hidden_states=enc.repeat_interleave(3, dim=0), **kw)
    assert torch.equal(exp, full) and torch.equal(es, fs)


def test_med_errors_are_loud(cuda):
    m, _ = _decoder("tiny", "bf16", cuda)
    sp = W.MED_SPECIAL["tiny"]
    enc = W.image_tokens(2, 5, 128, seed=0)
    with pytest.raises(RuntimeError):          # CPU tensors: no fallback
        m(torch.tensor([sp["prompt"]] * 2), encoder_hidden_states=enc)
    with pytest.raises(NotImplementedError):   # sampling is built as blip.py calls it: one beam
        m.generate(input_ids=torch.tensor([sp["prompt"]] * 2), do_sample=True, num_beams=3, eos_token_id=sp["eos"],
                   encoder_hidden_states=enc.to(cuda))
    with pytest.raises(NotImplementedError):   # the beam-search path of the reference never sets a repetition penalty
        m.generate(input_ids=torch.tensor([sp["prompt"]] * 2), num_beams=3, repetition_penalty=1.1, eos_token_id=sp["eos"],
                   encoder_hidden_states=enc.to(cuda))
    with pytest.raises(RuntimeError):          # top_k beyond what the sampling kernel's candidate list holds
        m.generate(input_ids=torch.tensor([sp["prompt"]] * 2), do_sample=True, top_k=5000, eos_token_id=sp["eos"],
                   encoder_hidden_states=enc.to(cuda))
    with pytest.raises(RuntimeError):          # max_length beyond the 64-token limit of the search state
        m.generate(input_ids=torch.tensor([sp["prompt"]] * 2), max_length=100, num_beams=3, eos_token_id=sp["eos"],
                   encoder_hidden_states=enc.to(cuda))


def test_itm_padding_columns_do_not_matter(cuda):
    """blip_itm.py:46 pads every caption to 35 tokens; the [ENC]-row logits are the same whether the all-padding columns are
    run (hidden requested) or dropped (cls only) — up to the different GEMM row count's tile rounding, i.e. exactly."""
    name = "tiny"
    m, sd = _itm(name, "bf16", cuda)
    c, sp = W.MED_CONFIGS[name], W.MED_SPECIAL[name]
    enc = W.image_tokens(3, 7, c["encoder_width"], seed=9).to(cuda)
    cap, mask = W.caption_ids(name, 3, 12, seed=4)
    cap[:, 0] = sp["enc"]
    pad_ids = torch.cat([cap, torch.zeros(3, 23, dtype=torch.long)], 1)
    pad_mask = torch.cat([mask, torch.zeros(3, 23, dtype=torch.long)], 1)
    hidden, _, full = m.run(pad_ids, pad_mask, enc, want_hidden=True, want_cls=True)       # all 35 columns
    _, _, trimmed = m.run(pad_ids, pad_mask, enc, want_hidden=False, want_cls=True)        # padding columns dropped
    assert hidden.shape[1] == 35
    assert torch.equal(full, trimmed)
    with torch.no_grad():
        ref = med_oracle.itm_logits(sd, enc.cpu(), pad_ids, pad_mask, c["num_attention_heads"], c["num_hidden_layers"])
    assert (trimmed.cpu() - ref).abs().max() < 8e-2


def test_generate_and_itm_on_vit_b_384_tokens(cuda):
    """The shipped pipeline config (image_size 384, vit 'base': 577 tokens of width 768 per frame): the decode cross-attention
    takes its K/V tiles in three TMA boxes, the ITM cross-attention runs 10 key blocks."""
    name = "base_b"
    c, sp = W.MED_CONFIGS[name], W.MED_SPECIAL[name]
    H, depth = c["num_attention_heads"], c["num_hidden_layers"]
    enc = W.image_tokens(3, 577, c["encoder_width"], seed=11)
    m, sd = _decoder(name, "bf16", cuda)
    prompt = torch.tensor([sp["prompt"]], dtype=torch.long).repeat(3, 1)
    out, scores, lens = m.generate(input_ids=prompt, max_length=12, min_length=5, num_beams=3, eos_token_id=sp["eos"],
                                   pad_token_id=sp["pad"], encoder_hidden_states=enc.to(cuda), return_scores=True)
    ref_toks, ref_scores, _ = med_oracle.generate(sd, enc, sp["prompt"], H, depth, num_beams=3, max_length=12, min_length=5,
                                                  eos=sp["eos"], pad=sp["pad"])
    got = [out[b, :int(lens[b])].tolist() for b in range(3)]
    print(f"generate base_b/577: {sum(g == r for g, r in zip(got, ref_toks))}/3 identical; scores {scores.cpu().numpy()} vs {ref_scores}")
    assert sum(g == r for g, r in zip(got, ref_toks)) >= 2
    assert all(g == r or s > rs - 0.1 for g, r, s, rs in zip(got, ref_toks, scores.cpu().numpy(), ref_scores))
    # teacher-forced logits of the generated sequences agree with the oracle on the same ids
    ids = out[:, :8].cpu()
    logits = m(ids, encoder_hidden_states=enc.to(cuda)).logits.cpu()
    with torch.no_grad():
        ref_logits, _ = med_oracle.decoder_logits(sd, "text_decoder.", ids, enc, H, depth)
    err = (logits - ref_logits).abs()
    std = float(ref_logits.std())
    assert err.max() < LOGIT_MAX["bf16"] * std and err.mean() < LOGIT_MEAN["bf16"] * std
    del m
    mi, sdi = _itm(name, "bf16", cuda)
    cap, mask = W.caption_ids(name, 3, 35, seed=6)
    cap[:, 0] = sp["enc"]
    _, _, cls = mi.run(cap, mask, enc.to(cuda), want_hidden=False, want_cls=True)
    with torch.no_grad():
        ref = med_oracle.itm_logits(sdi, enc, cap, mask, H, depth)
    assert (cls.cpu() - ref).abs().max() < 8e-2


@pytest.mark.parametrize("F,n_img,K,max_length,min_length", [(1, 3, 1, 6, 0), (2, 50, 2, 5, 5), (5, 300, 4, 9, 3), (3, 257, 3, 30, 10)])
def test_generate_edge_shapes(cuda, F, n_img, K, max_length, min_length):
    """One frame / greedy (1 beam) / a single step after the prompt (max_length = prompt + 1) / 4 beams / odd token counts (two
    TMA boxes at 257 and 300) / the defaults of BLIP_Decoder.generate (max_length 30, min_length 10)."""
    name = "tiny"
    m, sd = _decoder(name, "fp16", cuda)
    c, sp = W.MED_CONFIGS[name], W.MED_SPECIAL[name]
    enc = W.image_tokens(F, n_img, c["encoder_width"], seed=F + n_img)
    prompt = torch.tensor([sp["prompt"]], dtype=torch.long).repeat(F, 1)
    out, scores, lens = m.generate(input_ids=prompt, max_length=max_length, min_length=min_length, num_beams=K, eos_token_id=sp["eos"],
                                   pad_token_id=sp["pad"], encoder_hidden_states=enc.to(cuda), return_scores=True)
    ref_toks, ref_scores, _ = med_oracle.generate(sd, enc, sp["prompt"], c["num_attention_heads"], c["num_hidden_layers"], num_beams=K,
                                                  max_length=max_length, min_length=min_length, eos=sp["eos"], pad=sp["pad"])
    got = [out[b, :int(lens[b])].tolist() for b in range(F)]
    assert all(g == r or s > rs - 0.02 for g, r, s, rs in zip(got, ref_toks, scores.cpu().numpy(), ref_scores))
    assert sum(g == r for g, r in zip(got, ref_toks)) >= F - 1
    assert all(len(g) <= max_length and g[:4] == sp["prompt"] for g in got)
    assert out.shape == (F, max(len(g) for g in got))


def test_decoder_forward_long_sequence_and_ragged_frames(cuda):
    """Teacher-forced logits for 40-token sequences (the training max_length of blip.py:110) and a sequence->frame map."""
    name = "tiny"
    m, sd = _decoder(name, "fp16", cuda)
    c, sp = W.MED_CONFIGS[name], W.MED_SPECIAL[name]
    enc = W.image_tokens(2, 9, c["encoder_width"], seed=21)
    ids, _ = W.caption_ids(name, 5, 40, seed=8, min_words=38)
    ids[:, 0] = sp["bos"]
    frame_of = torch.tensor([1, 0, 0, 1, 1])
    _, logits, _ = m.bert.run(ids, None, enc.to(cuda), frame_of_seq=frame_of, causal=True, want_hidden=False, want_logits=True)
    with torch.no_grad():
        ref, _ = med_oracle.decoder_logits(sd, "text_decoder.", ids, enc[frame_of], c["num_attention_heads"], c["num_hidden_layers"])
    assert (logits.cpu() - ref).abs().max() < LOGIT_TOL["fp16"]


def test_generate_chunks_large_batches(cuda, monkeypatch):
    """A batch whose cross-attention K/V would not fit the workspace budget is decoded in chunks, with identical captions."""
    name = "tiny"
    m, _ = _decoder(name, "bf16", cuda)
    c, sp = W.MED_CONFIGS[name], W.MED_SPECIAL[name]
    enc = W.image_tokens(11, 40, c["encoder_width"], seed=13).to(cuda)
    kw = dict(input_ids=torch.tensor([sp["prompt"]], dtype=torch.long).repeat(11, 1), max_length=10, min_length=5, num_beams=3,
              eos_token_id=sp["eos"], pad_token_id=sp["pad"], encoder_hidden_states=enc, return_scores=True)
    full = m.generate(**kw)
    need3 = m.bert._native.lib.vidil_med_generate_workspace_bytes(m.bert._native.handle, 3, 40, 3, 10, 4)
    monkeypatch.setenv("VIDIL_MED_WORKSPACE_GB", str((need3 + 4096) / (1 << 30)))   # room for 3 frames at a time
    before = _lib.launch_count()
    chunked = m.generate(**kw)
    assert all(torch.equal(a, b) for a, b in zip(full, chunked))
    assert _lib.launch_count() - before > 3 * 100                    # four chunks' worth of kernels


def test_generate_stops_early_when_all_frames_are_done(cuda):
    """A decoder whose LM bias makes [SEP] overwhelmingly likely finishes every frame right after min_length: the search stops
    issuing decode steps (far fewer kernels than a full-length run) and still returns what the oracle's loop returns."""
    name = "tiny"
    c, sp = W.MED_CONFIGS[name], W.MED_SPECIAL[name]
    sd = W.med_state_dict(name, "decoder", seed=0)
    sd["text_decoder.cls.predictions.bias"] = sd["text_decoder.cls.predictions.bias"].clone()
    sd["text_decoder.cls.predictions.bias"][sp["eos"]] += 12.0
    m = BertLMHeadModel(_cfg(name), compute_dtype="fp16")
    m.load_state_dict({k[len("text_decoder."):]: v for k, v in sd.items()}, strict=False)
    m = m.to(cuda).eval()
    enc = W.image_tokens(5, 6, c["encoder_width"], seed=17)
    kw = dict(input_ids=torch.tensor([sp["prompt"]], dtype=torch.long).repeat(5, 1), min_length=6, num_beams=3, eos_token_id=sp["eos"],
              pad_token_id=sp["pad"], encoder_hidden_states=enc.to(cuda), return_scores=True)
    m.generate(max_length=40, **kw)                                            # packs the weights
    before = _lib.launch_count()
    out, scores, lens = m.generate(max_length=40, **kw)
    launches = _lib.launch_count() - before
    ref_toks, ref_scores, trace = med_oracle.generate(sd, enc, sp["prompt"], c["num_attention_heads"], c["num_hidden_layers"],
                                                      num_beams=3, max_length=40, min_length=6, eos=sp["eos"], pad=sp["pad"])
    assert len(trace) < 10 and all(trace[-1]["done"])                          # the oracle's loop broke early too
    got = [out[b, :int(lens[b])].tolist() for b in range(5)]
    assert got == ref_toks and all(g[-1] == sp["eos"] and len(g) <= 9 for g in got)
    assert np.allclose(scores.cpu().numpy(), np.asarray(ref_scores, dtype=np.float32), atol=2e-2)
    per_step = 2 * 17 + 8                                                       # kernels of one tiny decode step (2 layers)
    assert launches < 12 * per_step, f"{launches} kernels: the search did not stop early (a 36-step run takes > {30 * per_step})"


@pytest.mark.parametrize("dtype", ["fp16", "bf16"])
def test_generate_sample_tiny_vs_oracle(cuda, dtype):
    """BLIP_Decoder.generate(sample=True) from the image tokens on (blip.py:139-148: top_p 0.9, repetition_penalty 1.1, inherited
    top_k 50) against the oracle's cached decoder driving the same draws.  16-bit operands move a logit by ~1e-2, which moves
    a boundary of the cumulative distribution by ~1e-3 of probability: a few draws in a hundred may fall on the other side, and
    the sequences differ from there on — so most sequences, not all, are identical; every sequence obeys the rules."""
    name = "tiny"
    m, sd = _decoder(name, dtype, cuda)
    c, sp = W.MED_CONFIGS[name], W.MED_SPECIAL[name]
    F, n_img, max_length, min_length = 32, 5, 14, 6
    enc = W.image_tokens(F, n_img, c["encoder_width"], seed=5)
    prompt = torch.tensor([sp["prompt"]], dtype=torch.long).repeat(F, 1)
    uniforms = torch.rand(max_length - len(sp["prompt"]), F, generator=torch.Generator().manual_seed(7))
    out, logp, lens = m.generate(input_ids=prompt, max_length=max_length, min_length=min_length, do_sample=True, top_p=0.9,
                                 eos_token_id=sp["eos"], pad_token_id=sp["pad"], repetition_penalty=1.1,
                                 encoder_hidden_states=enc.to(cuda), return_scores=True, uniforms=uniforms)
    ref, ref_logp = med_oracle.generate_sample(sd, enc, sp["prompt"], c["num_attention_heads"], c["num_hidden_layers"], uniforms.numpy(),
                                               max_length=max_length, min_length=min_length, eos=sp["eos"], pad=sp["pad"],
                                               operand_dtype=_TORCH_DT[dtype])
    got = [out[b, :int(lens[b])].tolist() for b in range(F)]
    same = sum(g == r for g, r in zip(got, ref))
    print(f"generate(sample=True) tiny {dtype}: {same}/{F} sequences identical to the oracle's")
    # measured: fp16 28/32, bf16 16/32 (a sequence is ~10 draws; one draw on the other side of a boundary changes all that follow)
    assert same >= ((F * 3) // 4 if dtype == "fp16" else F // 3)
    for g in got:
        assert g[:4] == sp["prompt"] and sp["eos"] not in g[4:min_length] and len(g) <= max_length
        assert g[-1] == sp["eos"] or len(g) == max_length
    assert out.shape[1] == max(len(g) for g in got)
    # reproducible from the random stream: a generator seed gives the same captions twice, another seed other captions
    a = m.generate(input_ids=prompt, max_length=max_length, min_length=min_length, do_sample=True, top_p=0.9, eos_token_id=sp["eos"],
                   pad_token_id=sp["pad"], repetition_penalty=1.1, encoder_hidden_states=enc.to(cuda), generator=torch.Generator().manual_seed(1))
    b = m.generate(input_ids=prompt, max_length=max_length, min_length=min_length, do_sample=True, top_p=0.9, eos_token_id=sp["eos"],
                   pad_token_id=sp["pad"], repetition_penalty=1.1, encoder_hidden_states=enc.to(cuda), generator=torch.Generator().manual_seed(1))
    d = m.generate(input_ids=prompt, max_length=max_length, min_length=min_length, do_sample=True, top_p=0.9, eos_token_id=sp["eos"],
                   pad_token_id=sp["pad"], repetition_penalty=1.1, encoder_hidden_states=enc.to(cuda), generator=torch.Generator().manual_seed(2))
    assert torch.equal(a, b) and (a.shape != d.shape or not torch.equal(a, d))
