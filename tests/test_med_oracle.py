"""The CPU oracle of the med.py text stack (oracle/med_oracle.py) against fixtures produced by the reference's own
med.py (oracle/make_golden.py golden_med), plus known-answer checks of the restated v4.15 beam search."""
import os

import numpy as np
import pytest
import torch

from oracle import med_oracle, weights as W

TOL = 3e-4


@pytest.mark.parametrize("name,batch,seq_len,n_img,fname", [("tiny", 3, 9, 5, "med_tiny.npz"), ("base_l", 2, 8, 197, "med_base_l.npz")])
def test_med_oracle_matches_reference_fixture(golden_dir, name, batch, seq_len, n_img, fname):
    g = np.load(os.path.join(golden_dir, fname))
    c, sp = W.MED_CONFIGS[name], W.MED_SPECIAL[name]
    H, depth = c["num_attention_heads"], c["num_hidden_layers"]
    enc = W.image_tokens(batch, n_img, c["encoder_width"], seed=0)
    sd = W.med_state_dict(name, "decoder", seed=0)
    ids, _ = W.caption_ids(name, batch, seq_len, seed=0, min_words=seq_len - 2)
    ids[:, 0] = sp["bos"]
    vs = g["vocab"]
    with torch.no_grad():
        logits, _ = med_oracle.decoder_logits(sd, "text_decoder.", ids, enc, H, depth)
        scale = max(1.0, float(np.abs(g["logits"]).max()))
        assert np.abs(logits[:, :, vs].numpy() - g["logits"]).max() < TOL * scale
        assert abs(float(logits.std()) - float(g["logits_std"])) < 1e-3
        # cached decode: prompt, then one token at a time
        _, past = med_oracle.decoder_logits(sd, "text_decoder.", ids[:, :4], enc, H, depth)
        for t in range(4, seq_len):
            lg, past = med_oracle.decoder_logits(sd, "text_decoder.", ids[:, t:t + 1], enc, H, depth, past=past)
            assert np.abs(lg[:, 0, vs].numpy() - g["step_logits"][:, t - 4]).max() < TOL * scale
        sdi = W.med_state_dict(name, "itm", seed=0)
        cap, mask = W.caption_ids(name, batch, seq_len, seed=1)
        cap[:, 0] = sp["enc"]
        out = med_oracle.itm_logits(sdi, enc, cap, mask, H, depth)
    assert np.abs(out.numpy() - g["itm_logits"]).max() < TOL


def _table_logits(L):
    it = iter(L)
    return lambda ids, beam_idx: next(it)


def test_beam_search_known_answers():
    """Hand-worked cases of the v4.15 rules: eos banned before min_length, a finished hypothesis is scored by
    sum_logprobs / len, an eos ranked below num_beams is skipped, the search stops once the worst kept hypothesis is
    at least as good as the best candidate of the step, open beams are finalised at max_length."""
    K, eos = 2, 1
    # (a) token 2 dominant (.5), eos .3: with length_penalty 1 longer sequences of the dominant token keep winning, so the
    # best open beam at max_length is returned and there is no room for an eos
    row = np.log(np.array([0.01, 0.30, 0.50, 0.10, 0.05, 0.04], dtype=np.float32))
    L = [np.tile(row, (K, 1)) for _ in range(8)]
    toks, scores, trace = med_oracle.beam_search_from_logits(_table_logits(L), 1, [5], num_beams=K, max_length=6, min_length=3,
                                                            eos=eos, pad=0)
    assert trace[0]["tokens"].tolist() == [2, 3] and trace[0]["beam_idx"].tolist() == [0, 0]
    assert trace[1]["tokens"].tolist() == [2, 3] and trace[1]["beam_idx"].tolist() == [0, 0]   # tie -> lower flat index
    assert toks[0] == [5, 2, 2, 2, 2, 2] and abs(scores[0] - 5 * np.log(0.5) / 6) < 1e-5
    assert len(trace) == 5
    # (b) eos dominant (.6) once allowed: hypotheses [5,2,2] (score (2 ln .3 + ln .6)/3) and [5,2,2,2]; after the second one
    # worst_score == best candidate / cur_len, so the frame is done after 4 steps
    row = np.log(np.array([0.01, 0.60, 0.30, 0.05, 0.02, 0.02], dtype=np.float32))
    L = [np.tile(row, (K, 1)) for _ in range(8)]
    toks, scores, trace = med_oracle.beam_search_from_logits(_table_logits(L), 1, [5], num_beams=K, max_length=8, min_length=3,
                                                            eos=eos, pad=0)
    assert toks[0] == [5, 2, 2, eos]
    assert abs(scores[0] - (2 * np.log(0.3) + np.log(0.6)) / 3) < 1e-5
    assert len(trace) == 4 and trace[-1]["done"] == [True]
    assert trace[2]["tokens"].tolist() == [2, 3]            # ranks 1 and 3; the eos at rank 2 (>= num_beams) was skipped


def test_beam_search_open_beams_finalised_at_max_length():
    V, K = 5, 3
    rng = np.random.default_rng(0)
    L = [rng.standard_normal((K * 2, V)).astype(np.float32) for _ in range(10)]
    for l in L:
        l[:, 1] = -50.0                                   # eos never competitive
    toks, scores, trace = med_oracle.beam_search_from_logits(_table_logits(L), 2, [4, 3], num_beams=K, max_length=7,
                                                            min_length=0, eos=1, pad=0)
    assert all(len(t) == 7 and 1 not in t for t in toks)   # max_length reached: no room for an eos
    assert len(trace) == 5
    # the winner is the best open beam: score = beam_score / 7
    for b in range(2):
        assert abs(scores[b] - trace[-1]["beam_scores"][b * K] / 7) < 1e-6


def test_generate_runs_cached_and_uncached_identically():
    """The cached decoder path (prepare_inputs_for_generation + _reorder_cache) and a full re-forward of the current
    sequences give the same captions on the tiny model."""
    name = "tiny"
    c, sp = W.MED_CONFIGS[name], W.MED_SPECIAL[name]
    H, depth = c["num_attention_heads"], c["num_hidden_layers"]
    sd = W.med_state_dict(name, "decoder", seed=0)
    enc = W.image_tokens(4, 5, c["encoder_width"], seed=2)
    toks, scores, _ = med_oracle.generate(sd, enc, sp["prompt"], H, depth, num_beams=3, max_length=12, min_length=5, eos=sp["eos"])
    enc3 = enc.repeat_interleave(3, dim=0)

    def full(ids, beam_idx):
        with torch.no_grad():
            lg, _ = med_oracle.decoder_logits(sd, "text_decoder.", torch.from_numpy(ids), enc3, H, depth)
        return lg[:, -1].numpy()

    toks2, scores2, _ = med_oracle.beam_search_from_logits(full, 4, sp["prompt"], 3, 12, 5, sp["eos"], 0)
    assert toks == toks2
    assert np.allclose(scores, scores2, atol=1e-5)
    assert all(t[:4] == sp["prompt"] for t in toks)


def test_beam_search_single_beam_is_greedy_and_scores_are_consistent():
    """Independent checks of the restated search: with one beam it is greedy decoding (arg-max of the admissible log-probs each
    step, stop at eos), and for any beam count the returned score is the sum of the returned tokens' log-probs divided by the
    hypothesis length (eos excluded from the length, included in the sum), recomputed here from the logits tables."""
    V, eos, prompt, max_length, min_length = 40, 3, [7, 8], 12, 4
    rng = np.random.default_rng(5)
    for K in (1, 2, 3):
        S = max_length - len(prompt)
        L = [rng.standard_normal((2 * K, V)).astype(np.float32) * 3 for _ in range(S)]
        for l in L:
            l[:, eos] += 1.5
        # tables that depend on the step only (rows identical across beams), so that a sequence's log-prob is path-independent
        L = [np.tile(l[:1], (2 * K, 1)) for l in L]
        it = iter(L)
        toks, scores, _ = med_oracle.beam_search_from_logits(lambda ids, bi: next(it), 2, prompt, K, max_length, min_length, eos, 0)
        lp = [torch.log_softmax(torch.from_numpy(l[0]), -1).numpy() for l in L]
        for t, s in zip(toks, scores):
            gen = t[len(prompt):]
            total = sum(float(lp[i][tok]) for i, tok in enumerate(gen))
            if t[-1] == eos:
                assert abs(s - total / (len(t) - 1)) < 1e-5
            else:
                assert len(t) == max_length and abs(s - total / len(t)) < 1e-5
            assert eos not in t[:min_length]                                  # MinLengthLogitsProcessor
        if K == 1:
            greedy = list(prompt)
            for i in range(S):
                row = lp[i].copy()
                if len(greedy) < min_length:
                    row[eos] = -np.inf
                nxt = int(np.argmax(row))
                if nxt == eos:
                    break
                greedy.append(nxt)
            expect = greedy + ([eos] if len(greedy) < max_length else [])
            assert toks[0] == expect and toks[1] == expect


def test_decision_margins_bound_what_a_perturbation_can_change():
    """`margins`: a frame whose every comparison was decided by more than the accumulated perturbation of the candidate scores
    returns the same tokens under that perturbation (the property tests/test_gpu_med.py relies on to assert token equality
    against 16-bit kernels), and some frames below the bound do change — the gate is not vacuous."""
    rng = np.random.default_rng(0)
    V, K, eos, max_length, min_length = 40, 3, 1, 12, 5
    table = rng.standard_normal((V, max_length, V)).astype(np.float32)
    noise = rng.uniform(-1.0, 1.0, (V, V, max_length, V)).astype(np.float32)

    def lm(scale):
        def step(ids, beam_idx):
            cur = ids.shape[1]
            return table[ids[:, -1], cur - 1] + 0.5 * table[ids[:, 0], cur - 1] + scale * noise[ids[:, 0], ids[:, -1], cur - 1]
        return step

    n_above, changed = {5e-4: 0, 3e-2: 0}, {5e-4: 0, 3e-2: 0}
    for first in range(2, V):
        for second in range(2, 12):
            prompt = [first, second]
            m = []
            t0, _, _ = med_oracle.beam_search_from_logits(lm(0.0), 1, prompt, num_beams=K, max_length=max_length,
                                                          min_length=min_length, eos=eos, pad=0, margins=m)
            for eps in n_above:
                t1, _, _ = med_oracle.beam_search_from_logits(lm(eps), 1, prompt, num_beams=K, max_length=max_length,
                                                              min_length=min_length, eos=eos, pad=0)
                # |noise| <= eps: a log-probability moves by at most 2 * eps per step, the difference of two by twice that
                bound = 4 * eps * (max_length - len(prompt))
                n_above[eps] += m[0] > bound
                changed[eps] += t0 != t1
                assert t0 == t1 or m[0] <= bound, (prompt, eps, m[0])
    assert n_above[5e-4] > 50            # the small perturbation leaves many frames decided, and none of them moved
    assert changed[3e-2] > 20            # the large one does change captions — only those the margin flagged


def _sampling_lm(V, seed=0):
    table = np.random.default_rng(seed).standard_normal((V, V)).astype(np.float32) * 2.0
    return lambda ids, _: table[ids[:, -1]]


def test_sampling_oracle_known_answers():
    """u = 0 always takes the most probable surviving token; a banned eos cannot be drawn before min_length; a frame that has
    drawn eos is padded while the others go on; the sequence ends with eos."""
    V, eos, pad = 30, 1, 0
    lm = _sampling_lm(V)
    steps, B = 8, 3
    toks, logp = med_oracle.sample_from_logits(lm, B, [5, 7], np.zeros((steps, B), np.float32), max_length=10, min_length=6, eos=eos, pad=pad)
    ids = np.array([5, 7])
    expect = [5, 7]
    for _ in range(steps):
        kept, _ = med_oracle.process_sampling_scores(lm(np.array([expect]), None)[0], np.array(expect), len(expect), 6, eos)
        expect.append(int(kept[0]))
        if expect[-1] == eos:
            break
    assert toks[0] == expect and toks[1] == expect
    assert all(eos not in t[:6] for t in toks)
    # an eos-heavy model: every frame stops right at min_length
    table = np.zeros((V, V), np.float32)
    table[:, eos] = 10.0
    toks, _ = med_oracle.sample_from_logits(lambda ids, _: table[ids[:, -1]], 2, [5, 7], np.full((8, 2), 0.3, np.float32), max_length=10,
                                            min_length=4, eos=eos, pad=pad)
    assert all(len(t) == 5 and t[-1] == eos for t in toks)


def test_sampling_oracle_draws_follow_the_filtered_distribution():
    """Inverse-CDF draws with uniform numbers reproduce the renormalised top-k / top-p distribution (chi-square)."""
    V, eos = 50, 1
    rng = np.random.default_rng(3)
    logits = (rng.standard_normal(V) * 1.5).astype(np.float32)
    kept, v = med_oracle.process_sampling_scores(logits, np.array([3, 4]), 2, 0, eos, top_k=20, top_p=0.9, repetition_penalty=1.1)
    p = np.exp(v - v[0]); p /= p.sum()
    n = 20000
    toks, _ = med_oracle.sample_from_logits(lambda ids, _: np.tile(logits, (ids.shape[0], 1)), n, [3, 4],
                                            rng.random((1, n)).astype(np.float32), max_length=3, min_length=0, eos=eos, top_k=20)
    drawn = np.array([t[2] for t in toks])
    assert set(drawn.tolist()) <= set(kept.tolist())
    counts = np.array([(drawn == k).sum() for k in kept])
    chi2 = ((counts - n * p) ** 2 / (n * p)).sum()
    assert chi2 < 3 * len(kept) + 20
