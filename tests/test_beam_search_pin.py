"""Pins the oracle's beam search (oracle/med_oracle.beam_search_from_logits) against REAL transformers code.

The reference's captioner calls `self.text_decoder.generate(num_beams=3, ...)` (models/blip.py:150-158) and lets transformers
(v4.15-era, un-vendored) run the search.  That version is not installable here; the installed transformers (5.x) carries the
vectorised rewrite of the same search, which differs from v4.15 in exactly three places (med_oracle.RULES: the length a
finished hypothesis is normalised by, the length open beams are normalised by at max_length, and the operands of the stopping
heuristic).  The oracle implements both rule sets on ONE shared body — candidate selection over 2K continuations, the
eos-inside-the-first-K rule, MinLength, beam re-ordering, hypothesis bookkeeping, early termination, finalisation — so driving
the installed `generate()` with a small random language model and requiring token- and score-identical results under
rules="v5" pins everything the two versions share.  The three v4.15 lines themselves stay anchored on the published v4.15.0
source and on the hand-worked cases of tests/test_med_oracle.py.
"""
import numpy as np
import pytest
import torch

from oracle import med_oracle

transformers = pytest.importorskip("transformers")


def _lm(seed, vocab, gain):
    torch.manual_seed(seed)
    cfg = transformers.GPT2Config(vocab_size=vocab, n_positions=32, n_embd=32, n_layer=2, n_head=2, bos_token_id=0,
                                  eos_token_id=1, pad_token_id=0)
    m = transformers.GPT2LMHeadModel(cfg).double().eval()
    with torch.no_grad():
        for p in m.parameters():
            p.mul_(gain)                  # sharper next-token distributions: early eos, early stopping and max_length all occur
    return m


CASES = [  # seed, vocabulary, gain, prompt length, beams, max_length, min_length
    (0, 14, 6.0, 4, 3, 12, 6), (1, 9, 3.0, 2, 3, 20, 5), (2, 30, 10.0, 1, 3, 9, 0), (3, 200, 6.0, 4, 3, 20, 5),
    (4, 14, 3.0, 2, 4, 20, 5), (5, 9, 10.0, 1, 4, 9, 0), (6, 30, 6.0, 4, 5, 12, 6), (7, 200, 3.0, 2, 2, 20, 5),
    (8, 14, 10.0, 4, 2, 20, 5), (9, 9, 6.0, 4, 3, 20, 5),
]


@pytest.mark.parametrize("seed,vocab,gain,P,K,max_length,min_length", CASES)
def test_oracle_beam_search_equals_installed_transformers(seed, vocab, gain, P, K, max_length, min_length):
    eos, pad, B = 1, 0, 12
    m = _lm(seed, vocab, gain)
    prompt = torch.randint(2, vocab, (1, P)).repeat(B, 1)
    prompt[:, 0] = torch.randint(2, vocab, (B,))               # one "frame" per row
    with torch.no_grad():
        out = m.generate(input_ids=prompt, attention_mask=torch.ones_like(prompt), max_length=max_length, min_length=min_length,
                         num_beams=K, eos_token_id=eos, pad_token_id=pad, do_sample=False, use_cache=False,
                         return_dict_in_generate=True, output_scores=True, repetition_penalty=1.0, length_penalty=1.0,
                         early_stopping=False)

    def step(ids, beam_idx):
        with torch.no_grad():
            return m(torch.from_numpy(ids)).logits[:, -1].float().numpy()

    differs_under_v415 = 0
    for b in range(B):
        hf = out.sequences[b].tolist()
        gen = hf[P:]
        if eos in gen:
            hf = hf[:P + gen.index(eos) + 1]                  # drop the padding generate() adds after the eos
        toks, scores, _ = med_oracle.beam_search_from_logits(step, 1, prompt[b].tolist(), num_beams=K, max_length=max_length,
                                                             min_length=min_length, eos=eos, pad=pad, rules="v5")
        assert toks[0] == hf, (b, toks[0], hf)
        assert abs(scores[0] - float(out.sequences_scores[b])) < 1e-5
        t415, _, _ = med_oracle.beam_search_from_logits(step, 1, prompt[b].tolist(), num_beams=K, max_length=max_length,
                                                        min_length=min_length, eos=eos, pad=pad)
        differs_under_v415 += t415[0] != hf
    print(f"seed {seed}: {B}/{B} identical to transformers {transformers.__version__}; the v4.15 rules give another caption for "
          f"{differs_under_v415}")


def test_batched_search_equals_per_frame_search():
    """The oracle searches a batch of frames in one call on the product path's tests; frames must not interact."""
    m = _lm(3, 50, 6.0)
    B, K = 5, 3
    prompt = [7, 3]

    def make_step(rows):
        def step(ids, beam_idx):
            first = np.repeat(np.asarray(rows, dtype=np.int64), K)[:, None] + 2
            with torch.no_grad():
                return m(torch.from_numpy(np.concatenate([first, ids], axis=1))).logits[:, -1].float().numpy()
        return step

    for rules in med_oracle.RULES:
        all_t, all_s, _ = med_oracle.beam_search_from_logits(make_step(list(range(B))), B, prompt, num_beams=K, max_length=12,
                                                             min_length=4, eos=1, pad=0, rules=rules)
        for b in range(B):
            t, s, _ = med_oracle.beam_search_from_logits(make_step([b]), 1, prompt, num_beams=K, max_length=12, min_length=4,
                                                         eos=1, pad=0, rules=rules)
            assert t[0] == all_t[b] and abs(s[0] - all_s[b]) < 1e-6


def test_sampling_processors_equal_installed_transformers():
    """The processors and warpers of the sampling path (blip.py:141-148: repetition_penalty 1.1, min_length, the inherited
    top_k 50, top_p 0.9) as restated in med_oracle.process_sampling_scores keep exactly the tokens, with exactly the scores,
    that the installed transformers' own RepetitionPenaltyLogitsProcessor -> MinLengthLogitsProcessor -> TopKLogitsWarper ->
    TopPLogitsWarper chain keeps.  (Installed TopP sorts ascending and drops mass <= 1 - top_p, v4.15 sorts descending and drops
    what follows mass > top_p: the same set except at exact equality.  Rows with exactly tied scores are compared by the
    multiset of kept scores: which of two tied tokens survives a cut between them is sort-order dependent in transformers.)"""
    lp = pytest.importorskip("transformers.generation.logits_process")
    rng = np.random.default_rng(0)
    for trial in range(300):
        V = int(rng.choice([40, 200, 3000]))
        logits = (rng.standard_normal(V) * rng.choice([0.5, 2.0, 6.0])).astype(np.float32)
        tied = trial % 5 == 0
        if tied:
            logits[rng.integers(0, V, 10)] = logits[0]
        seq = rng.integers(0, V, size=int(rng.integers(1, 12)))
        min_length, eos = int(rng.choice([0, 5, 20])), 1
        top_k, top_p = int(rng.choice([5, 50])), float(rng.choice([0.5, 0.9, 0.99]))
        chain = lp.LogitsProcessorList([lp.RepetitionPenaltyLogitsProcessor(1.1), lp.MinLengthLogitsProcessor(min_length, eos),
                                        lp.TopKLogitsWarper(top_k), lp.TopPLogitsWarper(top_p)])
        out = chain(torch.from_numpy(seq)[None].long(), torch.from_numpy(logits)[None].clone())[0].numpy()
        toks, v = med_oracle.process_sampling_scores(logits, seq, len(seq), min_length, eos, top_k, top_p, 1.1)
        kept = np.nonzero(np.isfinite(out))[0]
        if tied:
            assert np.allclose(np.sort(out[kept]), np.sort(v), rtol=1e-6)
        else:
            assert set(kept.tolist()) == set(toks.tolist()), trial
            assert np.allclose(out[toks], v, rtol=1e-6)
        assert np.all(np.diff(v) <= 0)                             # descending: the order the draw walks


def test_sampling_loop_equals_installed_transformers_generate(monkeypatch):
    """The whole sampling loop — processor order, the inherited top_k, eos / pad handling of finished rows, the stop when every
    row is finished — against the installed `generate(do_sample=True, ...)` called with the arguments of blip.py:141-148.
    torch.multinomial is replaced, for the duration of the call, by the inverse-CDF lookup on caller-supplied uniform numbers
    that the oracle and the native kernel use (descending probability, ties by ascending token id), so the two loops consume
    the same random stream and must return the same sequences."""
    eos, pad = 1, 0
    total = same = 0
    for seed, vocab, gain, P, max_length, min_length in [(0, 60, 3.0, 4, 14, 6), (1, 200, 6.0, 2, 12, 0), (2, 30, 2.0, 3, 20, 5),
                                                         (3, 500, 4.0, 4, 20, 5)]:
        m = _lm(seed, vocab, gain)
        B = 16
        g = torch.Generator().manual_seed(100 + seed)
        prompt = torch.randint(2, vocab, (1, P), generator=g).repeat(B, 1)
        prompt[:, 0] = torch.randint(2, vocab, (B,), generator=g)
        uniforms = torch.rand(max_length - P, B, generator=g).numpy().astype(np.float32)
        state = {"step": 0}

        def fake_multinomial(probs, num_samples=1, **kw):
            assert num_samples == 1
            p = probs.double().numpy()
            out = np.zeros((p.shape[0], 1), dtype=np.int64)
            for b in range(p.shape[0]):
                order = np.lexsort((np.arange(p.shape[1]), -p[b]))
                order = order[p[b][order] > 0]
                cum = np.cumsum(p[b][order])
                j = int(np.searchsorted(cum, float(uniforms[state["step"], b]) * cum[-1], side="right"))
                out[b, 0] = order[min(j, len(order) - 1)]
            state["step"] += 1
            return torch.from_numpy(out)

        monkeypatch.setattr(torch, "multinomial", fake_multinomial)
        with torch.no_grad():
            hf = m.generate(input_ids=prompt, attention_mask=torch.ones_like(prompt), max_length=max_length, min_length=min_length,
                            do_sample=True, top_p=0.9, top_k=50, num_return_sequences=1, eos_token_id=eos, pad_token_id=pad,
                            repetition_penalty=1.1, use_cache=False)
        monkeypatch.undo()

        def step(ids, _):
            with torch.no_grad():
                return m(torch.from_numpy(ids)).logits[:, -1].float().numpy()

        ref, _ = med_oracle.sample_from_logits(step, B, prompt[0].tolist(), uniforms, max_length, min_length, eos, pad, 50, 0.9, 1.1) \
            if bool((prompt == prompt[0]).all()) else (None, None)
        if ref is None:                                        # per-row prompts: one oracle call per row with that row's uniforms
            ref = []
            for b in range(B):
                r, _ = med_oracle.sample_from_logits(step, 1, prompt[b].tolist(), uniforms[:, b:b + 1], max_length, min_length, eos,
                                                     pad, 50, 0.9, 1.1)
                ref.append(r[0])
        for b in range(B):
            seq = hf[b].tolist()
            gen = seq[P:]
            if eos in gen:
                assert all(t == pad for t in gen[gen.index(eos) + 1:])          # finished rows are padded
                seq = seq[:P + gen.index(eos) + 1]
            total += 1
            same += seq == ref[b]
        assert hf.shape[1] == max(len(r) for r in ref)                         # the loop stops when every row is finished
    print(f"sampling loop: {same}/{total} sequences identical to transformers {transformers.__version__}")
    assert same >= total - 1      # one draw in ~10^5 may sit within fp32 rounding of a cumulative-probability boundary
