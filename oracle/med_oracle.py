"""CPU restatement of the reference's `models/med.py` text stack on the CapFilt path — the parity oracle for
the caption decoder (BLIP_Decoder.generate, models/blip.py:127-167) and the ITM filter head
(BLIP_ITM.forward, models/blip_itm.py:41-58).

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
leg may import this module; the product path (vidil_b200/) never does.

Parity status
  * The network arithmetic (embeddings, BertLayer with self- and cross-attention, LM head, ITM head) is pinned
    against *outputs of the reference itself*: oracle/make_golden.py imports the unmodified
    /root/reference/models/med.py (behind the import shims SURVEY.md §8c lists), runs it on seeded inputs and
    commits the results under tests/golden/med_*.npz; tests/test_oracle_golden.py checks this file against them.
  * Beam search: **pinned to real transformers code for everything but three lines.**  The reference delegates to
    `transformers` `generate()` (un-vendored, unpinned, `docker/requirements.txt:9`; med.py is "based on v4.15.0" and only
    imports under transformers of that era, which cannot be installed here and whose `generate()` cannot drive med.py under
    the installed 5.5).  `beam_search_from_logits` restates the published v4.15.0 algorithm
    (`GenerationMixin.beam_search`, `BeamSearchScorer.process/finalize`, `BeamHypotheses.add/is_done`,
    `MinLengthLogitsProcessor`) with the arguments of the reference's call site (blip.py:150-158,
    run_video_CapFilt.py:102: num_beams 3, max_length 20, min_length 5, length_penalty 1.0, early_stopping False,
    repetition_penalty 1.0).  The same function body also runs the rule set of the installed transformers (`rules="v5"`,
    three lines differ — see RULES below), and under those rules it is token- and score-identical to the installed
    `generate(num_beams=K)` on 120 searches of random language models (tests/test_beam_search_pin.py).  What stays
    "parity unpinned" are the three v4.15-specific lines (hypothesis length normalisation, max_length normalisation,
    operands of the stopping heuristic): anchored on the published source, the call site and hand-worked cases only.

Plain functional tensor arithmetic on a parameter dict with the reference's state_dict keys.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------------------------------
# Operand-dtype emulation.  The native path keeps the residual stream, LayerNorm / softmax statistics and every accumulator
# in fp32 but feeds the tensor cores 16-bit operands and stores qkv / attention output / GELU hidden / cross K,V in 16 bits.
# With `operand_dtype` set (torch.bfloat16 or torch.float16, see `emulate`) the functions below round at exactly those
# points — weights and activations entering a Linear, the stored projections, the probabilities entering P @ V, the
# tensors between LM-head stages — and compute everything else in fp32 like the default oracle.  What remains between this
# emulation and the kernels is accumulation order and the approximate exp / erf (~1e-5 on BERT-base logits instead of the
# ~0.2 that bf16 operand rounding alone causes), so caption TOKENS can be asserted equal (tests/test_gpu_med.py); the plain
# fp32 oracle stays the reference restatement and the reported statistic.
# ----------------------------------------------------------------------------------------------------------------------
_OPERAND_DTYPE = None


class emulate:
    """with med_oracle.emulate(torch.bfloat16): ... — round 16-bit operands / stored tensors like the native kernels."""

    def __init__(self, dtype):
        self.dtype = dtype

    def __enter__(self):
        global _OPERAND_DTYPE
        self.prev, _OPERAND_DTYPE = _OPERAND_DTYPE, self.dtype
        return self

    def __exit__(self, *exc):
        global _OPERAND_DTYPE
        _OPERAND_DTYPE = self.prev


def _r(x: torch.Tensor) -> torch.Tensor:
    """Round to the emulated operand dtype (identity for the default fp32 oracle)."""
    return x if _OPERAND_DTYPE is None else x.to(_OPERAND_DTYPE).to(torch.float32)


def _linear(x, w, b):
    """A Linear as the tensor cores see it: 16-bit x and w (when emulating), exact products, fp32 accumulation, fp32 bias."""
    return F.linear(_r(x), _r(w), b)


def embeddings(sd: dict, pre: str, input_ids: torch.Tensor, past_len: int, eps: float) -> torch.Tensor:
    """BertEmbeddings.forward, med.py:74-96: word + absolute position -> LayerNorm (no token-type term)."""
    T = input_ids.shape[1]
    x = sd[pre + "embeddings.word_embeddings.weight"][input_ids]                              # :88
    x = x + sd[pre + "embeddings.position_embeddings.weight"][past_len:past_len + T][None]     # :85,:93-94
    D = x.shape[-1]
    return F.layer_norm(x, (D,), sd[pre + "embeddings.LayerNorm.weight"], sd[pre + "embeddings.LayerNorm.bias"], eps)


def _flash_emulated(s: torch.Tensor, v: torch.Tensor, block: int = 64) -> torch.Tensor:
    """softmax(s) @ v computed the way the flash-style tensor-core kernels do (see attention_block)."""
    m = torch.full(s.shape[:-1], float("-inf"), dtype=s.dtype, device=s.device)
    l = torch.zeros_like(m)
    o = torch.zeros(s.shape[:-1] + (v.shape[-1],), dtype=s.dtype, device=s.device)
    for j0 in range(0, s.shape[-1], block):
        sb = s[..., j0:j0 + block]
        m_new = torch.maximum(m, sb.max(dim=-1).values)
        corr = torch.exp(m - m_new)
        p = torch.exp(sb - m_new[..., None])
        l = l * corr + p.sum(dim=-1)
        o = o * corr[..., None] + _r(p) @ v[..., j0:j0 + block, :]
        m = m_new
    return o / l[..., None]


def _heads(x: torch.Tensor, H: int) -> torch.Tensor:
    B, T, D = x.shape
    return x.view(B, T, H, D // H).permute(0, 2, 1, 3)                                        # transpose_for_scores :141-144


def attention_block(sd: dict, p: str, x: torch.Tensor, kv_src: torch.Tensor, add_mask, H: int, eps: float, past=None):
    """BertAttention = BertSelfAttention (med.py:146-232) + BertSelfOutput (:235-246).
    `kv_src` is x for self-attention, the image tokens for cross-attention; `past` = (K, V) of earlier positions."""
    q = _heads(_r(_linear(x, sd[p + "self.query.weight"], sd[p + "self.query.bias"])), H)
    k = _heads(_r(_linear(kv_src, sd[p + "self.key.weight"], sd[p + "self.key.bias"])), H)
    v = _heads(_r(_linear(kv_src, sd[p + "self.value.weight"], sd[p + "self.value.bias"])), H)
    if past is not None:
        k = torch.cat([past[0], k], dim=2)                                                     # :173-174
        v = torch.cat([past[1], v], dim=2)
    s = q @ k.transpose(-1, -2) / math.sqrt(q.shape[-1])                                       # :184,:202
    if add_mask is not None:
        s = s + add_mask                                                                       # :205
    if _OPERAND_DTYPE is None:
        ctx = (s.softmax(dim=-1) @ v).permute(0, 2, 1, 3).reshape(x.shape)                     # :208, :222-226
    elif kv_src is x and past is not None:
        # one-token decode step of the self-attention (med_self_attn_decode_kernel): probabilities normalised in fp32 and
        # multiplied with the 16-bit cached V without being rounded
        ctx = _r((s.softmax(dim=-1) @ v).permute(0, 2, 1, 3).reshape(x.shape))
    else:
        # the tensor-core attentions (attention_x_kernel, cross_decode_mma_kernel): online softmax over 64-key blocks; the
        # un-normalised exponentials are rounded to the operand type for P.V, the row sum accumulates them unrounded, and the
        # output is divided by the sum at the end
        ctx = _r(_flash_emulated(s, v).permute(0, 2, 1, 3).reshape(x.shape))
    out = _linear(ctx, sd[p + "output.dense.weight"], sd[p + "output.dense.bias"])
    D = x.shape[-1]
    out = F.layer_norm(out + x, (D,), sd[p + "output.LayerNorm.weight"], sd[p + "output.LayerNorm.bias"], eps)  # :245
    return out, (k, v)


def layer(sd: dict, p: str, x, self_mask, enc, H: int, eps: float, past=None, mode: str = "multimodal"):
    """BertLayer.forward, med.py:333-384: self-attention -> (cross-attention) -> intermediate(GELU erf) -> output."""
    x, present = attention_block(sd, p + "attention.", x, x, self_mask, H, eps, past)
    if mode == "multimodal":
        x, _ = attention_block(sd, p + "crossattention.", x, enc, None, H, eps)                # all-ones image mask
    h = _r(F.gelu(_linear(x, sd[p + "intermediate.dense.weight"], sd[p + "intermediate.dense.bias"])))   # :297-303
    h = _linear(h, sd[p + "output.dense.weight"], sd[p + "output.dense.bias"])
    D = x.shape[-1]
    x = F.layer_norm(h + x, (D,), sd[p + "output.LayerNorm.weight"], sd[p + "output.LayerNorm.bias"], eps)   # :316
    return x, present


def self_mask(attention_mask: torch.Tensor, T: int, past_len: int, causal: bool) -> torch.Tensor:
    """get_extended_attention_mask, med.py:609-668: additive 0 / -10000 mask [B,1,T,past+T]."""
    am = attention_mask.to(torch.float32)
    if causal:
        ids = torch.arange(T, device=am.device)
        cm = (ids[None, :] <= ids[:, None]).to(torch.float32)                                  # :636
        if past_len:
            cm = torch.cat([torch.ones(T, past_len, device=am.device), cm], dim=-1)            # :641-649
        ext = cm[None, None] * am[:, None, None, :]                                            # :651
    else:
        ext = am[:, None, None, :]                                                             # :653
    return (1.0 - ext) * -10000.0                                                              # :667


def bert_forward(sd: dict, pre: str, input_ids, attention_mask, enc, H: int, depth: int, eps: float = 1e-12,
                 causal: bool = False, past=None, mode: str = "multimodal"):
    """BertModel.forward, med.py:700-809 (add_pooling_layer=False).  Returns (last_hidden_state, presents)."""
    B, T = input_ids.shape
    past_len = 0 if past is None else past[0][0].shape[2]
    if attention_mask is None:
        attention_mask = torch.ones(B, past_len + T, device=input_ids.device)
    x = embeddings(sd, pre, input_ids, past_len, eps)
    m = self_mask(attention_mask, T, past_len, causal)
    presents = []
    for i in range(depth):
        x, pkv = layer(sd, f"{pre}encoder.layer.{i}.", x, m, enc, H, eps, None if past is None else past[i], mode)
        presents.append(pkv)
    return x, presents


def lm_head(sd: dict, pre: str, x: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
    """BertOnlyMLMHead, med.py:501-541: dense -> GELU -> LayerNorm -> decoder (+ output-only bias)."""
    c = pre + "cls.predictions."
    h = _r(F.gelu(_linear(x, sd[c + "transform.dense.weight"], sd[c + "transform.dense.bias"])))
    D = h.shape[-1]
    h = _r(F.layer_norm(h, (D,), sd[c + "transform.LayerNorm.weight"], sd[c + "transform.LayerNorm.bias"], eps))
    return _linear(h, sd[c + "decoder.weight"], sd[c + "bias"])


def decoder_logits(sd: dict, pre: str, input_ids, image_embeds, H: int, depth: int, past=None):
    """BertLMHeadModel.forward, med.py:871-893 with is_decoder=True: logits for every input position."""
    x, presents = bert_forward(sd, pre + "bert.", input_ids, None, image_embeds, H, depth, causal=True, past=past)
    return lm_head(sd, pre, x), presents


def itm_logits(sd: dict, image_embeds, input_ids, attention_mask, H: int, depth: int) -> torch.Tensor:
    """BLIP_ITM.forward(match_head='itm'), blip_itm.py:49-57: multimodal encoder -> itm_head on the [CLS]/[ENC] row."""
    x, _ = bert_forward(sd, "text_encoder.", input_ids, attention_mask, image_embeds, H, depth, causal=False)
    return F.linear(x[:, 0, :], sd["itm_head.weight"], sd["itm_head.bias"])   # fp32 SIMT head in the native path too


# ----------------------------------------------------------------------------------------------------------------------
# Beam search — transformers v4.15.0 semantics (restated; see the header).
# ----------------------------------------------------------------------------------------------------------------------
class _BeamHyps:
    """BeamHypotheses (v4.15.0 generation_beam_search.py): keeps the num_beams best finished hypotheses."""

    def __init__(self, num_beams: int, length_penalty: float):
        self.num_beams, self.length_penalty = num_beams, length_penalty
        self.beams: list = []
        self.worst_score = 1e9

    def add(self, hyp: list, sum_logprobs: float, norm_len: int | None = None) -> float:
        """`norm_len`: the length the score is normalised by (v4.15: the hypothesis as stored, len(hyp)).
        Returns the gap (in sum-of-log-probability units) by which the comparisons made here were decided."""
        norm_len = len(hyp) if norm_len is None else norm_len
        score = sum_logprobs / (norm_len ** self.length_penalty)
        gap = math.inf
        if len(self.beams) >= self.num_beams:
            gap = abs(score - self.worst_score) * min(norm_len, min(n for _, _, n in self.beams)) ** self.length_penalty
        if len(self.beams) < self.num_beams or score > self.worst_score:
            self.beams.append((score, hyp, norm_len))
            if len(self.beams) > self.num_beams:
                ranked = sorted([(s, idx) for idx, (s, _, _) in enumerate(self.beams)])
                del self.beams[ranked[0][1]]
                self.worst_score = ranked[1][0]
            else:
                self.worst_score = min(score, self.worst_score)
        return gap

    def is_done(self, best_sum_logprobs: float, cur_len: int) -> bool:
        if len(self.beams) < self.num_beams:
            return False
        return self.worst_score >= best_sum_logprobs / cur_len ** self.length_penalty    # early_stopping False


# What the vectorised beam search of transformers >= 4.50 (checked against the installed 5.5, generation/utils.py
# `_beam_search` / `_update_finished_beams` / `_check_early_stop_heuristic`) does differently from v4.15 for the call-site
# arguments of blip.py:150-158.  Everything else below is shared by the two rule sets, which is what lets the installed
# transformers pin the restatement (tests/test_beam_search_pin.py) although it is not the version the reference ran:
#   * a finished hypothesis is normalised by its generated length INCLUDING the eos and EXCLUDING the prompt,
#     (cur_len + 1 - prompt_len); v4.15 divides by the stored hypothesis, prompt included, eos not (cur_len);
#   * beams still open at max_length are normalised by (max_length - prompt_len); v4.15 by max_length;
#   * the stopping heuristic compares the worst kept hypothesis with the best beam that stays OPEN after the step, over
#     (cur_len + 1 - prompt_len); v4.15 with the best of all 2K candidates of the step (eos included), over cur_len.
RULES = ("v4.15", "v5")


def topk_candidates(scores: np.ndarray, k: int):
    """torch.topk(largest, sorted) over the flattened [beams*V] row; equal scores are taken lowest index first (what a stable
    descending sort gives).  A partition finds the k-th value first so that the 90k-entry row is not fully sorted every step."""
    if scores.size <= 8 * k:
        order = np.argsort(-scores, kind="stable")[:k]
        return scores[order], order
    thr = np.partition(scores, scores.size - k)[scores.size - k]          # the k-th largest value
    above = np.flatnonzero(scores > thr)
    equal = np.flatnonzero(scores == thr)
    cand = np.concatenate([above, equal[:k - above.size]])                # ties at the threshold: lowest indices
    order = cand[np.argsort(-scores[cand], kind="stable")]
    return scores[order], order


def _step_margin(row: np.ndarray, K: int, V: int, eos: int, kept_min: float) -> float:
    """How far one step's selection is from going the other way (see `margins` below): the gap between the K-th surviving
    non-eos continuation and the best one dropped, and for every beam's eos candidate its distance from the score that
    separates rank K-1 from rank K among the other candidates (an eos candidate counts only inside the first K)."""
    eos_pos = np.arange(K) * V + eos
    eos_scores = row[eos_pos].copy()
    others = row.copy()
    others[eos_pos] = -np.inf
    top = np.sort(np.partition(others, others.size - (K + 1))[others.size - (K + 1):])[::-1]     # K+1 best non-eos, descending
    assert top[K - 1] == kept_min
    gap = float(top[K - 1] - top[K])
    for j in range(K):
        if np.isfinite(eos_scores[j]):
            rest = np.sort(np.concatenate([top[:K], np.delete(eos_scores, j)]))[::-1]               # candidates that can outrank it
            gap = min(gap, abs(float(eos_scores[j]) - float(rest[K - 1])))
    return gap


def beam_search_from_logits(step_logits, batch: int, prompt: list, num_beams: int = 3, max_length: int = 20,
                            min_length: int = 5, eos: int = 102, pad: int = 0, length_penalty: float = 1.0,
                            margins: list | None = None, rules: str = "v4.15"):
    """`step_logits(input_ids [batch*beams, n] (np.int64), beam_idx or None) -> fp32 logits [batch*beams, V]` is called
    once per step with the full current sequences and the beam reordering of the previous step.
    Returns (tokens list per frame incl. a trailing eos when it fits, scores, state trace per step).
    `margins` (a list, filled with one float per frame): the smallest gap, in sum-of-log-probability units, by which any
    comparison that shaped this frame's result was decided — which K continuations survive a step, whether an eos
    candidate ranks inside the first K, every BeamHypotheses comparison, the stopping test and the final choice.  A path
    whose candidate scores differ from this one's by less than half that gap necessarily returns the same tokens."""
    assert rules in RULES
    v5, P = rules == "v5", len(prompt)
    margin = np.full(batch, np.inf)
    K = num_beams
    ids = np.tile(np.asarray(prompt, dtype=np.int64)[None], (batch * K, 1))
    beam_scores = np.zeros((batch, K), dtype=np.float32)
    beam_scores[:, 1:] = -1e9
    beam_scores = beam_scores.reshape(-1)
    hyps = [_BeamHyps(K, length_penalty) for _ in range(batch)]
    done = [False] * batch
    beam_idx = None
    trace = []
    while True:
        cur_len = ids.shape[1]
        logits = np.asarray(step_logits(ids, beam_idx), dtype=np.float32)
        lp = torch.log_softmax(torch.from_numpy(logits), dim=-1).numpy()
        if cur_len < min_length:
            lp[:, eos] = -np.inf                                                          # MinLengthLogitsProcessor
        V = lp.shape[1]
        scores = (lp + beam_scores[:, None]).reshape(batch, K * V)
        nb_scores = np.zeros((batch, K), dtype=np.float32)
        nb_tokens = np.zeros((batch, K), dtype=np.int64)
        nb_idx = np.zeros((batch, K), dtype=np.int64)
        for b in range(batch):
            if done[b]:
                nb_tokens[b, :] = pad
                nb_idx[b, :] = b * K          # v4.15 writes 0 here; the rows of a finished frame are never read again
                continue
            cs, ci = topk_candidates(scores[b], 2 * K)
            slot = 0
            for rank in range(2 * K):
                tok, src = int(ci[rank] % V), int(ci[rank] // V)
                row = b * K + src
                if tok == eos:
                    if rank >= K:
                        continue
                    margin[b] = min(margin[b], hyps[b].add(ids[row].tolist(), float(cs[rank]), cur_len + 1 - P if v5 else cur_len))
                else:
                    nb_scores[b, slot], nb_tokens[b, slot], nb_idx[b, slot] = cs[rank], tok, row
                    slot += 1
                if slot == K:
                    break
            assert slot == K
            best_open, open_len = (float(nb_scores[b, 0]), cur_len + 1 - P) if v5 else (float(cs.max()), cur_len)
            if margins is not None:
                margin[b] = min(margin[b], _step_margin(scores[b], K, V, eos, float(nb_scores[b, K - 1])))
                if len(hyps[b].beams) >= K:
                    margin[b] = min(margin[b], abs(hyps[b].worst_score * open_len ** length_penalty - best_open))
            done[b] = done[b] or hyps[b].is_done(best_open, open_len)
        beam_scores = nb_scores.reshape(-1)
        beam_idx = nb_idx.reshape(-1)
        ids = np.concatenate([ids[beam_idx], nb_tokens.reshape(-1, 1)], axis=1)
        trace.append(dict(beam_scores=beam_scores.copy(), beam_idx=beam_idx.copy(), tokens=nb_tokens.reshape(-1).copy(),
                          done=list(done)))
        if all(done) or ids.shape[1] >= max_length:
            break
    out_tokens, out_scores = [], []
    for b in range(batch):                                                                # BeamSearchScorer.finalize
        if not done[b]:
            for j in range(K):
                margin[b] = min(margin[b], hyps[b].add(ids[b * K + j].tolist(), float(beam_scores[b * K + j]),
                                                       max_length - P if v5 else None))
        ranked = sorted(hyps[b].beams, key=lambda x: x[0])
        if len(ranked) > 1:
            margin[b] = min(margin[b], (ranked[-1][0] - ranked[-2][0]) * min(ranked[-1][2], ranked[-2][2]) ** length_penalty)
        best = ranked.pop()
        seq = list(best[1])
        if len(seq) < max_length:
            seq.append(eos)
        out_tokens.append(seq)
        out_scores.append(best[0])
    if margins is not None:
        margins[:] = [float(x) for x in margin]
    return out_tokens, out_scores, trace


# ----------------------------------------------------------------------------------------------------------------------
# Nucleus sampling — transformers v4.15.0 `GenerationMixin.sample` with the processors / warpers `generate(do_sample=True,
# top_p=..., repetition_penalty=1.1, min_length=...)` builds (blip.py:141-148): RepetitionPenaltyLogitsProcessor,
# MinLengthLogitsProcessor, then TopKLogitsWarper (the PretrainedConfig default top_k = 50 BLIP inherits) and
# TopPLogitsWarper.  The four processors are pinned against the installed transformers' own classes
# (tests/test_beam_search_pin.py::test_sampling_processors_equal_installed_transformers) and the whole loop against the
# installed `generate(do_sample=True, ...)` (::test_sampling_loop_equals_installed_transformers_generate, 64/64 sequences); the
# draw itself is an inverse-CDF lookup with caller-supplied uniform numbers, because torch.multinomial's stream cannot be
# reproduced by another implementation — that test swaps torch.multinomial for the same lookup, and what is compared with the
# native path is the same lookup on the same numbers.
# ----------------------------------------------------------------------------------------------------------------------
def process_sampling_scores(logits: np.ndarray, seq: np.ndarray, cur_len: int, min_length: int, eos: int, top_k: int = 50,
                            top_p: float = 0.9, repetition_penalty: float = 1.1):
    """One row: fp32 logits [V] and the tokens so far -> (token ids kept, in descending score order, ties by ascending id;
    their processed scores).  v4.15 logits_process.py: RepetitionPenalty :146-158, MinLength :96-107, TopK :223-241 (keeps ties
    at the k-th value), TopP :173-206 (a token goes once the probability mass before it exceeds top_p)."""
    x = np.asarray(logits, dtype=np.float32).copy()
    seen = np.unique(np.asarray(seq, dtype=np.int64))
    x[seen] = np.where(x[seen] < 0, x[seen] * np.float32(repetition_penalty), x[seen] / np.float32(repetition_penalty))
    if cur_len < min_length:
        x[eos] = -np.inf
    order = np.lexsort((np.arange(x.size), -x))                       # score descending, id ascending among equals
    order = order[np.isfinite(x[order])]
    if order.size > top_k:
        kth = x[order[top_k - 1]]
        order = order[x[order] >= kth]
    v = x[order]
    e = np.exp(v - v[0], dtype=np.float32)
    total = np.float32(0.0)
    for t in e:                                                        # the same sequential fp32 sums as the kernel
        total = np.float32(total + t)
    keep, cum = 0, np.float32(0.0)
    for j in range(order.size):
        if j > 0 and np.float32(cum / total) > np.float32(top_p):
            break
        cum = np.float32(cum + e[j])
        keep += 1
    return order[:keep], v[:keep]


def sample_from_logits(step_logits, batch: int, prompt: list, uniforms: np.ndarray, max_length: int = 20, min_length: int = 5,
                       eos: int = 102, pad: int = 0, top_k: int = 50, top_p: float = 0.9, repetition_penalty: float = 1.1):
    """`step_logits(input_ids [batch, n], None) -> fp32 logits [batch, V]` as in beam_search_from_logits.  uniforms
    [max_length - len(prompt), batch].  Returns (token lists incl. eos, sums of the drawn tokens' log-probabilities)."""
    ids = np.tile(np.asarray(prompt, dtype=np.int64)[None], (batch, 1))
    unfinished = np.ones(batch, dtype=bool)
    logp = np.zeros(batch, dtype=np.float32)
    step = 0
    while True:
        cur_len = ids.shape[1]
        logits = np.asarray(step_logits(ids, None), dtype=np.float32)
        nxt = np.full(batch, pad, dtype=np.int64)
        for b in range(batch):
            if not unfinished[b]:
                continue
            toks, v = process_sampling_scores(logits[b], ids[b], cur_len, min_length, eos, top_k, top_p, repetition_penalty)
            e = np.exp(v - v[0], dtype=np.float32)
            mass = np.float32(0.0)
            for t in e:
                mass = np.float32(mass + t)
            target = np.float32(np.float32(uniforms[step, b]) * mass)
            pick, run = len(toks) - 1, np.float32(0.0)
            for j in range(len(toks)):
                run = np.float32(run + e[j])
                if run > target:
                    pick = j
                    break
            nxt[b] = toks[pick]
            logp[b] += np.float32((v[pick] - v[0]) - np.log(mass, dtype=np.float32))
        ids = np.concatenate([ids, nxt[:, None]], axis=1)
        unfinished &= nxt != eos
        step += 1
        if not unfinished.any() or ids.shape[1] >= max_length:
            break
    out = []
    for b in range(batch):
        seq = ids[b].tolist()
        if eos in seq[len(prompt):]:
            seq = seq[:len(prompt) + seq[len(prompt):].index(eos) + 1]
        out.append(seq)
    return out, logp


@torch.no_grad()
def generate_sample(sd: dict, image_embeds: torch.Tensor, prompt: list, H: int, depth: int, uniforms, pre: str = "text_decoder.",
                    max_length: int = 20, min_length: int = 5, eos: int = 102, pad: int = 0, top_k: int = 50, top_p: float = 0.9,
                    repetition_penalty: float = 1.1, operand_dtype=None):
    """BLIP_Decoder.generate(sample=True), blip.py:139-148, from the image tokens on (cached decoder, one sequence per frame)."""
    if operand_dtype is not None:
        with emulate(operand_dtype):
            return generate_sample(sd, image_embeds, prompt, H, depth, uniforms, pre, max_length, min_length, eos, pad, top_k, top_p,
                                   repetition_penalty, None)
    state = {"past": None}

    def step(ids, _):
        inp = torch.from_numpy(ids if state["past"] is None else ids[:, -1:])
        logits, state["past"] = decoder_logits(sd, pre, inp, image_embeds, H, depth, state["past"])
        return logits[:, -1, :].float().numpy()

    return sample_from_logits(step, image_embeds.shape[0], prompt, np.asarray(uniforms, dtype=np.float32), max_length, min_length, eos,
                              pad, top_k, top_p, repetition_penalty)


@torch.no_grad()
def generate(sd: dict, image_embeds: torch.Tensor, prompt: list, H: int, depth: int, pre: str = "text_decoder.",
             num_beams: int = 3, max_length: int = 20, min_length: int = 5, eos: int = 102, pad: int = 0,
             length_penalty: float = 1.0, device=None, operand_dtype=None, margins: list | None = None):
    """BLIP_Decoder.generate(sample=False), blip.py:127-167, from the image tokens on: repeat_interleave the image
    tokens over the beams (:130), run the cached decoder (prepare_inputs_for_generation / _reorder_cache, med.py:929-955)."""
    if operand_dtype is not None:
        with emulate(operand_dtype):
            return generate(sd, image_embeds, prompt, H, depth, pre, num_beams, max_length, min_length, eos, pad, length_penalty,
                            device, None, margins)
    B = image_embeds.shape[0]
    dev = image_embeds.device if device is None else device   # `device`: where sd / image_embeds live (bench.py's eager-GPU baseline)
    enc = image_embeds.repeat_interleave(num_beams, dim=0)
    state = {"past": None}

    def step(ids, beam_idx):
        past = state["past"]
        if past is None:
            inp = torch.from_numpy(ids).to(dev)
        else:
            idx = torch.from_numpy(beam_idx).to(dev)
            past = [(k.index_select(0, idx), v.index_select(0, idx)) for k, v in past]
            inp = torch.from_numpy(ids[:, -1:]).to(dev)
        logits, presents = decoder_logits(sd, pre, inp, enc, H, depth, past)
        state["past"] = presents
        return logits[:, -1, :].float().cpu().numpy()

    return beam_search_from_logits(step, B, prompt, num_beams, max_length, min_length, eos, pad, length_penalty, margins)
