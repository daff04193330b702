"""CPU restatement of the visual-tokenization tail of the reference — parity oracle.

TEST INFRASTRUCTURE (see oracle/vit_oracle.py for the import rule).

Follows run_visual_tokenization.py: the similarity matmul (:276), the per-frame host argsort (:298-308),
the frequency aggregation over frames (:173-187), and the per-rank partition / rank-ordered merge both
driver scripts use (run_visual_tokenization.py:427-431,453-457; run_video_CapFilt.py:237-241,272-283).
numpy only (this is index/byte work); fixtures in tests/golden/tokenization_*.json come from executing the
reference's own lines (oracle/make_golden.py).
"""
from __future__ import annotations

from collections import defaultdict

import numpy as np


def sim_topk(image_embeds: np.ndarray, text_embeds: np.ndarray, k: int):
    """sims = image_embeds @ text_embeds.T (fp32, :276); per row np.argsort(score)[::-1][:k] (:306).

    Returns (scores [F,k] fp32, indices [F,k] int64), best first.
    """
    sims = np.asarray(image_embeds, dtype=np.float32) @ np.asarray(text_embeds, dtype=np.float32).T
    idx = np.empty((sims.shape[0], k), dtype=np.int64)
    for f in range(sims.shape[0]):          # :302-306, one host argsort per frame
        idx[f] = np.argsort(sims[f])[::-1][:k]
    return np.take_along_axis(sims, idx, axis=1), idx


def frame_tokens(indices_by_key: dict, phrases_by_key: dict, num_videos: int, num_frm: int):
    """:268,300-308 — list per video of per-frame {key: [phrase, ...]} from [V*num_frm, k] index arrays."""
    out = [[defaultdict(list) for _ in range(num_frm)] for _ in range(num_videos)]
    for key, idx in indices_by_key.items():
        idx = np.asarray(idx).reshape(num_videos, num_frm, -1)       # :298 score.view(V, num_frm, -1)
        for v in range(num_videos):
            for f in range(num_frm):
                out[v][f][key] = [phrases_by_key[key][int(i)] for i in idx[v, f]]
    return out


def aggregate_frame_tokens(tokens: list) -> dict:
    """:173-187 — count phrases over frames (rank-outer, frame-inner insertion order), stable sort by count
    descending, keep the first `topk` (= number of object phrases per frame)."""
    keys = tokens[0].keys()
    aggregated = {key: [] for key in keys}
    topk = len(tokens[0]["objects"])
    num_frm = len(tokens)
    for key in keys:
        if tokens[0][key] == []:
            continue
        count = defaultdict(int)
        for j in range(topk):
            for i in range(num_frm):
                count[tokens[i][key][j]] += 1
        cands = sorted([(t, c) for t, c in count.items()], key=lambda x: x[1], reverse=True)
        aggregated[key] = [t for t, _ in cands[:topk]]
    return aggregated


def shard_bounds(n_items: int, world_size: int, rank: int):
    """step = n // world + 1; [rank*step, min(n, rank*step + step))  (run_visual_tokenization.py:429-431)."""
    step = n_items // world_size + 1
    start = rank * step
    return start, max(start, min(n_items, start + step))


def merge_rank_dicts(per_rank: list) -> dict:
    """dict.update in rank order (run_visual_tokenization.py:453-457)."""
    out = {}
    for d in per_rank:
        out.update(d)
    return out
