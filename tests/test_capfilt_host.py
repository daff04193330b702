"""Host logic of the CapFilt driver (vidil_b200/capfilt.py) against the reference's own functions, exec()'d from
/root/reference/run_video_CapFilt.py with fake models, and the N>1 path on gloo (world_size 2)."""
import copy
import json
import os
import socket
import sys
import zlib

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import reference_shims as rs
from vidil_b200 import capfilt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORDS = ["a man", "a dog runs", "two people talk", "a car on a road", "someone cooks food", "a band plays", "water"]


def _h(s):
    return zlib.crc32(s.encode())


class FakeTok:
    """BertTokenizer call surface used by blip_itm.py:46-47; ids carry a hash of the text so the fake model can tell captions apart."""

    def __call__(self, texts, padding=None, truncation=True, max_length=35, return_tensors="pt"):
        ids = torch.zeros(len(texts), max_length, dtype=torch.long)
        mask = torch.zeros(len(texts), max_length, dtype=torch.long)
        for i, t in enumerate(texts):
            n = min(len(t.split()) + 2, max_length)
            ids[i, 0] = 101
            ids[i, 1] = _h(t) % 30000
            mask[i, :n] = 1
        return type("Enc", (), {"input_ids": ids, "attention_mask": mask})()


def _logit(text_hash, frame_val):
    return ((text_hash % 13) - 6) * 0.5 + frame_val


class FakeFilterer:
    tokenizer = FakeTok()

    def eval(self):
        return self

    def to(self, device):
        return self

    def __call__(self, images, captions, match_head='itm'):                 # the reference's call, :111
        l1 = torch.tensor([_logit(_h(c) % 30000, float(images[i].mean())) for i, c in enumerate(captions)])
        return torch.stack([torch.zeros_like(l1), l1], dim=1)

    def forward_ids(self, images, ids, mask, frame_of_seq=None, seqs_per_frame=0):   # the batched call of vidil_b200.capfilt
        frame = (lambda p: int(frame_of_seq[p])) if frame_of_seq is not None else (lambda p: p // seqs_per_frame)
        l1 = torch.tensor([_logit(int(ids[p, 1]), float(images[frame(p)].mean())) for p in range(ids.shape[0])])
        return torch.stack([torch.zeros_like(l1), l1], dim=1)


class FakeCaptioner:
    def eval(self):
        return self

    def to(self, device):
        return self

    def generate(self, images, sample=False, num_beams=3, max_length=20, min_length=5, top_p=0.9):
        assert (sample, num_beams, max_length, min_length) == (False, 3, 20, 5)
        return [WORDS[int(abs(float(f.sum())) * 7) % len(WORDS)] for f in images]


def fake_loader(video_path, strategy, num_frm):
    if "broken" in video_path:
        return None if "none" in video_path else (_ for _ in ()).throw(IOError("cannot decode"))
    rng = np.random.default_rng(_h(video_path))
    return rng.integers(0, 255, size=(num_frm, 6, 8, 3), dtype=np.uint8)


def fake_processor(frames_u8, image_size):
    return frames_u8.float().permute(0, 3, 1, 2)[:, :, :4, :4] / 255.0 - 0.5


def _data():
    d = [{"video_path": f"/v/video{i}.mp4", "text": [f"original caption {i}\nsecond line", "x"], "video_id": f"video{i}"} for i in range(7)]
    d[3]["video_path"] = "/v/broken3.mp4"
    return d


CONFIGS = [
    dict(caption=True, filter=True, filter_generated_only=False, keep_original_caption=True, do_sentence_tokenization=False),
    dict(caption=True, filter=True, filter_generated_only=True, keep_original_caption=False, do_sentence_tokenization=False),
    dict(caption=True, filter=False, filter_generated_only=False, keep_original_caption=False, do_sentence_tokenization=False),
    dict(caption=False, filter=True, filter_generated_only=False, keep_original_caption=True, do_sentence_tokenization=False),
    dict(caption=True, filter=True, filter_generated_only=False, keep_original_caption=True),     # sentence splitter on
]


def _config(extra, mode="max_filter", threshold=0.5):
    c = dict(image_size=224, vit="base", frm_sampling_strategy="uniform", num_frm_CapFilt=4, generation_mode="beam",
             threshold=threshold, filter_mode=mode, caption_model_ckpt="", filterer_model_ckpt="")
    c.update(extra)
    return c


def _reference_capfilt():
    """caption_frames, filter_captions and CapFilt executed from the reference's own source lines, with the libraries this
    container lacks (decord, spaCy, the checkpoints) replaced by the same fakes the native side gets."""
    import textwrap
    with open(os.path.join(rs.REFERENCE_ROOT, "run_video_CapFilt.py")) as f:
        lines = f.readlines()
    src = "".join(lines[92:127]) + "".join(lines[138:205])       # :93-127 (caption_frames, filter_captions) and :139-205 (CapFilt)
    cap, filt = FakeCaptioner(), FakeFilterer()

    class _Sent:
        def __init__(self, t):
            self.text = t

    class _Doc:
        def __init__(self, t):
            self.sents = [_Sent(p) for p in t.split(". ")]

    ns = {"torch": torch, "np": np, "tqdm": lambda x: x,
          "spacy": type("S", (), {"load": staticmethod(lambda *a, **k: (lambda t: _Doc(t)))}),
          "blip_decoder": lambda **k: cap, "blip_itm": lambda **k: filt,
          "load_video_from_path_decord": fake_loader,
          "process_frame": lambda frm, config, device: fake_processor(torch.as_tensor(frm)[None], config["image_size"])[0]}
    exec(textwrap.dedent(src), ns)  # noqa: S102 - trusted local reference source
    return ns["CapFilt"], ns["filter_captions"]


splitter = lambda t: t.split(". ")  # noqa: E731


@pytest.mark.skipif(not rs.reference_available(), reason="reference tree not mounted")
@pytest.mark.parametrize("extra", CONFIGS)
@pytest.mark.parametrize("mode,threshold", [("max_filter", 0.5), ("avg_filter", 0.3)])
def test_capfilt_matches_reference_control_flow(extra, mode, threshold):
    ref_capfilt, _ = _reference_capfilt()
    cfg = _config(extra, mode, threshold)
    ref_data, data = _data(), _data()
    ref_capfilt(ref_data, cfg, "cpu")
    capfilt.CapFilt(data, cfg, "cpu", captioner=FakeCaptioner(), filterer=FakeFilterer(), frame_loader=fake_loader,
                    sentence_splitter=splitter, frame_processor=fake_processor)
    assert data == ref_data
    assert any(item["text"] != item.get("unfiltered_text") for item in data if "unfiltered_text" in item) or not cfg["filter"]
    assert "unfiltered_text" not in data[3]                          # the undecodable video is skipped, not fatal


@pytest.mark.skipif(not rs.reference_available(), reason="reference tree not mounted")
@pytest.mark.parametrize("extra", CONFIGS[:3])
@pytest.mark.parametrize("video_batch", [3, 16])
def test_capfilt_video_batches_equal_reference(extra, video_batch):
    """Several videos per ViT pass / beam search / ITM call: every item ends up exactly as the reference leaves it."""
    ref_capfilt, _ = _reference_capfilt()
    cfg = _config(extra)
    ref_data, data = _data(), _data()
    ref_capfilt(ref_data, cfg, "cpu")
    capfilt.CapFilt(data, cfg, "cpu", captioner=FakeCaptioner(), filterer=FakeFilterer(), frame_loader=fake_loader,
                    sentence_splitter=splitter, frame_processor=fake_processor, video_batch=video_batch)
    assert data == ref_data


@pytest.mark.skipif(not rs.reference_available(), reason="reference tree not mounted")
def test_filter_captions_matches_reference():
    _, ref_filter = _reference_capfilt()
    images = torch.randn(5, 3, 4, 4) * 0.3
    texts = [f"{w} number {i}" for i, w in enumerate(WORDS * 2)]
    for mode, thr in [("max_filter", 0.5), ("avg_filter", 0.5), ("max_filter", 0.9)]:
        assert capfilt.filter_captions(FakeFilterer(), images, texts, thr, mode) == ref_filter(FakeFilterer(), images, texts, thr, mode)
    assert capfilt.filter_captions(FakeFilterer(), images, [], 0.5) == []


def test_dedup_and_collect():
    assert capfilt.dedup_exact(["a", "b", "a", "c", "b"]) == ["a", "b", "c"]
    items = [{"video_id": "v0", "video_path": "p0", "text": ["x"], "unfiltered_text": ["x", "y"]},
             {"video_id": "v1", "video_path": "p1", "text": ["orig"]},                                  # never loaded
             {"video_id": "v2", "video_path": "p2", "text": [], "unfiltered_text": ["z"]}]              # everything filtered out
    f, u = capfilt.collect_rank_outputs(items)
    assert f == {"v0": ["x"]} and u == {"v0": ["x", "y"], "v2": ["z"]}


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist

    from vidil_b200 import distributed as vd
    vd.init_distributed_mode(backend="gloo")
    out = capfilt.run(_data(), _config(CONFIGS[0]), "cpu", output_dir=out_dir, captioner=FakeCaptioner(), filterer=FakeFilterer(),
                      frame_loader=fake_loader, sentence_splitter=splitter, frame_processor=fake_processor)
    assert (out is None) == (rank != 0)
    dist.barrier()
    dist.destroy_process_group()


def test_capfilt_run_world2_equals_single_process(tmp_path):
    single = capfilt.run(_data(), _config(CONFIGS[0]), "cpu", output_dir=str(tmp_path / "one"), captioner=FakeCaptioner(),
                         filterer=FakeFilterer(), frame_loader=fake_loader, sentence_splitter=splitter, frame_processor=fake_processor)
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path / "two")), nprocs=2, join=True)
    for name, ref in zip(("video_text_CapFilt.json", "video_text_Cap.json"), single):
        got = json.load(open(tmp_path / "two" / name))
        assert got == ref and list(got) == list(ref)
        assert open(tmp_path / "two" / name).read() == json.dumps(ref, indent=4)
    assert "video3" not in single[1] and len(single[1]) == 6


def test_filter_captions_batched_is_ragged_safe():
    """Any number of captions (zero included) and frames per video: the one-call version equals per-video filter_captions."""
    rng = np.random.default_rng(3)
    f = FakeFilterer()
    for trial in range(20):
        n_videos = int(rng.integers(1, 6))
        images = [torch.from_numpy(rng.standard_normal((int(rng.integers(1, 6)), 3, 4, 4)).astype(np.float32)) * 0.4
                  for _ in range(n_videos)]
        texts = [[f"{WORDS[int(rng.integers(0, len(WORDS)))]} v{v} c{c}" for c in range(int(rng.integers(0, 5)))] for v in range(n_videos)]
        for mode, thr in (("max_filter", 0.5), ("avg_filter", 0.45)):
            got = capfilt.filter_captions_batched(f, images, texts, thr, mode)
            assert got == [capfilt.filter_captions(f, im, tx, thr, mode) for im, tx in zip(images, texts)]
    assert capfilt.filter_captions_batched(f, [torch.zeros(2, 3, 4, 4)], [[]], 0.5) == [[]]
