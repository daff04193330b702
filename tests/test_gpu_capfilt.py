"""The CapFilt driver end to end on a B200 with the native models (ViT-B/16 @224 + BERT-base decoder / ITM encoder, random
weights) and a stand-in tokenizer: decoded uint8 frames -> pre-processing -> captions -> de-duplication -> ITM filter."""
import copy

import numpy as np
import pytest
import torch

from oracle import weights as W
from vidil_b200 import capfilt
from vidil_b200.blip import BLIP_Decoder, BLIP_ITM

pytestmark = pytest.mark.gpu

WORDS = {1037: "a", 3861: "picture", 1997: "of"}
IDS = {v: k for k, v in WORDS.items()}


class StandInTokenizer:
    """The BertTokenizer surface blip.py / blip_itm.py use, over a vocabulary of 'w<id>' words (bert-base-uncased is not on disk)."""
    bos_token_id, sep_token_id, pad_token_id, cls_token_id, enc_token_id = 30522, 102, 0, 101, 30523
    special = {0, 101, 102, 30522, 30523}

    def _encode(self, text):
        return [IDS[w] if w in IDS else int(w[1:]) for w in text.split()]

    def __call__(self, text, padding=None, truncation=True, max_length=35, return_tensors=None):
        if isinstance(text, str):
            return type("Enc", (), {"input_ids": [101] + self._encode(text) + [102]})()
        ids = torch.zeros(len(text), max_length, dtype=torch.long)
        mask = torch.zeros(len(text), max_length, dtype=torch.long)
        for i, t in enumerate(text):
            row = ([101] + self._encode(t))[:max_length - 1] + [102]
            ids[i, :len(row)] = torch.tensor(row)
            mask[i, :len(row)] = 1
        return type("Enc", (), {"input_ids": ids, "attention_mask": mask})()

    def decode(self, ids, skip_special_tokens=True):
        return " ".join(WORDS.get(int(t), f"w{int(t)}") for t in ids if not (skip_special_tokens and int(t) in self.special))


def _models(cuda):
    torch.manual_seed(0)
    tok = StandInTokenizer()
    cap = BLIP_Decoder(image_size=224, vit="base", tokenizer=tok, prompt="a picture of ")
    itm = BLIP_ITM(image_size=224, vit="base", tokenizer=tok)
    with torch.no_grad():
        for m in (cap, itm):
            for n, p in m.named_parameters():
                p.normal_(0.0, 0.03)
                if ("norm" in n or "LayerNorm" in n) and n.endswith("weight"):
                    p.add_(1.0)
        cap.text_decoder.cls.predictions.decoder.weight.normal_(0.0, 0.1)
        itm.itm_head.weight.normal_(0.0, 0.5)
    return cap.to(cuda).eval(), itm.to(cuda).eval()


def _loader(video_path, strategy, num_frm):
    if "broken" in video_path:
        raise IOError("cannot decode")
    return W.u8_frames(num_frm, 120, 160, seed=int(video_path.split("video")[1].split(".")[0])).numpy()


def _data(n):
    d = [{"video_path": f"/v/video{i}.mp4", "text": [f"w{2000 + i} w{3000 + i}"], "video_id": f"video{i}"} for i in range(n)]
    d[2]["video_path"] = "/v/broken2.mp4"
    return d


CONFIG = dict(image_size=224, vit="base", frm_sampling_strategy="uniform", num_frm_CapFilt=4, generation_mode="beam", threshold=0.5,
              filter_mode="max_filter", caption=True, filter=True, filter_generated_only=True, keep_original_caption=False,
              do_sentence_tokenization=False)


def test_capfilt_end_to_end_and_video_batching(cuda):
    cap, itm = _models(cuda)
    one, many = _data(6), _data(6)
    capfilt.CapFilt(one, CONFIG, cuda, captioner=cap, filterer=itm, frame_loader=_loader, video_batch=1)
    capfilt.CapFilt(many, CONFIG, cuda, captioner=cap, filterer=itm, frame_loader=_loader, video_batch=4)
    assert one == many
    assert "unfiltered_text" not in one[2]
    for item in one[:2] + one[3:]:
        caps = item["unfiltered_text"]
        assert 1 <= len(caps) <= 4 and len(set(caps)) == len(caps)
        assert all(isinstance(c, str) and c and not c.startswith("a picture of") for c in caps)
        assert set(item["text"]) <= set(caps)
    # threshold 1.0 can never be exceeded by a probability: everything is filtered out, the captions stay in unfiltered_text
    none = _data(3)
    capfilt.CapFilt(none, dict(CONFIG, threshold=1.0), cuda, captioner=cap, filterer=itm, frame_loader=_loader, video_batch=3)
    assert [i["text"] for i in none[:2]] == [[], []] and none[0]["unfiltered_text"] == one[0]["unfiltered_text"]

    # the reference's own way of filtering (run_video_CapFilt.py:107-126: one filterer call per caption) keeps the same captions
    item = one[0]
    frames = capfilt.process_frames(torch.as_tensor(_loader(item["video_path"], "uniform", 4)).to(cuda), 224)
    ref_kept = []
    for t in item["unfiltered_text"]:
        itm_output = itm(frames, [t for _ in range(frames.size()[0])], match_head="itm")
        score = torch.nn.functional.softmax(itm_output, dim=1)[:, 1].cpu().numpy()
        if np.max(score) > CONFIG["threshold"]:
            ref_kept.append(t)
    assert ref_kept == item["text"]
    # and the captions themselves are what the captioner gives for this video alone
    assert capfilt.dedup_exact(cap.generate(frames, sample=False, num_beams=3, max_length=20, min_length=5)) == item["unfiltered_text"]


def test_capfilt_generation_mode_sample(cuda):
    """generation_mode "sample" (run_video_CapFilt.py:103-104 -> blip.py:139-148): captions are strings after the prompt, the
    run is reproducible under torch.manual_seed like the reference's torch.multinomial draws, and a different seed gives
    different captions."""
    cap, itm = _models(cuda)
    frames = capfilt.process_frames(torch.as_tensor(_loader("/v/video1.mp4", "uniform", 4)).to(cuda), 224)
    torch.manual_seed(5)
    a = cap.generate(frames, sample=True, top_p=0.9, max_length=20, min_length=5)
    torch.manual_seed(5)
    b = cap.generate(frames, sample=True, top_p=0.9, max_length=20, min_length=5)
    torch.manual_seed(6)
    c = cap.generate(frames, sample=True, top_p=0.9, max_length=20, min_length=5)
    assert a == b and a != c
    assert len(a) == 4 and all(isinstance(s, str) and s and not s.startswith("a picture of") for s in a)
    assert all(1 <= len(s.split()) <= 16 for s in a)              # max_length 20 minus the 4 prompt tokens
    assert capfilt.caption_frames(cap, frames, mode="sample") is not None
    data = _data(3)                                               # video 2 cannot be decoded and is skipped
    torch.manual_seed(7)
    capfilt.CapFilt(data, dict(CONFIG, generation_mode="sample"), cuda, captioner=cap, filterer=itm, frame_loader=_loader, video_batch=2)
    assert all(1 <= len(d["unfiltered_text"]) <= 4 for d in data[:2]) and "unfiltered_text" not in data[2]
