"""The CPU oracle against the committed fixtures, which are outputs of the reference itself
(oracle/make_golden.py: unmodified models/vit.py, transformers' CLIPModel, the reference's own top-k and
aggregation lines).  This is what pins the oracle."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import clip_oracle, clip_text_oracle, preprocess_oracle, tokenization_oracle, vit_oracle, weights as W

# Same ATen ops in the same order as the reference => equal up to thread-count-dependent reduction order.
TOL = 2e-4


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


@pytest.mark.parametrize("vit,size,batch,fname", [("tiny", 32, 2, "vit_tiny.npz"), ("large", 224, 1, "vit_large_224.npz"),
                                                  ("base", 384, 1, "vit_base_384.npz")])
def test_vit_oracle_matches_reference_fixture(golden_dir, vit, size, batch, fname):
    g = _load(golden_dir, fname)
    sd = W.vit_state_dict(vit, size, seed=0)
    x = W.frames(batch, size, seed=0)
    out, blocks = vit_oracle.vit_forward(sd, x, W.VIT_CONFIGS[vit][2], return_blocks=True)
    tok = g["tokens"]
    assert np.abs(out[:, tok].numpy() - g["out"]).max() < TOL
    assert abs(out.abs().mean().item() - float(g["out_mean_abs"])) < 1e-4
    for key in g.files:
        if key.startswith("block"):
            i = int(key[5:])
            ref = g[key]
            assert np.abs(blocks[i][:, tok].numpy() - ref).max() < TOL * max(1.0, np.abs(ref).max())


def test_vit_oracle_fp64_agrees_with_fp32(golden_dir):
    sd = W.vit_state_dict("tiny", 32, seed=0)
    x = W.frames(2, 32, seed=0)
    a = vit_oracle.vit_forward(sd, x, 2)
    b = vit_oracle.vit_forward(sd, x, 2, dtype=torch.float64)
    assert (a.double() - b).abs().max() < 1e-4


@pytest.mark.parametrize("name,batch,fname", [("tiny", 2, "clip_tiny.npz"), ("large14", 1, "clip_large14.npz")])
def test_clip_oracle_matches_transformers_fixture(golden_dir, name, batch, fname):
    g = _load(golden_dir, fname)
    c = W.CLIP_CONFIGS[name]
    sd = W.clip_vision_state_dict(name, seed=0)
    x = W.frames(batch, c["image_size"], seed=1)
    emb, hidden = clip_oracle.clip_vision_forward(sd, x, c["num_attention_heads"])
    assert np.abs(emb.numpy() - g["image_embeds"]).max() < 1e-5
    ref = g["last_hidden"]
    assert np.abs(hidden[:, g["tokens"]].numpy() - ref).max() < TOL * max(1.0, np.abs(ref).max())
    assert np.allclose(np.linalg.norm(emb.numpy(), axis=1), 1.0, atol=1e-5)


@pytest.mark.parametrize("name,batch,seq,fname", [("tiny", 5, 12, "clip_text_tiny.npz"), ("large14", 3, 20, "clip_text_large14.npz")])
def test_clip_text_oracle_matches_transformers_fixture(golden_dir, name, batch, seq, fname):
    """The fixture was produced WITH the tokenizer-style attention_mask; the oracle applies the causal mask only —
    padding after the EOS token cannot reach the pooled EOS row."""
    g = _load(golden_dir, fname)
    c = W.CLIP_TEXT_CONFIGS[name]
    emb, _ = clip_text_oracle.clip_text_forward(W.clip_text_state_dict(name, seed=0), W.token_ids(name, batch, seq, seed=0),
                                                c["num_attention_heads"], c["eos_token_id"])
    assert np.abs(emb.numpy() - g["text_embeds"]).max() < 1e-5
    assert np.allclose(np.linalg.norm(emb.numpy(), axis=1), 1.0, atol=1e-5)


def test_preprocess_oracle_is_bit_identical_to_pil_fixture(golden_dir):
    """Index/byte work: the restatement of Pillow's fixed-point resize + ToTensor + Normalize must reproduce the outputs of
    the reference's own process_frame lines exactly (SHA-256 of the float32 bytes)."""
    import hashlib
    for case in json.load(open(os.path.join(golden_dir, "preprocess.json"))):
        frames = W.u8_frames(2, case["H"], case["W"], seed=case["H"] + case["W"]).numpy()
        out = np.stack([preprocess_oracle.process_frame(f, case["S"]) for f in frames])
        assert hashlib.sha256(np.ascontiguousarray(out).tobytes()).hexdigest() == case["sha256"], case["H"]
        st = max(1, case["S"] // 4)
        assert np.array_equal(out[:, :, ::st, ::st], np.asarray(case["sample"], dtype=np.float32))


def test_clip_preprocess_oracle_is_bit_identical_to_hf_processor_fixture(golden_dir):
    """run_visual_tokenization.py:138-140: the restatement of transformers' CLIPImageProcessor (shortest edge 224 bicubic,
    centre crop, rescale, normalise) reproduces the processor's own pixel_values exactly (SHA-256 of the float32 bytes)."""
    import hashlib
    for case in json.load(open(os.path.join(golden_dir, "clip_preprocess.json"))):
        frames = W.u8_frames(2, case["H"], case["W"], seed=case["H"] * 7 + case["W"]).numpy()
        out = np.stack([preprocess_oracle.clip_process_frame(f, case["S"]) for f in frames])
        assert hashlib.sha256(np.ascontiguousarray(out).tobytes()).hexdigest() == case["sha256"], (case["H"], case["W"])
        assert np.array_equal(out[:, :, ::56, ::56], np.asarray(case["sample"], dtype=np.float32))


def test_clip_preprocess_oracle_vs_live_hf_processor():
    import warnings
    m = pytest.importorskip("transformers.models.clip.image_processing_pil_clip")
    warnings.filterwarnings("ignore")
    proc = m.CLIPImageProcessorPil()
    for (H, Wd) in [(97, 131), (300, 200)]:
        frame = W.u8_frames(1, H, Wd, seed=2).numpy()[0]
        want = proc(images=[frame], return_tensors="np")["pixel_values"][0]
        assert np.array_equal(want, preprocess_oracle.clip_process_frame(frame, 224))


def test_preprocess_oracle_vs_live_pil():
    torchvision = pytest.importorskip("torchvision")
    from torchvision import transforms
    from torchvision.transforms.functional import InterpolationMode
    frame = W.u8_frames(1, 97, 131, seed=1).numpy()[0]
    t = transforms.Compose([transforms.ToPILImage(), transforms.Resize((64, 64), interpolation=InterpolationMode.BICUBIC),
                            transforms.ToTensor(), transforms.Normalize(preprocess_oracle.MEAN, preprocess_oracle.STD)])
    assert np.array_equal(t(frame).numpy(), preprocess_oracle.process_frame(frame, 64))


def test_topk_oracle_matches_reference_lines(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "tokenization.json")))
    img = W.unit_rows(g["F"], g["D"], seed=g["img_seed"]).numpy()
    bank = W.unit_rows(g["T"], g["D"], seed=g["bank_seed"]).numpy()
    _, idx = tokenization_oracle.sim_topk(img, bank, g["k"])
    assert idx.reshape(g["F"] // 8, 8, g["k"]).tolist() == g["topk_indices"]


def test_aggregation_oracle_matches_reference_function(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "tokenization.json")))
    assert len(g["aggregate_cases"]) >= 5
    for case in g["aggregate_cases"]:
        assert tokenization_oracle.aggregate_frame_tokens(case["frame_tokens"]) == case["aggregated"]


def test_sharding_oracle_matches_reference_formula(golden_dir):
    for case in json.load(open(os.path.join(golden_dir, "sharding.json"))):
        for rank, sl in enumerate(case["slices"]):
            s, e = tokenization_oracle.shard_bounds(case["n"], case["world"], rank)
            assert ([s, e] if e > s else None) == sl


def test_flops_per_frame_matches_baseline():
    assert abs(vit_oracle.flops_per_frame(1024, 24, 197) / 1e9 - 123.107) < 1e-2          # BLIP ViT-L/16@224
    assert abs(vit_oracle.flops_per_frame(768, 12, 577) / 1e9 - 110.967) < 1e-2           # ViT-B/16@384
    assert abs(vit_oracle.flops_per_frame(1024, 24, 257, patch_size=14, proj_dim=768) / 1e9 - 162.026) < 1e-2
