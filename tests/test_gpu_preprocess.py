"""Frame pre-processing on a B200 (vidil_preprocess_frames): byte / integer work, so the bar is bit-exactness against the
oracle and against the committed outputs of the reference's own process_frame lines (PIL + torchvision)."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import preprocess_oracle, weights as W
from vidil_b200 import preprocess

pytestmark = pytest.mark.gpu


def test_preprocess_matches_pil_fixture_bit_for_bit(cuda, golden_dir):
    for case in json.load(open(os.path.join(golden_dir, "preprocess.json"))):
        frames = W.u8_frames(2, case["H"], case["W"], seed=case["H"] + case["W"])
        out = preprocess.process_frames(frames.to(cuda), case["S"]).cpu().numpy()
        assert out.shape == (2, 3, case["S"], case["S"]) and out.dtype == np.float32
        assert hashlib.sha256(np.ascontiguousarray(out).tobytes()).hexdigest() == case["sha256"], (case["H"], case["W"])


@pytest.mark.parametrize("H,Wd,S,B", [(240, 320, 224, 5), (33, 17, 224, 2), (1080, 1920, 224, 1), (224, 224, 384, 3), (1, 1, 16, 2)])
def test_preprocess_equals_oracle(cuda, H, Wd, S, B):
    frames = W.u8_frames(B, H, Wd, seed=3)
    got = preprocess.process_frames(frames.to(cuda), S).cpu().numpy()
    want = np.stack([preprocess_oracle.process_frame(f, S) for f in frames.numpy()])
    assert np.array_equal(got, want)


def test_process_frame_signature_and_batch(cuda):
    """The reference's call: process_frame(frame, config, device) per frame (run_video_CapFilt.py:161); a 256-frame batch
    gives the same rows as 256 single calls."""
    frames = W.u8_frames(256, 120, 160, seed=9)
    batch = preprocess.process_frames(frames.to(cuda), 224)
    one = preprocess.process_frame(frames[17].numpy(), {"image_size": 224}, cuda)
    assert tuple(one.shape) == (3, 224, 224) and one.is_cuda and torch.equal(one, batch[17])
    assert torch.equal(preprocess.process_frames(frames[:0].to(cuda), 224), torch.empty(0, 3, 224, 224, device=cuda))
    with pytest.raises(RuntimeError, match="CUDA"):
        preprocess.process_frames(frames, 224)
    with pytest.raises(RuntimeError, match="uint8"):
        preprocess.process_frames(frames.float().to(cuda), 224)


def test_preprocessed_frames_feed_the_encoder(cuda):
    """uint8 frames -> process_frames -> ViT: the chain the CapFilt driver runs per video."""
    from vidil_b200.vision_transformer import VisionTransformer
    m = VisionTransformer(img_size=32, patch_size=16, embed_dim=128, depth=2, num_heads=2, compute_dtype="fp16")
    m.load_state_dict(W.vit_state_dict("tiny", 32, seed=0))
    m = m.to(cuda).eval()
    frames = W.u8_frames(4, 60, 80, seed=2)
    x = preprocess.process_frames(frames.to(cuda), 32)
    from oracle import vit_oracle
    ref = vit_oracle.vit_forward(W.vit_state_dict("tiny", 32, seed=0), torch.from_numpy(
        np.stack([preprocess_oracle.process_frame(f, 32) for f in frames.numpy()])), 2)
    assert (m(x).cpu() - ref).abs().max() < 2e-2


def test_encode_u8_stream_equals_the_step_by_step_path(cuda):
    from vidil_b200.vision_transformer import VisionTransformer
    m = VisionTransformer(img_size=32, patch_size=16, embed_dim=128, depth=2, num_heads=2, compute_dtype="bf16")
    m.load_state_dict(W.vit_state_dict("tiny", 32, seed=0))
    m = m.to(cuda).eval()
    batches = [W.u8_frames(3, 48, 64, seed=40 + i).pin_memory() for i in range(5)]
    want = [m(preprocess.process_frames(b.to(cuda), 32)).cpu() for b in batches]
    got = [o.clone() for o in preprocess.encode_u8_stream(m, iter(batches), 32)]
    assert len(got) == 5 and all(torch.equal(a, b) for a, b in zip(got, want))


# ---- CLIP side: transformers' CLIPImageProcessor (run_visual_tokenization.py:138-140) ----------------------------------
def test_clip_preprocess_matches_hf_processor_fixture_bit_for_bit(cuda, golden_dir):
    """Byte / integer work: SHA-256 of the float32 pixel_values equals the digest of what transformers' CLIPImageProcessor (PIL
    backend) produced for the same seeded frames (oracle/make_golden.py), 10 geometries: landscape, portrait, square,
    up-scaling, 1080p, and a short side one pixel off the target."""
    for case in json.load(open(os.path.join(golden_dir, "clip_preprocess.json"))):
        frames = W.u8_frames(2, case["H"], case["W"], seed=case["H"] * 7 + case["W"])
        out = preprocess.clip_process_frames(frames.to(cuda), case["S"]).cpu().numpy()
        assert out.shape == (2, 3, 224, 224) and out.dtype == np.float32
        assert hashlib.sha256(np.ascontiguousarray(out).tobytes()).hexdigest() == case["sha256"], (case["H"], case["W"])


@pytest.mark.parametrize("H,Wd,S,B", [(240, 320, 224, 5), (320, 240, 224, 3), (33, 17, 224, 2), (17, 33, 64, 2), (64, 64, 64, 1),
                                      (2160, 3840, 224, 1), (1, 5, 16, 2)])
def test_clip_preprocess_equals_oracle(cuda, H, Wd, S, B):
    frames = W.u8_frames(B, H, Wd, seed=4)
    got = preprocess.clip_process_frames(frames.to(cuda), S).cpu().numpy()
    want = np.stack([preprocess_oracle.clip_process_frame(f, S) for f in frames.numpy()])
    assert np.array_equal(got, want)


def test_clip_processor_dropin_call_surface(cuda):
    """`processor(images=frames, return_tensors="pt")` with a video's frames as numpy arrays / PIL images / one uint8 tensor, and
    frames of different geometries in one call (predict_video batches across videos): rows in the caller's order."""
    proc = preprocess.VidilCLIPProcessor(cuda)
    a = W.u8_frames(3, 120, 160, seed=1).numpy()
    b = W.u8_frames(2, 200, 150, seed=2).numpy()
    want_a = np.stack([preprocess_oracle.clip_process_frame(f) for f in a])
    want_b = np.stack([preprocess_oracle.clip_process_frame(f) for f in b])
    out = proc(images=[f for f in a], return_tensors="pt")["pixel_values"]
    assert out.is_cuda and np.array_equal(out.cpu().numpy(), want_a)
    assert np.array_equal(proc(images=torch.from_numpy(a))["pixel_values"].cpu().numpy(), want_a)
    from PIL import Image
    assert np.array_equal(proc(images=[Image.fromarray(f) for f in a])["pixel_values"].cpu().numpy(), want_a)
    mixed = proc(images=[a[0], b[0], a[1], b[1], a[2]])["pixel_values"].cpu().numpy()
    assert np.array_equal(mixed, np.stack([want_a[0], want_b[0], want_a[1], want_b[1], want_a[2]]))
    with pytest.raises(RuntimeError, match="without a tokenizer"):
        proc(text=["a photo"])


def test_clip_u8_stream_equals_preprocess_then_tower(cuda):
    """encode_u8_stream on the CLIP tower: uint8 batches in, image_embeds out, identical to pre-processing then encoding."""
    from vidil_b200.clip import CLIPVisionB200
    c = W.CLIP_CONFIGS["tiny"]
    m = CLIPVisionB200(**c, compute_dtype="bf16")
    m.load_state_dict(W.clip_vision_state_dict("tiny", seed=0))
    m = m.to(cuda).eval()
    batches = [W.u8_frames(b, 48, 64, seed=10 + i).pin_memory() for i, b in enumerate([4, 4, 2])]
    want = [m(preprocess.clip_process_frames(b.to(cuda), c["image_size"])).cpu() for b in batches]
    got = [o.clone() for o in m.encode_u8_stream(iter(batches))]
    assert len(got) == 3 and all(torch.equal(a, b) for a, b in zip(got, want))
