"""Generate the committed parity fixtures under tests/golden/ by executing the reference itself.

TEST INFRASTRUCTURE, build-container only (needs /root/reference and, for the CLIP fixtures, the installed
transformers package).  Run from the repository root:

    python -m oracle.make_golden

What is produced (all inputs/weights are regenerated from seeds by oracle/weights.py, so only outputs are
stored):

  vit_tiny.npz            reference VisionTransformer (D=128, depth 2, 32x32 px), full output, 2 frames
  vit_large_224.npz       reference ViT-L/16 @224 (create_vit('large', 224)), 1 frame: CLS + every 4th token,
                          plus the residual stream after blocks 0, 11, 23 for the same tokens
  vit_base_384.npz        reference ViT-B/16 @384 (the shipped pipeline config), 1 frame: CLS + every 8th token
  clip_tiny.npz           transformers.CLIPModel vision path (D=128, 2 layers), full outputs, 2 frames
  clip_large14.npz        transformers.CLIPModel ViT-L/14 @224, 1 frame: image_embeds + CLS/every 8th token of
                          last_hidden_state
  clip_text_tiny.npz      transformers.CLIPModel text path (D=128, 2 layers, 12 tokens), 5 phrases: text_embeds + pooled
  clip_text_large14.npz   transformers.CLIPModel text tower of clip-vit-large-patch14's shape, 3 phrases of 20 tokens
  med_tiny.npz            reference med.py (D=128, 2 layers): teacher-forced decoder logits, cached one-token steps, ITM logits
  med_base_l.npz          reference med.py at BERT-base shape with ViT-L tokens (197 x 1024): the same, every 97th vocabulary entry
  clip_preprocess.json    transformers' CLIPImageProcessor (PIL backend) pixel_values on synthetic uint8 frames of 10 geometries
  preprocess.json         torchvision/PIL process_frame (run_video_CapFilt.py:128-137) on synthetic uint8 frames of 7 geometries:
                          SHA-256 of the float32 output + a 5x5 sample per channel
  tokenization.json       sim top-k indices computed with the reference's own lines (:276, :306), and
                          aggregate_frame_tokens executed from run_visual_tokenization.py:173-187
  sharding.json           per-rank slices from the reference's partition formula for several (n, world) pairs
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from . import reference_shims as rs
from . import weights as W

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _save(name, **arrays):
    path = os.path.join(GOLDEN_DIR, name)
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrays.items()})
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KB)")


@torch.no_grad()
def golden_vit(vit, image_size, batch, token_stride, tap_blocks, fname):
    sd = W.vit_state_dict(vit, image_size, seed=0)
    model = rs.build_reference_vit(vit, image_size, sd)
    x = W.frames(batch, image_size, seed=0)
    taps = {}
    hooks = []
    for i in tap_blocks:
        hooks.append(model.blocks[i].register_forward_hook(lambda m, a, out, i=i: taps.__setitem__(i, out.clone())))
    out = model(x)
    for h in hooks:
        h.remove()
    n_tok = out.shape[1]
    tok = np.arange(0, n_tok, token_stride)
    arrays = dict(tokens=tok, out=out[:, tok].numpy(), out_mean_abs=np.float32(out.abs().mean().item()),
                  out_std=np.float32(out.std().item()))
    for i, t in taps.items():
        arrays[f"block{i}"] = t[:, tok].numpy()
    _save(fname, **arrays)


def _clip_model(name):
    from transformers import CLIPConfig, CLIPModel
    c = W.CLIP_CONFIGS[name]
    vision = dict(hidden_size=c["hidden_size"], intermediate_size=c["intermediate_size"],
                  num_hidden_layers=c["num_hidden_layers"], num_attention_heads=c["num_attention_heads"],
                  image_size=c["image_size"], patch_size=c["patch_size"], layer_norm_eps=1e-5, hidden_act="quick_gelu")
    text = dict(hidden_size=64, intermediate_size=128, num_hidden_layers=1, num_attention_heads=1,
                max_position_embeddings=8, vocab_size=64, hidden_act="quick_gelu")  # unused by the image path
    cfg = CLIPConfig(text_config=text, vision_config=vision, projection_dim=c["projection_dim"])
    cfg._attn_implementation = "eager"
    model = CLIPModel(cfg).eval()
    sd = W.clip_vision_state_dict(name, seed=0)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith(("text_model.", "text_projection", "logit_scale")) for k in missing), missing
    return model


@torch.no_grad()
def golden_clip(name, batch, token_stride, fname):
    model = _clip_model(name)
    c = W.CLIP_CONFIGS[name]
    x = W.frames(batch, c["image_size"], seed=1)
    vout = model.vision_model(pixel_values=x)
    # exactly CLIPModel.forward's image branch (modeling_clip.py:916-923)
    emb = model.visual_projection(vout.pooler_output)
    emb = emb / emb.norm(p=2, dim=-1, keepdim=True)
    tok = np.arange(0, vout.last_hidden_state.shape[1], token_stride)
    _save(fname, tokens=tok, image_embeds=emb.numpy(), last_hidden=vout.last_hidden_state[:, tok].numpy())


@torch.no_grad()
def golden_clip_text(name, batch, seq_len, fname):
    """transformers' CLIPModel text branch (get_text_features + normalisation, as CLIPModel.forward does)."""
    from transformers import CLIPConfig, CLIPModel
    c = W.CLIP_TEXT_CONFIGS[name]
    text = dict(vocab_size=c["vocab_size"], max_position_embeddings=c["max_position_embeddings"], hidden_size=c["hidden_size"],
                intermediate_size=c["intermediate_size"], num_hidden_layers=c["num_hidden_layers"],
                num_attention_heads=c["num_attention_heads"], layer_norm_eps=1e-5, hidden_act="quick_gelu",
                eos_token_id=c["eos_token_id"], bos_token_id=c["eos_token_id"] - 1, pad_token_id=c["eos_token_id"])
    vision = dict(hidden_size=64, intermediate_size=128, num_hidden_layers=1, num_attention_heads=1, image_size=28, patch_size=14)
    cfg = CLIPConfig(text_config=text, vision_config=vision, projection_dim=c["projection_dim"])
    cfg._attn_implementation = "eager"
    model = CLIPModel(cfg).eval()
    missing, unexpected = model.load_state_dict(W.clip_text_state_dict(name, seed=0), strict=False)
    assert not unexpected and all(k.startswith(("vision_model.", "visual_projection", "logit_scale")) for k in missing), missing
    ids = W.token_ids(name, batch, seq_len, seed=0)
    mask = torch.ones_like(ids)
    for b in range(batch):                       # what the tokenizer would report: 1 up to and including the first EOS
        e = int((ids[b] == c["eos_token_id"]).int().argmax())
        mask[b, e + 1:] = 0
    out = model.text_model(input_ids=ids, attention_mask=mask)
    emb = model.text_projection(out.pooler_output)
    emb = emb / emb.norm(p=2, dim=-1, keepdim=True)
    _save(fname, text_embeds=emb.numpy(), pooled=out.pooler_output.numpy())


@torch.no_grad()
def golden_med(name, batch, seq_len, n_img, vocab_stride, fname):
    """Reference med.py executed on seeded inputs: teacher-forced decoder logits, one cached decode step, and the ITM
    logits of padded captions (the three ways run_video_CapFilt.py drives the text stack)."""
    c, sp = W.MED_CONFIGS[name], W.MED_SPECIAL[name]
    enc = W.image_tokens(batch, n_img, c["encoder_width"], seed=0)
    ones = torch.ones(batch, n_img, dtype=torch.long)
    sd = W.med_state_dict(name, "decoder", seed=0)
    dec, _ = rs.build_reference_med(name, "decoder", sd)
    ids, _ = W.caption_ids(name, batch, seq_len, seed=0, min_words=seq_len - 2)
    ids[:, 0] = sp["bos"]
    full = dec(ids, attention_mask=torch.ones_like(ids), encoder_hidden_states=enc, encoder_attention_mask=ones,
               return_dict=True, use_cache=True, is_decoder=True)
    # cached path exactly as generate() drives it: prompt first, then one token at a time (med.py:929-948)
    o = dec(ids[:, :4], attention_mask=torch.ones_like(ids[:, :4]), encoder_hidden_states=enc, encoder_attention_mask=ones,
            return_dict=True, use_cache=True, is_decoder=True)
    step_logits = []
    for t in range(4, seq_len):
        o = dec(ids[:, t:t + 1], attention_mask=torch.ones_like(ids[:, :t + 1]), encoder_hidden_states=enc,
                encoder_attention_mask=ones, past_key_values=o.past_key_values, return_dict=True, use_cache=True, is_decoder=True)
        step_logits.append(o.logits[:, 0])
    vs = np.arange(0, c["vocab_size"], vocab_stride)
    arrays = dict(vocab=vs, logits=full.logits[:, :, vs].numpy(), step_logits=torch.stack(step_logits, 1)[:, :, vs].numpy(),
                  logits_std=np.float32(full.logits.std().item()))
    sdi = W.med_state_dict(name, "itm", seed=0)
    enc_m, head = rs.build_reference_med(name, "itm", sdi)
    cap, mask = W.caption_ids(name, batch, seq_len, seed=1)
    cap[:, 0] = sp["enc"]
    out = enc_m(cap, attention_mask=mask, encoder_hidden_states=enc, encoder_attention_mask=ones, return_dict=True)
    arrays["itm_hidden_cls"] = out.last_hidden_state[:, 0].numpy()
    arrays["itm_logits"] = head(out.last_hidden_state[:, 0, :]).numpy()
    _save(fname, **arrays)


PREPROCESS_CASES = [(240, 320, 224), (360, 640, 384), (224, 224, 224), (100, 150, 224), (480, 270, 224), (7, 9, 32),
                    (720, 1280, 224)]


def golden_preprocess():
    """The reference's process_frame lines executed as they stand (run_video_CapFilt.py:128-137): torchvision ToPILImage ->
    Resize BICUBIC -> ToTensor -> Normalize.  Outputs are stored as SHA-256 digests (bit-exactness) plus a small sample."""
    import hashlib

    from torchvision import transforms
    from torchvision.transforms.functional import InterpolationMode
    cases = []
    for (H, Wd, S) in PREPROCESS_CASES:
        frames = W.u8_frames(2, H, Wd, seed=H + Wd).numpy()
        transform = transforms.Compose([
            transforms.ToPILImage(),
            transforms.Resize((S, S), interpolation=InterpolationMode.BICUBIC),
            transforms.ToTensor(),
            transforms.Normalize((0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711))])
        out = np.stack([transform(f).numpy() for f in frames])
        cases.append({"H": H, "W": Wd, "S": S, "sha256": hashlib.sha256(np.ascontiguousarray(out).tobytes()).hexdigest(),
                      "sample": out[:, :, ::max(1, S // 4), ::max(1, S // 4)].tolist()})
    with open(os.path.join(GOLDEN_DIR, "preprocess.json"), "w") as f:
        json.dump(cases, f)
    print("wrote preprocess.json")


CLIP_PREPROCESS_CASES = [(240, 320), (320, 240), (224, 224), (360, 640), (720, 1280), (225, 300), (100, 150), (500, 333),
                         (1080, 1920), (224, 225)]


def golden_clip_preprocess():
    """transformers' CLIPImageProcessor with its PIL backend — what `processor(images=frames, return_tensors="pt")` at
    run_visual_tokenization.py:138-140 runs — on synthetic uint8 frames: SHA-256 of the float32 pixel_values plus a sample."""
    import hashlib
    import warnings

    from transformers.models.clip.image_processing_pil_clip import CLIPImageProcessorPil
    warnings.filterwarnings("ignore")
    proc = CLIPImageProcessorPil()
    cases = []
    for (H, Wd) in CLIP_PREPROCESS_CASES:
        frames = W.u8_frames(2, H, Wd, seed=H * 7 + Wd).numpy()
        out = proc(images=[f for f in frames], return_tensors="np")["pixel_values"]
        assert out.shape == (2, 3, 224, 224) and out.dtype == np.float32
        cases.append({"H": H, "W": Wd, "S": 224, "sha256": hashlib.sha256(np.ascontiguousarray(out).tobytes()).hexdigest(),
                      "sample": out[:, :, ::56, ::56].tolist()})
    with open(os.path.join(GOLDEN_DIR, "clip_preprocess.json"), "w") as f:
        json.dump(cases, f)
    print("wrote clip_preprocess.json")


def golden_tokenization():
    # similarity + per-frame argsort exactly as run_visual_tokenization.py:276,298-306
    F_, T, D, k = 64, 1000, 768, 5
    img = W.unit_rows(F_, D, seed=0)
    bank = W.unit_rows(T, D, seed=1)
    sims_matrix = img @ bank.t()                                     # :276
    score = sims_matrix.view(F_ // 8, 8, -1).cpu().numpy()           # :298-299
    inds = [[np.argsort(score[j][f])[::-1][:k].tolist() for f in range(8)] for j in range(F_ // 8)]   # :306
    agg = rs.extract_reference_function("run_visual_tokenization.py", 173, 187, "aggregate_frame_tokens")
    cases = []
    rng = np.random.Generator(np.random.PCG64(7))
    vocab = [f"phrase {i}" for i in range(12)]
    for _ in range(6):
        num_frm, topk = int(rng.integers(2, 9)), int(rng.integers(1, 6))
        frames = []
        for _f in range(num_frm):
            frames.append({key: [vocab[int(i)] for i in rng.integers(0, len(vocab), size=topk)]
                           for key in ("objects", "attributes", "scenes", "verbs")})
        frames[0]["verbs"] = [] if rng.integers(0, 2) else frames[0]["verbs"]
        if frames[0]["verbs"] == []:
            for fr in frames:
                fr["verbs"] = []
        cases.append({"frame_tokens": frames, "aggregated": agg(frames)})
    with open(os.path.join(GOLDEN_DIR, "tokenization.json"), "w") as f:
        json.dump({"F": F_, "T": T, "D": D, "k": k, "img_seed": 0, "bank_seed": 1, "topk_indices": inds,
                   "aggregate_cases": cases}, f)
    print("wrote tokenization.json")


def golden_sharding():
    out = []
    for n, world in [(0, 1), (1, 1), (7, 2), (8, 2), (2990, 8), (6250, 8), (5, 8), (16, 4), (1000, 3)]:
        data = list(range(n))
        slices = []
        for rank in range(world):
            step = len(data) // world + 1      # run_visual_tokenization.py:429-431 / run_video_CapFilt.py:239-241
            start = rank * step
            end = min(len(data), start + step)
            s = data[start:end]
            slices.append([s[0], s[-1] + 1] if s else None)
        out.append({"n": n, "world": world, "slices": slices})
    with open(os.path.join(GOLDEN_DIR, "sharding.json"), "w") as f:
        json.dump(out, f)
    print("wrote sharding.json")


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    golden_vit("tiny", 32, 2, 1, [0, 1], "vit_tiny.npz")
    golden_vit("large", 224, 1, 4, [0, 11, 23], "vit_large_224.npz")
    golden_vit("base", 384, 1, 8, [0, 11], "vit_base_384.npz")
    rs.uninstall_shims()
    golden_clip("tiny", 2, 1, "clip_tiny.npz")
    golden_clip("large14", 1, 8, "clip_large14.npz")
    golden_clip_text("tiny", 5, 12, "clip_text_tiny.npz")
    golden_clip_text("large14", 3, 20, "clip_text_large14.npz")
    golden_med("tiny", 3, 9, 5, 1, "med_tiny.npz")
    golden_med("base_l", 2, 8, 197, 97, "med_base_l.npz")
    golden_preprocess()
    golden_clip_preprocess()
    golden_tokenization()
    golden_sharding()


if __name__ == "__main__":
    main()
