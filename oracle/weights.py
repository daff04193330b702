"""Deterministic synthetic parameters for the towers on the hot path.

TEST INFRASTRUCTURE — only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may
import anything under oracle/.

There are no pretrained checkpoints offline, and a 303 M-parameter state_dict cannot be committed, so both
sides of every parity test regenerate the same tensors from a seed.  numpy's PCG64 stream is
platform-independent, unlike torch's per-ISA vectorised samplers, so the golden fixtures made in the build
container stay valid on the GPU box.

Keys and shapes are exactly the reference's `VisionTransformer.state_dict()` (models/vit.py:144-161, probed:
294 tensors for ViT-L/16), so a dict from here also exercises `load_state_dict` on the drop-in module.
"""
from __future__ import annotations

import numpy as np
import torch

VIT_CONFIGS = {
    # name: (embed_dim, depth, num_heads) — models/blip.py:309-322
    "tiny": (128, 2, 2),  # not a reference config; small enough for pure-CPU tests
    "base": (768, 12, 12),
    "large": (1024, 24, 16),
}


def _normal(rng: np.random.Generator, shape, std: float, mean: float = 0.0) -> torch.Tensor:
    a = rng.standard_normal(size=shape, dtype=np.float32)
    a *= np.float32(std)
    if mean != 0.0:
        a += np.float32(mean)
    return torch.from_numpy(a)


def vit_state_dict(vit: str = "large", image_size: int = 224, patch_size: int = 16, seed: int = 0,
                   exercise_affine: bool = True) -> dict:
    """Synthetic BLIP-ViT parameters.

    Weights ~ N(0, 0.02) like the reference's trunc_normal_(std=.02) init (models/vit.py:163-174; the ±2
    truncation at 100 sigma is immaterial).  With `exercise_affine` biases ~ N(0, 0.02) and LayerNorm
    gamma ~ N(1, 0.02), beta ~ N(0, 0.02) instead of the init's exact 0 / 1, so the bias and affine paths of
    the kernels are actually tested.
    """
    D, depth, _ = VIT_CONFIGS[vit]
    P = (image_size // patch_size) ** 2
    rng = np.random.Generator(np.random.PCG64(seed))
    bstd = 0.02 if exercise_affine else 0.0

    def bias(n):
        return _normal(rng, (n,), bstd) if exercise_affine else torch.zeros(n)

    def gamma(n):
        return _normal(rng, (n,), 0.02, 1.0) if exercise_affine else torch.ones(n)

    sd = {}
    sd["cls_token"] = _normal(rng, (1, 1, D), 0.02)
    sd["pos_embed"] = _normal(rng, (1, P + 1, D), 0.02)
    sd["patch_embed.proj.weight"] = _normal(rng, (D, 3, patch_size, patch_size), 0.02)
    sd["patch_embed.proj.bias"] = bias(D)
    for i in range(depth):
        p = f"blocks.{i}."
        sd[p + "norm1.weight"] = gamma(D)
        sd[p + "norm1.bias"] = bias(D)
        sd[p + "attn.qkv.weight"] = _normal(rng, (3 * D, D), 0.02)
        sd[p + "attn.qkv.bias"] = bias(3 * D)
        sd[p + "attn.proj.weight"] = _normal(rng, (D, D), 0.02)
        sd[p + "attn.proj.bias"] = bias(D)
        sd[p + "norm2.weight"] = gamma(D)
        sd[p + "norm2.bias"] = bias(D)
        sd[p + "mlp.fc1.weight"] = _normal(rng, (4 * D, D), 0.02)
        sd[p + "mlp.fc1.bias"] = bias(4 * D)
        sd[p + "mlp.fc2.weight"] = _normal(rng, (D, 4 * D), 0.02)
        sd[p + "mlp.fc2.bias"] = bias(D)
    sd["norm.weight"] = gamma(D)
    sd["norm.bias"] = bias(D)
    return sd


# openai/clip-vit-large-patch14 vision tower (configs/pipeline_config/pipeline_config_msrvtt_test.yaml:15);
# "tiny" is a CPU-sized stand-in with the same structure.
CLIP_CONFIGS = {
    "tiny": dict(hidden_size=128, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2, image_size=28,
                 patch_size=14, projection_dim=64),
    "large14": dict(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24, num_attention_heads=16,
                    image_size=224, patch_size=14, projection_dim=768),
}


def clip_vision_state_dict(name: str = "large14", seed: int = 0) -> dict:
    """Synthetic parameters under transformers' CLIPModel key names (vision tower + visual_projection)."""
    c = CLIP_CONFIGS[name]
    D, I, L = c["hidden_size"], c["intermediate_size"], c["num_hidden_layers"]
    P = (c["image_size"] // c["patch_size"]) ** 2
    rng = np.random.Generator(np.random.PCG64(seed))
    v = "vision_model."
    sd = {}
    sd[v + "embeddings.class_embedding"] = _normal(rng, (D,), 0.02)
    sd[v + "embeddings.patch_embedding.weight"] = _normal(rng, (D, 3, c["patch_size"], c["patch_size"]), 0.02)
    sd[v + "embeddings.position_embedding.weight"] = _normal(rng, (P + 1, D), 0.02)
    sd[v + "pre_layrnorm.weight"] = _normal(rng, (D,), 0.02, 1.0)
    sd[v + "pre_layrnorm.bias"] = _normal(rng, (D,), 0.02)
    for i in range(L):
        p = f"{v}encoder.layers.{i}."
        sd[p + "layer_norm1.weight"] = _normal(rng, (D,), 0.02, 1.0)
        sd[p + "layer_norm1.bias"] = _normal(rng, (D,), 0.02)
        for nm in ("q_proj", "k_proj", "v_proj", "out_proj"):
            sd[p + f"self_attn.{nm}.weight"] = _normal(rng, (D, D), 0.02)
            sd[p + f"self_attn.{nm}.bias"] = _normal(rng, (D,), 0.02)
        sd[p + "layer_norm2.weight"] = _normal(rng, (D,), 0.02, 1.0)
        sd[p + "layer_norm2.bias"] = _normal(rng, (D,), 0.02)
        sd[p + "mlp.fc1.weight"] = _normal(rng, (I, D), 0.02)
        sd[p + "mlp.fc1.bias"] = _normal(rng, (I,), 0.02)
        sd[p + "mlp.fc2.weight"] = _normal(rng, (D, I), 0.02)
        sd[p + "mlp.fc2.bias"] = _normal(rng, (D,), 0.02)
    sd[v + "post_layernorm.weight"] = _normal(rng, (D,), 0.02, 1.0)
    sd[v + "post_layernorm.bias"] = _normal(rng, (D,), 0.02)
    sd["visual_projection.weight"] = _normal(rng, (c["projection_dim"], D), 0.02)
    return sd


# openai/clip-vit-large-patch14 text tower; "tiny" is a CPU-sized stand-in with the same structure.
CLIP_TEXT_CONFIGS = {
    "tiny": dict(vocab_size=96, max_position_embeddings=16, hidden_size=128, intermediate_size=512, num_hidden_layers=2,
                 num_attention_heads=2, projection_dim=64, eos_token_id=95),
    "large14": dict(vocab_size=49408, max_position_embeddings=77, hidden_size=768, intermediate_size=3072,
                    num_hidden_layers=12, num_attention_heads=12, projection_dim=768, eos_token_id=49407),
}


def clip_text_state_dict(name: str = "large14", seed: int = 0) -> dict:
    """Synthetic parameters under transformers' CLIPModel key names (text tower + text_projection)."""
    c = CLIP_TEXT_CONFIGS[name]
    D, I, L = c["hidden_size"], c["intermediate_size"], c["num_hidden_layers"]
    rng = np.random.Generator(np.random.PCG64(3_000_017 + seed))
    t = "text_model."
    sd = {}
    sd[t + "embeddings.token_embedding.weight"] = _normal(rng, (c["vocab_size"], D), 0.02)
    sd[t + "embeddings.position_embedding.weight"] = _normal(rng, (c["max_position_embeddings"], D), 0.01)
    for i in range(L):
        p = f"{t}encoder.layers.{i}."
        sd[p + "layer_norm1.weight"] = _normal(rng, (D,), 0.02, 1.0)
        sd[p + "layer_norm1.bias"] = _normal(rng, (D,), 0.02)
        for nm in ("q_proj", "k_proj", "v_proj", "out_proj"):
            sd[p + f"self_attn.{nm}.weight"] = _normal(rng, (D, D), 0.03)
            sd[p + f"self_attn.{nm}.bias"] = _normal(rng, (D,), 0.02)
        sd[p + "layer_norm2.weight"] = _normal(rng, (D,), 0.02, 1.0)
        sd[p + "layer_norm2.bias"] = _normal(rng, (D,), 0.02)
        sd[p + "mlp.fc1.weight"] = _normal(rng, (I, D), 0.03)
        sd[p + "mlp.fc1.bias"] = _normal(rng, (I,), 0.02)
        sd[p + "mlp.fc2.weight"] = _normal(rng, (D, I), 0.03)
        sd[p + "mlp.fc2.bias"] = _normal(rng, (D,), 0.02)
    sd[t + "final_layer_norm.weight"] = _normal(rng, (D,), 0.02, 1.0)
    sd[t + "final_layer_norm.bias"] = _normal(rng, (D,), 0.02)
    sd["text_projection.weight"] = _normal(rng, (c["projection_dim"], D), 0.03)
    return sd


def token_ids(name: str, batch: int, seq_len: int, seed: int = 0) -> torch.Tensor:
    """Synthetic tokenised phrases: BOS, 1..seq_len-2 word ids, EOS, then padding (= EOS id, as CLIP's tokenizer pads)."""
    c = CLIP_TEXT_CONFIGS[name]
    rng = np.random.Generator(np.random.PCG64(4_000_037 + seed))
    eos = c["eos_token_id"]
    ids = np.full((batch, seq_len), eos, dtype=np.int64)
    for b in range(batch):
        n_words = int(rng.integers(1, seq_len - 1))
        ids[b, 0] = eos - 1                                            # BOS is eos - 1 in CLIP's vocabulary
        ids[b, 1:1 + n_words] = rng.integers(1, eos - 1, size=n_words)
    return torch.from_numpy(ids)


def frames(batch: int, image_size: int = 224, seed: int = 0) -> torch.Tensor:
    """Synthetic post-Normalize frames ~ N(0,1), fp32 NCHW."""
    rng = np.random.Generator(np.random.PCG64(1_000_003 + seed))
    return torch.from_numpy(rng.standard_normal(size=(batch, 3, image_size, image_size), dtype=np.float32))


def unit_rows(n: int, d: int, seed: int) -> torch.Tensor:
    """n unit-norm fp32 rows — stand-ins for CLIP image/text embeddings."""
    rng = np.random.Generator(np.random.PCG64(2_000_003 + seed))
    a = rng.standard_normal(size=(n, d), dtype=np.float32)
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    return torch.from_numpy(a)


def u8_frames(batch: int, height: int, width: int, seed: int = 0) -> torch.Tensor:
    """Synthetic decoded video frames: uint8 [B, H, W, 3] with smooth structure, noise and saturated patches (so that the
    bicubic overshoot clips at both ends)."""
    rng = np.random.Generator(np.random.PCG64(5_000_011 + seed))
    yy, xx = np.mgrid[0:height, 0:width].astype(np.float32)
    out = np.empty((batch, height, width, 3), dtype=np.uint8)
    for b in range(batch):
        f = rng.uniform(0.01, 0.2, size=(3, 2)).astype(np.float32)
        img = np.stack([127.5 + 100 * np.sin(f[c, 0] * xx + b) * np.cos(f[c, 1] * yy) for c in range(3)], axis=-1)
        img += rng.normal(0, 25, size=img.shape).astype(np.float32)
        img[: height // 8, : width // 4] = 255
        img[-(height // 8):, -(width // 4):] = 0
        out[b] = np.clip(img, 0, 255).astype(np.uint8)
    return torch.from_numpy(out)


# BLIP's mixture-of-encoder-decoder text stack (configs/med_config.json + blip.py:95-97: encoder_width = vision width).
# "tiny" is a CPU-sized stand-in with the same structure; "base_l" pairs BERT-base with ViT-L tokens (width 1024),
# "base_b" with ViT-B tokens (width 768, the shipped pipeline config).
MED_CONFIGS = {
    "tiny": dict(vocab_size=200, hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=512,
                 max_position_embeddings=64, encoder_width=128, layer_norm_eps=1e-12),
    "base_l": dict(vocab_size=30524, hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072,
                   max_position_embeddings=512, encoder_width=1024, layer_norm_eps=1e-12),
    "base_b": dict(vocab_size=30524, hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072,
                   max_position_embeddings=512, encoder_width=768, layer_norm_eps=1e-12),
}
# bert-base-uncased ids the reference relies on (blip.py:283-291): [PAD] 0, [CLS] 101, [SEP] 102 (eos), [DEC] 30522 (bos),
# [ENC] 30523; 'a picture of' = 1037 3861 1997.  The tiny vocabulary keeps the same roles at small ids.
MED_SPECIAL = {
    "tiny": dict(pad=0, cls=101, eos=102, bos=198, enc=199, prompt=[198, 37, 61, 97]),
    "base_l": dict(pad=0, cls=101, eos=102, bos=30522, enc=30523, prompt=[30522, 1037, 3861, 1997]),
    "base_b": dict(pad=0, cls=101, eos=102, bos=30522, enc=30523, prompt=[30522, 1037, 3861, 1997]),
}


def med_state_dict(name: str = "base_l", kind: str = "decoder", seed: int = 0, logit_std: float = 0.1) -> dict:
    """Synthetic parameters under the reference's key names: kind 'decoder' -> BLIP_Decoder.text_decoder
    (BertLMHeadModel: 'text_decoder.bert.*', 'text_decoder.cls.predictions.*'); kind 'itm' -> BLIP_ITM
    ('text_encoder.*', 'itm_head.*').  The LM decoder matrix gets std `logit_std` so that the logits of a random-init
    model are spread out like a trained one's (std ~ sqrt(D) * logit_std) instead of being a near-tie everywhere."""
    c = MED_CONFIGS[name]
    D, I, L, E, V = c["hidden_size"], c["intermediate_size"], c["num_hidden_layers"], c["encoder_width"], c["vocab_size"]
    rng = np.random.Generator(np.random.PCG64((6_000_011 if kind == "decoder" else 7_000_003) + seed))
    pre = "text_decoder.bert." if kind == "decoder" else "text_encoder."
    sd = {}
    sd[pre + "embeddings.word_embeddings.weight"] = _normal(rng, (V, D), 0.05)
    sd[pre + "embeddings.position_embeddings.weight"] = _normal(rng, (c["max_position_embeddings"], D), 0.05)
    sd[pre + "embeddings.LayerNorm.weight"] = _normal(rng, (D,), 0.02, 1.0)
    sd[pre + "embeddings.LayerNorm.bias"] = _normal(rng, (D,), 0.02)
    for i in range(L):
        p = f"{pre}encoder.layer.{i}."
        for att, kv_in in (("attention", D), ("crossattention", E)):
            sd[p + att + ".self.query.weight"] = _normal(rng, (D, D), 0.04)
            sd[p + att + ".self.query.bias"] = _normal(rng, (D,), 0.02)
            sd[p + att + ".self.key.weight"] = _normal(rng, (D, kv_in), 0.04)
            sd[p + att + ".self.key.bias"] = _normal(rng, (D,), 0.02)
            sd[p + att + ".self.value.weight"] = _normal(rng, (D, kv_in), 0.04)
            sd[p + att + ".self.value.bias"] = _normal(rng, (D,), 0.02)
            sd[p + att + ".output.dense.weight"] = _normal(rng, (D, D), 0.04)
            sd[p + att + ".output.dense.bias"] = _normal(rng, (D,), 0.02)
            sd[p + att + ".output.LayerNorm.weight"] = _normal(rng, (D,), 0.02, 1.0)
            sd[p + att + ".output.LayerNorm.bias"] = _normal(rng, (D,), 0.02)
        sd[p + "intermediate.dense.weight"] = _normal(rng, (I, D), 0.04)
        sd[p + "intermediate.dense.bias"] = _normal(rng, (I,), 0.02)
        sd[p + "output.dense.weight"] = _normal(rng, (D, I), 0.04)
        sd[p + "output.dense.bias"] = _normal(rng, (D,), 0.02)
        sd[p + "output.LayerNorm.weight"] = _normal(rng, (D,), 0.02, 1.0)
        sd[p + "output.LayerNorm.bias"] = _normal(rng, (D,), 0.02)
    if kind == "decoder":
        h = "text_decoder.cls.predictions."
        sd[h + "transform.dense.weight"] = _normal(rng, (D, D), 0.04)
        sd[h + "transform.dense.bias"] = _normal(rng, (D,), 0.02)
        sd[h + "transform.LayerNorm.weight"] = _normal(rng, (D,), 0.02, 1.0)
        sd[h + "transform.LayerNorm.bias"] = _normal(rng, (D,), 0.02)
        sd[h + "decoder.weight"] = _normal(rng, (V, D), logit_std)
        sd[h + "bias"] = _normal(rng, (V,), 0.5)
    else:
        sd["itm_head.weight"] = _normal(rng, (2, D), 0.1)
        sd["itm_head.bias"] = _normal(rng, (2,), 0.1)
    return sd


def image_tokens(batch: int, n_tokens: int, width: int, seed: int = 0) -> torch.Tensor:
    """Stand-in ViT output (post-LayerNorm tokens, ~unit scale) for tests that exercise the text stack alone."""
    rng = np.random.Generator(np.random.PCG64(8_000_009 + seed))
    return torch.from_numpy(rng.standard_normal(size=(batch, n_tokens, width), dtype=np.float32))


def caption_ids(name: str, batch: int, seq_len: int, seed: int = 0, min_words: int | None = None):
    """Synthetic tokenised captions as the ITM tokenizer call lays them out (blip_itm.py:46-47, padding='max_length'):
    [CLS]-slot id, words, [SEP], then [PAD]; returns (ids int64 [B,T], attention_mask int64 [B,T])."""
    c, s = MED_CONFIGS[name], MED_SPECIAL[name]
    rng = np.random.Generator(np.random.PCG64(9_000_011 + seed))
    ids = np.full((batch, seq_len), s["pad"], dtype=np.int64)
    mask = np.zeros((batch, seq_len), dtype=np.int64)
    lo = 1 if min_words is None else min_words
    for b in range(batch):
        n_words = int(rng.integers(lo, seq_len - 1))
        ids[b, 0] = s["cls"]
        ids[b, 1:1 + n_words] = rng.integers(103, min(c["vocab_size"], 30522) - 4, size=n_words)
        ids[b, 1 + n_words] = s["eos"]
        mask[b, :n_words + 2] = 1
    return torch.from_numpy(ids), torch.from_numpy(mask)
