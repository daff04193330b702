"""CPU restatement of the CLIP image tower the reference calls through transformers — parity oracle.

TEST INFRASTRUCTURE (see oracle/vit_oracle.py for the import rule).

The reference does not contain this arithmetic: `run_visual_tokenization.py:9,347,350` instantiates
`transformers.CLIPModel` ("openai/clip-vit-large-patch14", pipeline_config_msrvtt_test.yaml:15) and reads
`outputs.image_embeds` (`:138-142`).  `transformers` is an un-vendored, unpinned dependency
(docker/requirements.txt:9); the published algorithm restated here is transformers' modeling_clip.py as
installed in this image (5.5.0): CLIPVisionEmbeddings.forward :202-219, eager_attention_forward :261-279,
CLIPAttention.forward :300-336, CLIPMLP :339-351 with hidden_act="quick_gelu", CLIPEncoderLayer.forward
:363-385, CLIPVisionTransformer.forward :667-691, CLIPModel.forward :916-923.

Pinned against outputs of that library (a CLIPModel built from an explicit CLIPConfig with the synthetic
weights of oracle/weights.py; no download): fixtures tests/golden/clip_*.npz made by oracle/make_golden.py.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def quick_gelu(x: torch.Tensor) -> torch.Tensor:
    """transformers.activations.QuickGELUActivation: x * sigmoid(1.702 x)."""
    return x * torch.sigmoid(1.702 * x)


@torch.no_grad()
def clip_vision_forward(sd: dict, pixel_values: torch.Tensor, num_heads: int, eps: float = 1e-5,
                        dtype: torch.dtype = torch.float32):
    """pixel_values [F,3,S,S] -> (image_embeds [F,proj] unit-norm, last_hidden_state [F,P+1,D]).

    `sd` uses CLIPModel.state_dict() key names (vision_model.* and visual_projection.weight).
    """
    sd = {k: v.to(dtype) for k, v in sd.items()}
    x = pixel_values.to(dtype)
    v = "vision_model."
    B = x.shape[0]
    w = sd[v + "embeddings.patch_embedding.weight"]
    pe = F.conv2d(x, w, None, stride=w.shape[-1]).flatten(2).transpose(1, 2)           # :210-211 (no bias)
    cls = sd[v + "embeddings.class_embedding"].expand(B, 1, -1)                         # :213
    h = torch.cat([cls, pe], dim=1) + sd[v + "embeddings.position_embedding.weight"]    # :214,218
    D = h.shape[-1]
    h = F.layer_norm(h, (D,), sd[v + "pre_layrnorm.weight"], sd[v + "pre_layrnorm.bias"], eps)   # :677
    n_layers = 1 + max(int(k.split(".")[3]) for k in sd if k.startswith(v + "encoder.layers."))
    hd = D // num_heads
    for i in range(n_layers):
        p = f"{v}encoder.layers.{i}."
        res = h
        y = F.layer_norm(h, (D,), sd[p + "layer_norm1.weight"], sd[p + "layer_norm1.bias"], eps)  # :372
        N = y.shape[1]
        q = F.linear(y, sd[p + "self_attn.q_proj.weight"], sd[p + "self_attn.q_proj.bias"])       # :310-312
        k = F.linear(y, sd[p + "self_attn.k_proj.weight"], sd[p + "self_attn.k_proj.bias"])
        val = F.linear(y, sd[p + "self_attn.v_proj.weight"], sd[p + "self_attn.v_proj.bias"])
        q = q.view(B, N, num_heads, hd).transpose(1, 2)                                           # :314-316
        k = k.view(B, N, num_heads, hd).transpose(1, 2)
        val = val.view(B, N, num_heads, hd).transpose(1, 2)
        att = torch.matmul(q, k.transpose(-1, -2)) * hd ** -0.5                                   # :271, scale :291
        att = F.softmax(att, dim=-1, dtype=torch.float32 if dtype == torch.float32 else dtype).to(q.dtype)  # :274
        o = torch.matmul(att, val).transpose(1, 2).reshape(B, N, D)                               # :277-278,333
        o = F.linear(o, sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"])  # :334
        h = res + o                                                                               # :378
        res = h
        y = F.layer_norm(h, (D,), sd[p + "layer_norm2.weight"], sd[p + "layer_norm2.bias"], eps)  # :381
        y = F.linear(y, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])                         # :348
        y = quick_gelu(y)                                                                         # :349
        y = F.linear(y, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])                         # :350
        h = res + y                                                                               # :383
    last_hidden = h
    pooled = F.layer_norm(h[:, 0, :], (D,), sd[v + "post_layernorm.weight"], sd[v + "post_layernorm.bias"], eps)  # :685-686
    emb = F.linear(pooled, sd["visual_projection.weight"])                                        # :917
    emb = emb / torch.pow(torch.sum(torch.pow(emb, 2), dim=-1, keepdim=True), 0.5)                # :923, :57-65
    return emb, last_hidden
