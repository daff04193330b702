"""CPU restatement of the CLIP text tower the reference calls through transformers — parity oracle.

TEST INFRASTRUCTURE (see oracle/vit_oracle.py for the import rule).

The reference computes the phrase bank with `CLIPModel(**inputs).text_embeds` (run_visual_tokenization.py:84-96);
the arithmetic is transformers' modeling_clip.py (un-vendored, unpinned; 5.5.0 installed here): CLIPTextEmbeddings.forward
:234-258, CLIPEncoderLayer :363-385 with a causal mask (:546-557), final_layer_norm :562, EOS pooling :564-585,
text_projection + L2 normalisation in CLIPModel.forward.  Pinned against outputs of that library: fixtures
tests/golden/clip_text_*.npz (oracle/make_golden.py).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .clip_oracle import quick_gelu


@torch.no_grad()
def clip_text_forward(sd: dict, input_ids: torch.Tensor, num_heads: int, eos_token_id: int, eps: float = 1e-5,
                      dtype: torch.dtype = torch.float32):
    """input_ids [B, L] -> (text_embeds [B, proj] unit-norm, last_hidden_state after final_layer_norm [B, L, D]).
    `sd` uses CLIPModel.state_dict() key names (text_model.* and text_projection.weight)."""
    sd = {k: v.to(dtype) for k, v in sd.items()}
    t = "text_model."
    B, L = input_ids.shape
    h = sd[t + "embeddings.token_embedding.weight"][input_ids] + sd[t + "embeddings.position_embedding.weight"][:L]
    D = h.shape[-1]
    hd = D // num_heads
    causal = torch.full((L, L), float("-inf"), dtype=dtype).triu(1)
    n_layers = 1 + max(int(k.split(".")[3]) for k in sd if k.startswith(t + "encoder.layers."))
    for i in range(n_layers):
        p = f"{t}encoder.layers.{i}."
        y = F.layer_norm(h, (D,), sd[p + "layer_norm1.weight"], sd[p + "layer_norm1.bias"], eps)
        q = F.linear(y, sd[p + "self_attn.q_proj.weight"], sd[p + "self_attn.q_proj.bias"]).view(B, L, num_heads, hd).transpose(1, 2)
        k = F.linear(y, sd[p + "self_attn.k_proj.weight"], sd[p + "self_attn.k_proj.bias"]).view(B, L, num_heads, hd).transpose(1, 2)
        v = F.linear(y, sd[p + "self_attn.v_proj.weight"], sd[p + "self_attn.v_proj.bias"]).view(B, L, num_heads, hd).transpose(1, 2)
        att = torch.matmul(q, k.transpose(-1, -2)) * hd ** -0.5 + causal
        att = F.softmax(att, dim=-1)
        o = torch.matmul(att, v).transpose(1, 2).reshape(B, L, D)
        h = h + F.linear(o, sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"])
        y = F.layer_norm(h, (D,), sd[p + "layer_norm2.weight"], sd[p + "layer_norm2.bias"], eps)
        y = F.linear(quick_gelu(F.linear(y, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])), sd[p + "mlp.fc2.weight"],
                     sd[p + "mlp.fc2.bias"])
        h = h + y
    h = F.layer_norm(h, (D,), sd[t + "final_layer_norm.weight"], sd[t + "final_layer_norm.bias"], eps)
    if eos_token_id == 2:
        pos = input_ids.to(torch.int).argmax(dim=-1)
    else:
        pos = (input_ids.to(torch.int) == eos_token_id).int().argmax(dim=-1)
    pooled = h[torch.arange(B), pos]
    emb = F.linear(pooled, sd["text_projection.weight"])
    return emb / emb.norm(p=2, dim=-1, keepdim=True), h
