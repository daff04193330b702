"""CPU restatement of the reference BLIP ViT forward (models/vit.py) — the parity oracle.

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
leg may import this module; the product path (vidil_b200/) never does.

Parity status: the reference ships no tests or golden vectors for this path, so the
oracle is pinned against *outputs of the reference itself*: oracle/make_golden.py imports the unmodified
/root/reference/models/vit.py (behind import shims for the absent timm/fairscale packages), runs it on
seeded inputs and commits the results under tests/golden/; tests/test_oracle_golden.py checks this
restatement against those fixtures (bit-for-bit on the same torch build: it issues the same ATen ops in the
same order).

The restatement is plain functional tensor arithmetic on a parameter dict with the reference's
state_dict keys; no nn.Module, no autograd.  `dtype=torch.float64` gives a higher-precision ground truth.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def patch_embed(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None) -> torch.Tensor:
    """timm PatchEmbed as used at models/vit.py:144-145,182: Conv2d(3, D, k=ps, s=ps) -> flatten(2).transpose(1,2).

    timm is an un-vendored, unpinned dependency of the reference (docker/requirements.txt:19; upstream BLIP
    used 0.4.12); its PatchEmbed.forward is `self.proj(x).flatten(2).transpose(1, 2)` with norm = Identity.
    """
    ps = weight.shape[-1]
    return F.conv2d(x, weight, bias, stride=ps).flatten(2).transpose(1, 2)


def attention(x: torch.Tensor, sd: dict, prefix: str, num_heads: int) -> torch.Tensor:
    """Attention.forward, models/vit.py:70-86 (dropouts are identity in eval)."""
    B, N, C = x.shape
    qkv = F.linear(x, sd[prefix + "qkv.weight"], sd[prefix + "qkv.bias"])              # :72
    qkv = qkv.reshape(B, N, 3, num_heads, C // num_heads).permute(2, 0, 3, 1, 4)        # :72
    q, k, v = qkv[0], qkv[1], qkv[2]                                                    # :73
    scale = (C // num_heads) ** -0.5                                                    # :49
    attn = (q @ k.transpose(-2, -1)) * scale                                            # :75
    attn = attn.softmax(dim=-1)                                                         # :76
    x = (attn @ v).transpose(1, 2).reshape(B, N, C)                                     # :83
    return F.linear(x, sd[prefix + "proj.weight"], sd[prefix + "proj.bias"])            # :84


def mlp(x: torch.Tensor, sd: dict, prefix: str) -> torch.Tensor:
    """Mlp.forward, models/vit.py:35-41: fc1 -> nn.GELU() (exact erf) -> fc2."""
    x = F.linear(x, sd[prefix + "fc1.weight"], sd[prefix + "fc1.bias"])
    x = F.gelu(x)
    return F.linear(x, sd[prefix + "fc2.weight"], sd[prefix + "fc2.bias"])


def block(x: torch.Tensor, sd: dict, i: int, num_heads: int, eps: float = 1e-6) -> torch.Tensor:
    """Block.forward, models/vit.py:107-110 (drop_path is identity in eval; LayerNorm eps=1e-6, :142)."""
    p = f"blocks.{i}."
    D = x.shape[-1]
    h = F.layer_norm(x, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], eps)
    x = x + attention(h, sd, p + "attn.", num_heads)                                    # :108
    h = F.layer_norm(x, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], eps)
    return x + mlp(h, sd, p + "mlp.")                                                   # :109


@torch.no_grad()
def vit_forward(sd: dict, x: torch.Tensor, num_heads: int, depth: int | None = None, eps: float = 1e-6,
                dtype: torch.dtype = torch.float32, return_blocks: bool = False):
    """VisionTransformer.forward, models/vit.py:180-194: frames [B,3,S,S] -> all tokens [B, P+1, D] after `norm`."""
    sd = {k: v.to(dtype) for k, v in sd.items()}
    x = x.to(dtype)
    if depth is None:
        depth = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("blocks."))
    B = x.shape[0]
    x = patch_embed(x, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"])      # :182
    cls = sd["cls_token"].expand(B, -1, -1)                                             # :184
    x = torch.cat((cls, x), dim=1)                                                      # :185
    x = x + sd["pos_embed"][:, :x.size(1), :]                                           # :187
    outs = []
    for i in range(depth):                                                              # :190-191
        x = block(x, sd, i, num_heads, eps)
        if return_blocks:
            outs.append(x)
    D = x.shape[-1]
    x = F.layer_norm(x, (D,), sd["norm.weight"], sd["norm.bias"], eps)                  # :192
    return (x, outs) if return_blocks else x


def flops_per_frame(embed_dim: int, depth: int, tokens: int, patch_size: int = 16, mlp_ratio: int = 4,
                    proj_dim: int = 0) -> float:
    """Algorithmic FLOPs (2*m*n*k over every GEMM) of one frame — SURVEY.md §8d / BASELINE.md §3."""
    D, N = embed_dim, tokens
    per_layer = 2 * N * D * 3 * D + 2 * N * D * D + 2 * 2 * N * D * mlp_ratio * D + 2 * 2 * N * N * D
    patch = 2 * (N - 1) * (3 * patch_size * patch_size) * D
    return float(depth * per_layer + patch + 2 * D * proj_dim)


def interpolate_pos_embed(pos_embed_checkpoint: torch.Tensor, num_patches: int, num_extra_tokens: int = 1):
    """interpolate_pos_embed, models/vit.py:281-305: bicubic resize of the grid part of a checkpoint's pos_embed."""
    D = pos_embed_checkpoint.shape[-1]
    orig = int((pos_embed_checkpoint.shape[-2] - num_extra_tokens) ** 0.5)
    new = int(num_patches ** 0.5)
    if orig == new:
        return pos_embed_checkpoint
    extra = pos_embed_checkpoint[:, :num_extra_tokens]
    pos = pos_embed_checkpoint[:, num_extra_tokens:].reshape(-1, orig, orig, D).permute(0, 3, 1, 2)
    pos = F.interpolate(pos, size=(new, new), mode="bicubic", align_corners=False)
    pos = pos.permute(0, 2, 3, 1).flatten(1, 2)
    return torch.cat((extra, pos), dim=1)


assert math.isclose(flops_per_frame(1024, 24, 197) / 1e9, 123.107, rel_tol=1e-4)
