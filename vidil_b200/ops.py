"""Torch-tensor front ends of the operator-level C-ABI entry points.

PyTorch is used for device memory and streams only: every function hands raw device pointers of
contiguous CUDA tensors to libvidil_b200.so and enqueues on torch's current stream.  CPU tensors are
rejected — there is no host implementation.
"""
from __future__ import annotations

import torch

from . import _lib


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _dev_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"vidil_b200: {name} must be a CUDA tensor (no CPU path exists)")
    return t.contiguous().float()


def _workspace(nbytes: int, device) -> torch.Tensor:
    # torch's caching allocator returns >=512-byte aligned blocks; the ABI wants 1024
    buf = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
    off = (-buf.data_ptr()) % 1024
    return buf[off:off + nbytes]


def linear(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None = None, *, epilogue: int = _lib.EPI_STORE,
           dtype: str = "bf16", cta_group: int = 2, out: torch.Tensor | None = None, pos: torch.Tensor | None = None,
           patches_per_frame: int = 0) -> torch.Tensor:
    """epilogue(a[M,K] @ w[N,K]^T + bias) through the tcgen05 GEMM; operands rounded to `dtype`, fp32 result.

    EPI_RESID accumulates into `out` (fp32 [M,N]); EPI_PATCH scatters rows into `out`
    ([frames*(P+1), N]) adding pos[1+p] — the two in-place modes of the forward.
    """
    lib = _lib.load()
    a = _dev_f32(a, "a")
    w = _dev_f32(w, "w")
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K
    if bias is not None:
        bias = _dev_f32(bias, "bias")
    if pos is not None:
        pos = _dev_f32(pos, "pos")
    if out is None:
        assert epilogue not in (_lib.EPI_RESID, _lib.EPI_PATCH), "in-place epilogues need `out`"
        out = torch.empty(M, N, dtype=torch.float32, device=a.device)
    assert out.is_cuda and out.dtype == torch.float32 and out.is_contiguous()
    ws = _workspace(lib.vidil_op_linear_workspace_bytes(M, N, K), a.device)
    st = lib.vidil_op_linear(a.data_ptr(), w.data_ptr(), bias.data_ptr() if bias is not None else None, out.data_ptr(),
                             M, N, K, epilogue, _lib.DTYPES[dtype], cta_group,
                             pos.data_ptr() if pos is not None else None, patches_per_frame, ws.data_ptr(), ws.numel(),
                             _stream())
    _lib.check(st, "vidil_op_linear")
    return out


def layernorm(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float) -> torch.Tensor:
    lib = _lib.load()
    x = _dev_f32(x, "x")
    D = x.shape[-1]
    rows = x.numel() // D
    out = torch.empty_like(x)
    st = lib.vidil_op_layernorm(x.data_ptr(), _dev_f32(weight, "weight").data_ptr(), _dev_f32(bias, "bias").data_ptr(),
                                out.data_ptr(), rows, D, float(eps), _stream())
    _lib.check(st, "vidil_op_layernorm")
    return out


def attention(qkv: torch.Tensor, num_heads: int, scale: float | None = None, dtype: str = "bf16") -> torch.Tensor:
    """qkv [B, N, 3*H*64] (the fused-QKV Linear output, models/vit.py:72) -> [B, N, H*64]."""
    lib = _lib.load()
    qkv = _dev_f32(qkv, "qkv")
    B, N, C3 = qkv.shape
    H = num_heads
    assert C3 == 3 * H * 64, "head_dim must be 64"
    if scale is None:
        scale = 64 ** -0.5
    out = torch.empty(B, N, H * 64, dtype=torch.float32, device=qkv.device)
    ws = _workspace(lib.vidil_op_attention_workspace_bytes(B, N, H), qkv.device)
    st = lib.vidil_op_attention(qkv.data_ptr(), out.data_ptr(), B, N, H, float(scale), _lib.DTYPES[dtype], ws.data_ptr(),
                                ws.numel(), _stream())
    _lib.check(st, "vidil_op_attention")
    return out


def sim_topk(image_embeds: torch.Tensor, text_embeds: torch.Tensor, k: int):
    """Top-k of image_embeds @ text_embeds.t() per row: (scores fp32 [F,k], indices int32 [F,k]), best first.

    Replaces `sims_matrix = image_embeds @ text_embeds.t()` + `.cpu().numpy()` + `np.argsort(...)[::-1][:k]`
    (run_visual_tokenization.py:276,299,306) without materialising the matrix on the host.
    """
    lib = _lib.load()
    img = _dev_f32(image_embeds, "image_embeds")
    bank = _dev_f32(text_embeds, "text_embeds")
    F, D = img.shape
    T = bank.shape[0]
    assert bank.shape[1] == D
    scores = torch.empty(F, k, dtype=torch.float32, device=img.device)
    idx = torch.empty(F, k, dtype=torch.int32, device=img.device)
    if F == 0:
        return scores, idx
    ws = _workspace(lib.vidil_sim_topk_workspace_bytes(F, T, D), img.device)
    st = lib.vidil_sim_topk(img.data_ptr(), bank.data_ptr(), F, T, D, k, scores.data_ptr(), idx.data_ptr(), ws.data_ptr(),
                            ws.numel(), _stream())
    _lib.check(st, "vidil_sim_topk")
    return scores, idx
