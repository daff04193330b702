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


def attention(qkv: torch.Tensor, num_heads: int, scale: float | None = None, dtype: str = "bf16", causal: bool = False) -> torch.Tensor:
    """qkv [B, N, 3*H*64] (the fused-QKV Linear output, models/vit.py:72) -> [B, N, H*64].  causal: query i sees keys 0..i
    (the CLIP text tower)."""
    lib = _lib.load()
    qkv = _dev_f32(qkv, "qkv")
    B, N, C3 = qkv.shape
    H = num_heads
    assert C3 == 3 * H * 64, "head_dim must be 64"
    if scale is None:
        scale = 64 ** -0.5
    out = torch.empty(B, N, H * 64, dtype=torch.float32, device=qkv.device)
    ws = _workspace(lib.vidil_op_attention_workspace_bytes(B, N, H), qkv.device)
    fn = lib.vidil_op_attention_causal if causal else lib.vidil_op_attention
    st = fn(qkv.data_ptr(), out.data_ptr(), B, N, H, float(scale), _lib.DTYPES[dtype], ws.data_ptr(), ws.numel(), _stream())
    _lib.check(st, "vidil_op_attention")
    return out


class SimBank:
    """A phrase bank prepared once for many `topk` calls (vidil_sim_bank_create): fp16 copy for the tensor cores, fp32 copy
    for the exact re-scoring, largest row norm for the error bound — all owned by the native handle."""

    def __init__(self, text_embeds: torch.Tensor):
        import ctypes
        self.lib = _lib.load()
        bank = _dev_f32(text_embeds, "text_embeds")
        self.T, self.D = bank.shape
        self.device = bank.device
        self.handle = ctypes.c_void_p()
        self._ws = None   # workspace, kept between calls: same address -> the library reuses its encoded TMA maps
        with torch.cuda.device(self.device):
            st = self.lib.vidil_sim_bank_create(bank.data_ptr(), self.T, self.D, _stream(), ctypes.byref(self.handle))
        _lib.check(st, "vidil_sim_bank_create")

    def __del__(self):
        try:
            if self.handle:
                self.lib.vidil_sim_bank_destroy(self.handle)
                self.handle = None
        except Exception:  # noqa: BLE001 - interpreter teardown
            pass

    def topk(self, image_embeds: torch.Tensor, k: int):
        img = _dev_f32(image_embeds, "image_embeds")
        F, D = img.shape
        assert D == self.D
        scores = torch.empty(F, k, dtype=torch.float32, device=img.device)
        idx = torch.empty(F, k, dtype=torch.int32, device=img.device)
        if F == 0:
            return scores, idx
        with torch.cuda.device(img.device):
            need = self.lib.vidil_sim_bank_topk_workspace_bytes(self.handle, F)
            if self._ws is None or self._ws.numel() < need:
                self._ws = _workspace(need, img.device)
            ws = self._ws
            st = self.lib.vidil_sim_bank_topk(self.handle, img.data_ptr(), F, k, scores.data_ptr(), idx.data_ptr(), ws.data_ptr(),
                                              ws.numel(), _stream())
        _lib.check(st, "vidil_sim_bank_topk")
        return scores, idx


_bank_cache: dict = {}


def _cached_bank(text_embeds: torch.Tensor) -> SimBank:
    """The phrase banks of a run are the same tensors call after call (visual_tokenization.tokens_from_embeddings): their
    prepared form is cached on tensor identity + version.  The entry pins the tensor, so its address cannot be recycled."""
    key = (text_embeds.data_ptr(), tuple(text_embeds.shape), tuple(text_embeds.stride()), text_embeds.dtype, text_embeds._version)
    hit = _bank_cache.get(key)
    if hit is None:
        if len(_bank_cache) >= 16:
            _bank_cache.pop(next(iter(_bank_cache)))
        hit = (SimBank(text_embeds), text_embeds)
        _bank_cache[key] = hit
    return hit[0]


def sim_topk(image_embeds: torch.Tensor, text_embeds: torch.Tensor, k: int, cache_bank: bool = True):
    """Top-k of image_embeds @ text_embeds.t() per row: (scores fp32 [F,k], indices int32 [F,k]), best first.

    Replaces `sims_matrix = image_embeds @ text_embeds.t()` + `.cpu().numpy()` + `np.argsort(...)[::-1][:k]`
    (run_visual_tokenization.py:276,299,306): the [F,T] matrix is never written, on the device or the host.  With
    cache_bank the phrase bank's fp16 / fp32 device copies are prepared once per bank tensor (SimBank)."""
    if cache_bank:
        return _cached_bank(text_embeds).topk(image_embeds, k)
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
