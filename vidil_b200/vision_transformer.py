"""Drop-in for the reference's `models/vit.py` on the frame-encoding path.

`VisionTransformer` keeps the reference constructor signature (models/vit.py:118-121), the exact
`state_dict()` schema (cls_token, pos_embed, patch_embed.proj.*, blocks.N.{norm1,attn.qkv,attn.proj,norm2,
mlp.fc1,mlp.fc2}.*, norm.* — 294 tensors for ViT-L/16), the attributes other reference code reads
(`patch_embed.num_patches`, `pos_embed`, `embed_dim`) and the call `visual_encoder(image) -> [B, N+1, D]`
fp32 (models/blip.py:128, models/blip_itm.py:43).  The arithmetic is not here: `forward` packs the weights
once into a native encoder handle and runs the sm_100a kernels of libvidil_b200.so
(include/vidil_b200.h: vidil_vit_forward).  Parameters stay ordinary fp32 `nn.Parameter`s so
`load_checkpoint(...)` / `load_state_dict` (models/blip.py:332-354) work unchanged; they are re-packed
automatically when they change.

Inference only (the hot path runs under `@torch.no_grad()`: run_video_CapFilt.py:139,
run_visual_tokenization.py:161): no autograd graph is built, dropout/drop-path rates are accepted and
ignored exactly as `.eval()` makes them inert in the reference, and there is no CPU implementation — a CPU
module or CPU input raises.
"""
from __future__ import annotations

import ctypes
from functools import partial

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib


class PatchEmbed(nn.Module):
    """Parameter holder with timm's PatchEmbed attribute surface (used at models/vit.py:144-147, 284)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768):
        super().__init__()
        self.img_size = (img_size, img_size)
        self.patch_size = (patch_size, patch_size)
        self.grid_size = (img_size // patch_size, img_size // patch_size)
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)


class Mlp(nn.Module):
    """Parameter holder mirroring models/vit.py:23-33."""

    def __init__(self, in_features, hidden_features):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.fc2 = nn.Linear(hidden_features, in_features)


class Attention(nn.Module):
    """Parameter holder mirroring models/vit.py:44-55."""

    def __init__(self, dim, num_heads):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)


class Block(nn.Module):
    """Parameter holder mirroring models/vit.py:89-105."""

    def __init__(self, dim, num_heads, mlp_ratio, norm_layer):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads)
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))


class NativeEncoder:
    """Owns one `vidil_encoder*` plus the torch-allocated workspace it runs in."""

    def __init__(self, cfg: _lib.EncoderCfg):
        self.lib = _lib.load()
        self.cfg = cfg
        self.handle = ctypes.c_void_p()
        _lib.check(self.lib.vidil_encoder_create(ctypes.byref(cfg), ctypes.byref(self.handle)), "vidil_encoder_create")
        self.tokens = self.lib.vidil_encoder_tokens(self.handle)
        self._ws = None
        self._scratch = None
        self._pipe = None

    def __del__(self):
        try:
            if self.handle:
                self.lib.vidil_encoder_destroy(self.handle)
                self.handle = None
        except Exception:  # noqa: BLE001 - interpreter teardown
            pass

    def load(self, name: str, tensor: torch.Tensor) -> None:
        t = tensor.detach().to(dtype=torch.float32).contiguous()
        if not t.is_cuda:
            raise RuntimeError("vidil_b200: parameters must live on a CUDA device (no CPU path exists)")
        st = self.lib.vidil_encoder_load(self.handle, name.encode(), t.data_ptr(), t.numel(),
                                         torch.cuda.current_stream().cuda_stream)
        _lib.check(st, f"vidil_encoder_load({name})")

    def pipeline_scratch(self, batch: int, device) -> torch.Tensor:
        """Device scratch of the two-slot host pipeline.  It only ever grows; the library fixes the slot offsets from the
        batch it was first bound with, so later, smaller (ragged last) batches reuse it as is.  Growing it means the old
        buffer is about to be freed under copy streams torch's allocator does not know: drain both slots first."""
        need = self.lib.vidil_encoder_host_pipeline_scratch_bytes(self.handle, batch)
        if self._pipe is None or self._pipe.numel() < need or self._pipe.device != device:
            if self._pipe is not None:
                self.host_wait(0)
                self.host_wait(1)
                torch.cuda.current_stream().synchronize()
            self._pipe = self._aligned(need, device)
        return self._pipe

    def host_submit(self, frames: torch.Tensor, out: torch.Tensor, slot: int, device) -> None:
        """Enqueue H2D -> forward -> D2H of one host batch into `slot` (0/1); returns immediately."""
        B = frames.shape[0]
        scratch = self.pipeline_scratch(B, device)
        fn = self.lib.vidil_encoder_host_submit if out.dtype == torch.float32 else self.lib.vidil_encoder_host_submit16
        st = fn(self.handle, frames.data_ptr(), B, out.data_ptr(), slot, scratch.data_ptr(), scratch.numel(),
                torch.cuda.current_stream().cuda_stream)
        _lib.check(st, "vidil_encoder_host_submit")

    def host_wait(self, slot: int) -> None:
        _lib.check(self.lib.vidil_encoder_host_wait(self.handle, slot), "vidil_encoder_host_wait")

    def set_profiling(self, enable: bool) -> None:
        _lib.check(self.lib.vidil_encoder_set_profiling(self.handle, int(enable)), "vidil_encoder_set_profiling")

    def read_profile(self) -> dict:
        """{class: {ms, flops, bytes, launches}} of the kernels enqueued since the last read (synchronises)."""
        st = _lib.KernelStats()
        _lib.check(self.lib.vidil_encoder_read_profile(self.handle, ctypes.byref(st)), "vidil_encoder_read_profile")
        return {name: dict(ms=st.ms[i], flops=st.flops[i], bytes=st.bytes[i], launches=int(st.launches[i]))
                for i, name in enumerate(_lib.KCLASSES)}

    @staticmethod
    def _aligned(nbytes: int, device) -> torch.Tensor:
        buf = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
        off = (-buf.data_ptr()) % 1024
        return buf[off:off + nbytes]

    def workspace(self, batch: int, device) -> torch.Tensor:
        need = self.lib.vidil_encoder_workspace_bytes(self.handle, batch)
        if self._ws is None or self._ws.numel() < need or self._ws.device != device:
            self._ws = self._aligned(need, device)  # grows monotonically; smaller batches reuse the same base
        return self._ws

    def host_scratch(self, batch: int, device) -> torch.Tensor:
        need = self.lib.vidil_encoder_host_scratch_bytes(self.handle, batch)
        if self._scratch is None or self._scratch.numel() < need or self._scratch.device != device:
            self._scratch = self._aligned(need, device)
        return self._scratch


def _check_frames(x: torch.Tensor, img_size: int) -> torch.Tensor:
    if x.dim() != 4 or x.shape[1] != 3 or x.shape[2] != img_size or x.shape[3] != img_size:
        raise RuntimeError(f"expected frames of shape [B, 3, {img_size}, {img_size}], got {tuple(x.shape)}")
    return x


class VisionTransformer(nn.Module):
    """B200-native Vision Transformer with the reference's interface (models/vit.py:113-198)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12,
                 num_heads=12, mlp_ratio=4., qkv_bias=True, qk_scale=None, representation_size=None,
                 drop_rate=0., attn_drop_rate=0., drop_path_rate=0., norm_layer=None,
                 use_grad_checkpointing=False, ckpt_layer=0, compute_dtype="bf16", cta_group=0,
                 cache_identical_inputs=False):
        super().__init__()
        if in_chans != 3 or not qkv_bias or qk_scale is not None or mlp_ratio != 4.:
            raise ValueError("vidil_b200.VisionTransformer supports the configurations models/blip.py:create_vit "
                             "builds: in_chans=3, qkv_bias=True, qk_scale=None, mlp_ratio=4")
        if embed_dim != num_heads * 64:
            raise ValueError("head_dim must be 64 (embed_dim == 64 * num_heads), as in ViT-B/16 and ViT-L/16")
        if norm_layer is not None and getattr(norm_layer, "func", norm_layer) is not nn.LayerNorm:
            raise ValueError("only nn.LayerNorm is supported as norm_layer")
        self.num_features = self.embed_dim = embed_dim
        self.depth, self.num_heads = depth, num_heads
        self.img_size, self.patch_size = img_size, patch_size
        self.compute_dtype = compute_dtype
        self.cta_group = cta_group
        norm_layer = norm_layer or partial(nn.LayerNorm, eps=1e-6)

        self.patch_embed = PatchEmbed(img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim)
        num_patches = self.patch_embed.num_patches
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, num_patches + 1, embed_dim))
        self.blocks = nn.ModuleList([Block(embed_dim, num_heads, mlp_ratio, norm_layer) for _ in range(depth)])
        self.norm = norm_layer(embed_dim)
        self.ln_eps = float(self.norm.eps)

        # same initialisation as models/vit.py:163-174
        nn.init.trunc_normal_(self.pos_embed, std=.02)
        nn.init.trunc_normal_(self.cls_token, std=.02)
        self.apply(self._init_weights)

        self._native = None
        self._packed_sig = None
        # run_video_CapFilt.py:110-112 calls the filterer once per caption with the SAME frame tensor, re-running this
        # tower every time.  With cache_identical_inputs the output of the previous call is returned
        # when the input is the same tensor object, unmodified, and the weights have not changed.  Opt-in: the returned
        # tensor is then shared between calls, so callers must not write into it.
        self.cache_identical_inputs = cache_identical_inputs
        self._cache_key = None
        self._cache_out = None
        self._cache_in = None

    @staticmethod
    def _init_weights(m):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    @torch.jit.ignore
    def no_weight_decay(self):
        return {'pos_embed', 'cls_token'}

    # -- native handle ---------------------------------------------------------------------------
    def _encoder_cfg(self) -> _lib.EncoderCfg:
        return _lib.EncoderCfg(img_size=self.img_size, patch_size=self.patch_size, embed_dim=self.embed_dim,
                               depth=self.depth, num_heads=self.num_heads, mlp_dim=4 * self.embed_dim,
                               ln_eps=self.ln_eps, act=_lib.ACT_GELU_ERF, patch_bias=1, pre_ln=0, proj_dim=0,
                               dtype=_lib.DTYPES[self.compute_dtype], cta_group=self.cta_group)

    def _signature(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def _ensure_packed(self) -> NativeEncoder:
        sig = self._signature()
        if self._native is None:
            self._native = NativeEncoder(self._encoder_cfg())
            self._packed_sig = None
        if sig != self._packed_sig:
            for name, p in self.state_dict().items():
                self._native.load(name, p)
            _lib.check(self._native.lib.vidil_encoder_check_loaded(self._native.handle), "vidil_encoder_check_loaded")
            self._packed_sig = sig
        return self._native

    # -- forward ---------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x, register_blk=-1):
        """frames [B,3,S,S] (CUDA, float) -> tokens [B, N+1, D] fp32 after the final LayerNorm (vit.py:180-194)."""
        if register_blk != -1:
            raise NotImplementedError("attention-map hooks (register_blk) are a training/visualisation feature "
                                      "outside the inference hot path")
        if not x.is_cuda:
            raise RuntimeError("vidil_b200: frames must be on a CUDA device; use encode_host() for host buffers")
        _check_frames(x, self.img_size)
        with torch.cuda.device(x.device):
            enc = self._ensure_packed()
            key = None
            if self.cache_identical_inputs:
                key = (x.data_ptr(), tuple(x.shape), tuple(x.stride()), x.dtype, x._version, self._packed_sig)
                if key == self._cache_key and self._cache_out is not None:
                    return self._cache_out
            x_orig = x  # the cache key is THIS tensor's address and version: it must stay alive while the entry does
            x = x.contiguous().float()
            B = x.shape[0]
            out = torch.empty(B, enc.tokens, self.embed_dim, dtype=torch.float32, device=x.device)
            if B == 0:
                return out
            ws = enc.workspace(B, x.device)
            st = enc.lib.vidil_vit_forward(enc.handle, x.data_ptr(), B, out.data_ptr(), ws.data_ptr(), ws.numel(),
                                           torch.cuda.current_stream().cuda_stream)
            _lib.check(st, "vidil_vit_forward")
            if key is not None:
                # Pin the ORIGINAL input (not the converted copy): while it is referenced here the caching allocator cannot
                # hand its address to another tensor, so (data_ptr, _version) identifies these very frames.
                self._cache_key, self._cache_out, self._cache_in = key, out, x_orig
        return out

    _TORCH_DTYPES = {"bf16": torch.bfloat16, "bfloat16": torch.bfloat16, "fp16": torch.float16, "float16": torch.float16,
                     "half": torch.float16}

    @property
    def token_dtype16(self) -> torch.dtype:
        """torch dtype of the 16-bit token output (the handle's tensor-core operand type)."""
        return self._TORCH_DTYPES[self.compute_dtype]

    @torch.no_grad()
    def forward_tokens16(self, x: torch.Tensor) -> torch.Tensor:
        """frames [B,3,S,S] (CUDA) -> tokens [B, N+1, D] in the 16-bit operand type (vidil_vit_forward16): the final LayerNorm
        writes 16 bits instead of fp32.  For consumers that take 16-bit tokens anyway and for host transfers."""
        if not x.is_cuda:
            raise RuntimeError("vidil_b200: frames must be on a CUDA device")
        _check_frames(x, self.img_size)
        with torch.cuda.device(x.device):
            enc = self._ensure_packed()
            x = x.contiguous().float()
            B = x.shape[0]
            out = torch.empty(B, enc.tokens, self.embed_dim, dtype=self.token_dtype16, device=x.device)
            if B == 0:
                return out
            ws = enc.workspace(B, x.device)
            st = enc.lib.vidil_vit_forward16(enc.handle, x.data_ptr(), B, out.data_ptr(), ws.data_ptr(), ws.numel(),
                                             torch.cuda.current_stream().cuda_stream)
            _lib.check(st, "vidil_vit_forward16")
        return out

    @torch.no_grad()
    def encode_host(self, frames: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """Host-buffer call: frames is a CPU fp32 tensor (pinned for full PCIe rate); the result comes back as a
        CPU tensor.  Copies run inside the call (vidil_vit_forward_host)."""
        if frames.is_cuda:
            raise RuntimeError("encode_host takes host tensors; call the module directly for CUDA tensors")
        _check_frames(frames, self.img_size)
        dev = self.cls_token.device
        if dev.type != "cuda":
            raise RuntimeError("vidil_b200: the module must be moved to a CUDA device first")
        with torch.cuda.device(dev):
            enc = self._ensure_packed()
            frames = frames.contiguous().float()
            B = frames.shape[0]
            if out is None:
                out = torch.empty(B, enc.tokens, self.embed_dim, dtype=torch.float32, pin_memory=True)
            scratch = enc.host_scratch(B, dev)
            st = enc.lib.vidil_vit_forward_host(enc.handle, frames.data_ptr(), B, out.data_ptr(), scratch.data_ptr(),
                                                scratch.numel(), torch.cuda.current_stream().cuda_stream)
            _lib.check(st, "vidil_vit_forward_host")
        return out

    @torch.no_grad()
    def encode_host_stream(self, batches, outs=None, half_tokens: bool = False):
        """Encode a stream of host batches with copies overlapped with compute.  half_tokens: deliver the tokens in the
        16-bit operand type (token_dtype16) — half the device-to-host bytes.

        `batches`: iterable of CPU fp32 tensors [B,3,S,S] (pinned for full PCIe rate; B may vary — a ragged last batch
        reuses the slots of the full ones, a larger batch drains the pipeline and re-binds it).  Yields one CPU
        tensor [B, N+1, D] per batch, in order.  `outs`: optional pair of pinned output tensors to cycle through (the
        yielded tensor is then only valid until two batches later).  Batch k+1's H2D runs during batch k's forward and
        batch k's D2H during batch k+1's (vidil_encoder_host_submit / _wait, two slots)."""
        dev = self.cls_token.device
        if dev.type != "cuda":
            raise RuntimeError("vidil_b200: the module must be moved to a CUDA device first")
        with torch.cuda.device(dev):
            enc = self._ensure_packed()
            pending = []  # (slot, out)
            k = 0
            for frames in batches:
                if frames.is_cuda:
                    raise RuntimeError("encode_host_stream takes host tensors")
                _check_frames(frames, self.img_size)
                frames = frames.contiguous().float()
                slot = k & 1
                if len(pending) == 2:  # the slot about to be reused must have delivered its result
                    s0, o0, _ = pending.pop(0)
                    enc.host_wait(s0)
                    yield o0
                out = outs[slot] if outs is not None else torch.empty(
                    frames.shape[0], enc.tokens, self.embed_dim, dtype=self.token_dtype16 if half_tokens else torch.float32,
                    pin_memory=True)
                if out.dtype != (self.token_dtype16 if half_tokens else torch.float32):
                    raise RuntimeError(f"encode_host_stream: output buffers must be {self.token_dtype16 if half_tokens else torch.float32}")
                enc.host_submit(frames, out, slot, dev)
                pending.append((slot, out, frames))  # keep the input alive until its copy has run
                k += 1
            for s0, o0, _ in pending:
                enc.host_wait(s0)
                yield o0

    @torch.jit.ignore()
    def load_pretrained(self, checkpoint_path, prefix=''):
        raise NotImplementedError("Flax .npz loading (models/vit.py:201-278) is not on the inference path; "
                                  "load BLIP checkpoints through load_state_dict as models/blip.py does")


def interpolate_pos_embed(pos_embed_checkpoint, visual_encoder):
    """Same contract as models/vit.py:281-305 (called from models/blip.py:load_checkpoint): bicubic-resize the
    grid part of a checkpoint's position table to this encoder's grid; the class-token row is kept."""
    embedding_size = pos_embed_checkpoint.shape[-1]
    num_patches = visual_encoder.patch_embed.num_patches
    num_extra_tokens = visual_encoder.pos_embed.shape[-2] - num_patches
    orig_size = int((pos_embed_checkpoint.shape[-2] - num_extra_tokens) ** 0.5)
    new_size = int(num_patches ** 0.5)
    if orig_size == new_size:
        return pos_embed_checkpoint
    extra_tokens = pos_embed_checkpoint[:, :num_extra_tokens]
    grid = pos_embed_checkpoint[:, num_extra_tokens:].reshape(-1, orig_size, orig_size, embedding_size).permute(0, 3, 1, 2)
    grid = F.interpolate(grid, size=(new_size, new_size), mode='bicubic', align_corners=False)
    grid = grid.permute(0, 2, 3, 1).flatten(1, 2)
    print('reshape position embedding from %d to %d' % (orig_size ** 2, new_size ** 2))
    return torch.cat((extra_tokens, grid), dim=1)
