"""Drop-in for the CLIP image path of `run_visual_tokenization.py`.

The reference builds `transformers.CLIPModel` (:347-350) and calls `model(**processor_out)`, reading
`.image_embeds` for frames (:138-142) and `.text_embeds` for the ontology phrases (:88-92).
`CLIPVisionB200` runs the vision tower, visual_projection and L2 normalisation natively
(vidil_clip_forward), `CLIPTextB200` the text tower, text_projection and normalisation (vidil_clip_text_forward,
SURVEY.md §8f "next" #3); `VidilCLIPModel` wraps an existing CLIPModel so that the same `model(**inputs)` call
returns native `image_embeds` / `text_embeds`.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch
import torch.nn as nn

from . import _lib
from .vision_transformer import NativeEncoder, _check_frames


# Default operand type of the CLIP towers: fp16.  Measured end to end at the north-star shape (bench.py `sim` record, 256 frames x
# 10 000 phrases, top-5, against the fp32 tower): bf16 operands flip 93 of 1 280 index positions (embedding error 8.7e-4), fp16
# operands 1 (embedding error 9.7e-5, the flipped pair 2.6e-5 apart in fp32 score).  CLIP's activations sit well inside fp16
# range (the residual stream stays fp32), and the tensor-core rate is the same.
class CLIPVisionB200(nn.Module):
    """CLIP vision tower + projection on the native path.  Parameters use transformers' CLIPModel key names
    (`vision_model.*`, `visual_projection.weight`) so `load_state_dict(hf_model.state_dict(), strict=False)`
    populates it."""

    def __init__(self, hidden_size=1024, intermediate_size=4096, num_hidden_layers=24, num_attention_heads=16,
                 image_size=224, patch_size=14, projection_dim=768, layer_norm_eps=1e-5, hidden_act="quick_gelu",
                 compute_dtype="fp16", cta_group=0):
        super().__init__()
        if hidden_size != 64 * num_attention_heads:
            raise ValueError("head_dim must be 64")
        if hidden_act not in ("quick_gelu", "gelu"):
            raise ValueError(f"unsupported hidden_act {hidden_act!r}")
        self.cfg = dict(hidden_size=hidden_size, intermediate_size=intermediate_size,
                        num_hidden_layers=num_hidden_layers, num_attention_heads=num_attention_heads,
                        image_size=image_size, patch_size=patch_size, projection_dim=projection_dim,
                        layer_norm_eps=layer_norm_eps, hidden_act=hidden_act)
        self.compute_dtype, self.cta_group = compute_dtype, cta_group
        D, I = hidden_size, intermediate_size
        P = (image_size // patch_size) ** 2

        def p(*shape):
            return nn.Parameter(torch.zeros(*shape))

        params = {
            "vision_model.embeddings.class_embedding": p(D),
            "vision_model.embeddings.patch_embedding.weight": p(D, 3, patch_size, patch_size),
            "vision_model.embeddings.position_embedding.weight": p(P + 1, D),
            "vision_model.pre_layrnorm.weight": p(D), "vision_model.pre_layrnorm.bias": p(D),
            "vision_model.post_layernorm.weight": p(D), "vision_model.post_layernorm.bias": p(D),
            "visual_projection.weight": p(projection_dim, D),
        }
        for i in range(num_hidden_layers):
            pre = f"vision_model.encoder.layers.{i}."
            for ln in ("layer_norm1", "layer_norm2"):
                params[pre + ln + ".weight"] = p(D)
                params[pre + ln + ".bias"] = p(D)
            for nm in ("q_proj", "k_proj", "v_proj", "out_proj"):
                params[pre + f"self_attn.{nm}.weight"] = p(D, D)
                params[pre + f"self_attn.{nm}.bias"] = p(D)
            params[pre + "mlp.fc1.weight"], params[pre + "mlp.fc1.bias"] = p(I, D), p(I)
            params[pre + "mlp.fc2.weight"], params[pre + "mlp.fc2.bias"] = p(D, I), p(D)
        # nn.ParameterDict forbids dots in keys; keep the HF names via a flat, escaped registry
        self._names = list(params)
        for name, prm in params.items():
            self.register_parameter(name.replace(".", "__"), prm)
        self._native = None
        self._packed_sig = None

    # state_dict with transformers' key names -----------------------------------------------------
    def state_dict(self, *args, **kwargs):
        sd = super().state_dict(*args, **kwargs)
        return {k.replace("__", "."): v for k, v in sd.items()}

    def load_state_dict(self, state_dict, strict=True, assign=False):
        own = {n: getattr(self, n.replace(".", "__")) for n in self._names}
        missing = [n for n in own if n not in state_dict]
        unexpected = [k for k in state_dict if k not in own]
        if strict and (missing or unexpected):
            raise RuntimeError(f"missing keys {missing[:4]}..., unexpected keys {unexpected[:4]}...")
        with torch.no_grad():
            for n, prm in own.items():
                if n in state_dict:
                    prm.copy_(state_dict[n])
        return SimpleNamespace(missing_keys=missing, unexpected_keys=unexpected)

    @classmethod
    def from_hf(cls, hf_model, compute_dtype="fp16", cta_group=0):
        """Build from a transformers CLIPModel / CLIPVisionModelWithProjection instance."""
        vc = hf_model.config.vision_config if hasattr(hf_model.config, "vision_config") else hf_model.config
        m = cls(hidden_size=vc.hidden_size, intermediate_size=vc.intermediate_size,
                num_hidden_layers=vc.num_hidden_layers, num_attention_heads=vc.num_attention_heads,
                image_size=vc.image_size, patch_size=vc.patch_size,
                projection_dim=getattr(hf_model.config, "projection_dim", vc.projection_dim),
                layer_norm_eps=vc.layer_norm_eps, hidden_act=vc.hidden_act, compute_dtype=compute_dtype,
                cta_group=cta_group)
        m.load_state_dict(hf_model.state_dict(), strict=False)
        dev = next(hf_model.parameters()).device
        return m.to(dev).eval()

    # native handle ------------------------------------------------------------------------------------
    def _packed_tensors(self):
        """(native name, tensor) pairs: HF names mapped to the ABI's names; q/k/v fused to one [3D, D] matrix in
        the [3, H, 64] row order the attention kernel reads."""
        g = lambda n: getattr(self, n.replace(".", "__"))  # noqa: E731
        v = "vision_model."
        c = self.cfg
        yield "cls_token", g(v + "embeddings.class_embedding")
        yield "pos_embed", g(v + "embeddings.position_embedding.weight")
        yield "patch_embed.proj.weight", g(v + "embeddings.patch_embedding.weight")
        yield "pre_norm.weight", g(v + "pre_layrnorm.weight")
        yield "pre_norm.bias", g(v + "pre_layrnorm.bias")
        for i in range(c["num_hidden_layers"]):
            s, d = f"{v}encoder.layers.{i}.", f"blocks.{i}."
            yield d + "norm1.weight", g(s + "layer_norm1.weight")
            yield d + "norm1.bias", g(s + "layer_norm1.bias")
            yield d + "attn.qkv.weight", torch.cat([g(s + f"self_attn.{n}.weight") for n in ("q_proj", "k_proj", "v_proj")])
            yield d + "attn.qkv.bias", torch.cat([g(s + f"self_attn.{n}.bias") for n in ("q_proj", "k_proj", "v_proj")])
            yield d + "attn.proj.weight", g(s + "self_attn.out_proj.weight")
            yield d + "attn.proj.bias", g(s + "self_attn.out_proj.bias")
            yield d + "norm2.weight", g(s + "layer_norm2.weight")
            yield d + "norm2.bias", g(s + "layer_norm2.bias")
            for fc in ("fc1", "fc2"):
                yield d + f"mlp.{fc}.weight", g(s + f"mlp.{fc}.weight")
                yield d + f"mlp.{fc}.bias", g(s + f"mlp.{fc}.bias")
        yield "norm.weight", g(v + "post_layernorm.weight")
        yield "norm.bias", g(v + "post_layernorm.bias")
        yield "head.proj.weight", g("visual_projection.weight")

    def _ensure_packed(self) -> NativeEncoder:
        c = self.cfg
        sig = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._native is None:
            cfg = _lib.EncoderCfg(img_size=c["image_size"], patch_size=c["patch_size"], embed_dim=c["hidden_size"],
                                  depth=c["num_hidden_layers"], num_heads=c["num_attention_heads"],
                                  mlp_dim=c["intermediate_size"], ln_eps=c["layer_norm_eps"],
                                  act=_lib.ACT_QUICK_GELU if c["hidden_act"] == "quick_gelu" else _lib.ACT_GELU_ERF,
                                  patch_bias=0, pre_ln=1, proj_dim=c["projection_dim"],
                                  dtype=_lib.DTYPES[self.compute_dtype], cta_group=self.cta_group)
            self._native = NativeEncoder(cfg)
            self._packed_sig = None
        if sig != self._packed_sig:
            with torch.no_grad():
                for name, t in self._packed_tensors():
                    self._native.load(name, t)
            _lib.check(self._native.lib.vidil_encoder_check_loaded(self._native.handle), "vidil_encoder_check_loaded")
            self._packed_sig = sig
        return self._native

    @torch.no_grad()
    def forward(self, pixel_values: torch.Tensor, return_hidden: bool = False):
        """pixel_values [F,3,S,S] CUDA -> image_embeds [F, projection_dim] fp32, unit L2 norm.
        With return_hidden also the vision tower's last_hidden_state [F, P+1, D]."""
        if not pixel_values.is_cuda:
            raise RuntimeError("vidil_b200: pixel_values must be on a CUDA device; use encode_host() for host buffers")
        c = self.cfg
        _check_frames(pixel_values, c["image_size"])
        with torch.cuda.device(pixel_values.device):
            enc = self._ensure_packed()
            x = pixel_values.contiguous().float()
            B = x.shape[0]
            emb = torch.empty(B, c["projection_dim"], dtype=torch.float32, device=x.device)
            hid = torch.empty(B, enc.tokens, c["hidden_size"], dtype=torch.float32, device=x.device) if return_hidden else None
            if B > 0:
                ws = enc.workspace(B, x.device)
                st = enc.lib.vidil_clip_forward(enc.handle, x.data_ptr(), B, emb.data_ptr(),
                                                hid.data_ptr() if hid is not None else None, ws.data_ptr(), ws.numel(),
                                                torch.cuda.current_stream().cuda_stream)
                _lib.check(st, "vidil_clip_forward")
        return (emb, hid) if return_hidden else emb

    @torch.no_grad()
    def encode_host(self, pixel_values: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """Host-buffer call (vidil_clip_forward_host): CPU fp32 frames in, CPU image_embeds out."""
        if pixel_values.is_cuda:
            raise RuntimeError("encode_host takes host tensors")
        c = self.cfg
        _check_frames(pixel_values, c["image_size"])
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("vidil_b200: the module must be moved to a CUDA device first")
        with torch.cuda.device(dev):
            enc = self._ensure_packed()
            x = pixel_values.contiguous().float()
            B = x.shape[0]
            if out is None:
                out = torch.empty(B, c["projection_dim"], dtype=torch.float32, pin_memory=True)
            scratch = enc.host_scratch(B, dev)
            st = enc.lib.vidil_clip_forward_host(enc.handle, x.data_ptr(), B, out.data_ptr(), scratch.data_ptr(),
                                                 scratch.numel(), torch.cuda.current_stream().cuda_stream)
            _lib.check(st, "vidil_clip_forward_host")
        return out


    @torch.no_grad()
    def encode_host_stream(self, batches, outs=None):
        """Pipelined host API (vidil_encoder_host_submit / _wait, two slots): `batches` is an iterable of CPU fp32 frame
        tensors [B,3,S,S] (pinned for full PCIe rate); yields one CPU tensor of image_embeds [B, projection_dim] per batch,
        in order.  Batch k+1's H2D runs during batch k's forward."""
        c = self.cfg
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("vidil_b200: the module must be moved to a CUDA device first")
        with torch.cuda.device(dev):
            enc = self._ensure_packed()
            pending = []
            k = 0
            for frames in batches:
                if frames.is_cuda:
                    raise RuntimeError("encode_host_stream takes host tensors")
                _check_frames(frames, c["image_size"])
                frames = frames.contiguous().float()
                slot = k & 1
                if len(pending) == 2:
                    s0, o0, _ = pending.pop(0)
                    enc.host_wait(s0)
                    yield o0
                out = outs[slot] if outs is not None else torch.empty(frames.shape[0], c["projection_dim"],
                                                                      dtype=torch.float32, pin_memory=True)
                enc.host_submit(frames, out, slot, dev)
                pending.append((slot, out, frames))
                k += 1
            for s0, o0, _ in pending:
                enc.host_wait(s0)
                yield o0


    def encode_u8_stream(self, batches_u8, outs=None):
        """Decoded frames in, image_embeds out: `batches_u8` is an iterable of pinned CPU uint8 tensors [B, H, W, 3]; each
        is uploaded as bytes, pre-processed on the GPU exactly as transformers' CLIPImageProcessor would
        (vidil_clip_preprocess_frames) and encoded; uploads and downloads overlap the neighbouring batches' kernels."""
        from . import preprocess
        return preprocess.encode_u8_stream(self, batches_u8, self.cfg["image_size"], outs=outs, recipe="clip")


class CLIPTextB200(nn.Module):
    """CLIP text tower + text_projection on the native path (vidil_clip_text_forward).  Parameters use transformers'
    CLIPModel key names (`text_model.*`, `text_projection.weight`)."""

    def __init__(self, vocab_size=49408, max_position_embeddings=77, hidden_size=768, intermediate_size=3072,
                 num_hidden_layers=12, num_attention_heads=12, projection_dim=768, layer_norm_eps=1e-5,
                 hidden_act="quick_gelu", eos_token_id=49407, compute_dtype="fp16", cta_group=0):
        super().__init__()
        if hidden_size != 64 * num_attention_heads:
            raise ValueError("head_dim must be 64")
        if hidden_act not in ("quick_gelu", "gelu"):
            raise ValueError(f"unsupported hidden_act {hidden_act!r}")
        self.cfg = dict(vocab_size=vocab_size, max_position_embeddings=max_position_embeddings, hidden_size=hidden_size,
                        intermediate_size=intermediate_size, num_hidden_layers=num_hidden_layers,
                        num_attention_heads=num_attention_heads, projection_dim=projection_dim,
                        layer_norm_eps=layer_norm_eps, hidden_act=hidden_act, eos_token_id=eos_token_id)
        self.compute_dtype, self.cta_group = compute_dtype, cta_group
        D, I = hidden_size, intermediate_size

        def p(*shape):
            return nn.Parameter(torch.zeros(*shape))

        t = "text_model."
        params = {
            t + "embeddings.token_embedding.weight": p(vocab_size, D),
            t + "embeddings.position_embedding.weight": p(max_position_embeddings, D),
            t + "final_layer_norm.weight": p(D), t + "final_layer_norm.bias": p(D),
            "text_projection.weight": p(projection_dim, D),
        }
        for i in range(num_hidden_layers):
            pre = f"{t}encoder.layers.{i}."
            for ln in ("layer_norm1", "layer_norm2"):
                params[pre + ln + ".weight"], params[pre + ln + ".bias"] = p(D), p(D)
            for nm in ("q_proj", "k_proj", "v_proj", "out_proj"):
                params[pre + f"self_attn.{nm}.weight"], params[pre + f"self_attn.{nm}.bias"] = p(D, D), p(D)
            params[pre + "mlp.fc1.weight"], params[pre + "mlp.fc1.bias"] = p(I, D), p(I)
            params[pre + "mlp.fc2.weight"], params[pre + "mlp.fc2.bias"] = p(D, I), p(D)
        self._names = list(params)
        for name, prm in params.items():
            self.register_parameter(name.replace(".", "__"), prm)
        self._handle = None
        self._packed_sig = None
        self._ws = None

    def __del__(self):
        try:
            if self._handle:
                _lib.load().vidil_text_encoder_destroy(self._handle)
                self._handle = None
        except Exception:  # noqa: BLE001 - interpreter teardown
            pass

    def state_dict(self, *args, **kwargs):
        sd = super().state_dict(*args, **kwargs)
        return {k.replace("__", "."): v for k, v in sd.items()}

    def load_state_dict(self, state_dict, strict=True, assign=False):
        own = {n: getattr(self, n.replace(".", "__")) for n in self._names}
        missing = [n for n in own if n not in state_dict]
        unexpected = [k for k in state_dict if k not in own]
        if strict and (missing or unexpected):
            raise RuntimeError(f"missing keys {missing[:4]}..., unexpected keys {unexpected[:4]}...")
        with torch.no_grad():
            for n, prm in own.items():
                if n in state_dict:
                    prm.copy_(state_dict[n])
        return SimpleNamespace(missing_keys=missing, unexpected_keys=unexpected)

    @classmethod
    def from_hf(cls, hf_model, compute_dtype="fp16", cta_group=0):
        tc = hf_model.config.text_config if hasattr(hf_model.config, "text_config") else hf_model.config
        m = cls(vocab_size=tc.vocab_size, max_position_embeddings=tc.max_position_embeddings, hidden_size=tc.hidden_size,
                intermediate_size=tc.intermediate_size, num_hidden_layers=tc.num_hidden_layers,
                num_attention_heads=tc.num_attention_heads,
                projection_dim=getattr(hf_model.config, "projection_dim", tc.projection_dim),
                layer_norm_eps=tc.layer_norm_eps, hidden_act=tc.hidden_act, eos_token_id=tc.eos_token_id,
                compute_dtype=compute_dtype, cta_group=cta_group)
        m.load_state_dict(hf_model.state_dict(), strict=False)
        dev = next(hf_model.parameters()).device
        return m.to(dev).eval()

    def _packed_tensors(self):
        g = lambda n: getattr(self, n.replace(".", "__"))  # noqa: E731
        t = "text_model."
        yield "token_embedding", g(t + "embeddings.token_embedding.weight")
        yield "position_embedding", g(t + "embeddings.position_embedding.weight")
        for i in range(self.cfg["num_hidden_layers"]):
            s, d = f"{t}encoder.layers.{i}.", f"blocks.{i}."
            yield d + "norm1.weight", g(s + "layer_norm1.weight")
            yield d + "norm1.bias", g(s + "layer_norm1.bias")
            yield d + "attn.qkv.weight", torch.cat([g(s + f"self_attn.{n}.weight") for n in ("q_proj", "k_proj", "v_proj")])
            yield d + "attn.qkv.bias", torch.cat([g(s + f"self_attn.{n}.bias") for n in ("q_proj", "k_proj", "v_proj")])
            yield d + "attn.proj.weight", g(s + "self_attn.out_proj.weight")
            yield d + "attn.proj.bias", g(s + "self_attn.out_proj.bias")
            yield d + "norm2.weight", g(s + "layer_norm2.weight")
            yield d + "norm2.bias", g(s + "layer_norm2.bias")
            for fc in ("fc1", "fc2"):
                yield d + f"mlp.{fc}.weight", g(s + f"mlp.{fc}.weight")
                yield d + f"mlp.{fc}.bias", g(s + f"mlp.{fc}.bias")
        yield "norm.weight", g(t + "final_layer_norm.weight")
        yield "norm.bias", g(t + "final_layer_norm.bias")
        yield "head.proj.weight", g("text_projection.weight")

    def _ensure_packed(self):
        import ctypes
        lib = _lib.load()
        c = self.cfg
        sig = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._handle is None:
            cfg = _lib.TextCfg(vocab_size=c["vocab_size"], max_positions=c["max_position_embeddings"], embed_dim=c["hidden_size"],
                               depth=c["num_hidden_layers"], num_heads=c["num_attention_heads"], mlp_dim=c["intermediate_size"],
                               ln_eps=c["layer_norm_eps"],
                               act=_lib.ACT_QUICK_GELU if c["hidden_act"] == "quick_gelu" else _lib.ACT_GELU_ERF,
                               proj_dim=c["projection_dim"], dtype=_lib.DTYPES[self.compute_dtype], cta_group=self.cta_group)
            h = ctypes.c_void_p()
            _lib.check(lib.vidil_text_encoder_create(ctypes.byref(cfg), ctypes.byref(h)), "vidil_text_encoder_create")
            self._handle = h
            self._packed_sig = None
        if sig != self._packed_sig:
            st = torch.cuda.current_stream().cuda_stream
            with torch.no_grad():
                for name, t in self._packed_tensors():
                    t = t.detach().float().contiguous()
                    if not t.is_cuda:
                        raise RuntimeError("vidil_b200: parameters must live on a CUDA device (no CPU path exists)")
                    _lib.check(lib.vidil_text_encoder_load(self._handle, name.encode(), t.data_ptr(), t.numel(), st),
                               f"vidil_text_encoder_load({name})")
            _lib.check(lib.vidil_text_encoder_check_loaded(self._handle), "vidil_text_encoder_check_loaded")
            self._packed_sig = sig
        return lib

    def eos_positions(self, input_ids: torch.Tensor) -> torch.Tensor:
        """The pooled position of each sequence, as CLIPTextTransformer.forward picks it (modeling_clip.py:564-585):
        argmax of the ids for legacy configs with eos_token_id == 2, otherwise the first eos_token_id."""
        if self.cfg["eos_token_id"] == 2:
            return input_ids.to(torch.int).argmax(dim=-1).to(torch.int32)
        return (input_ids.to(torch.int) == self.cfg["eos_token_id"]).int().argmax(dim=-1).to(torch.int32)

    @torch.no_grad()
    def forward(self, input_ids: torch.Tensor, attention_mask: torch.Tensor | None = None) -> torch.Tensor:
        """input_ids [B, L] (CUDA, integer) -> text_embeds [B, projection_dim] fp32, unit L2 norm.  attention_mask is
        accepted and ignored: under the causal mask, padding after the EOS token cannot reach the pooled EOS row."""
        if not input_ids.is_cuda:
            raise RuntimeError("vidil_b200: input_ids must be on a CUDA device (no CPU path exists)")
        if input_ids.dim() != 2:
            raise RuntimeError(f"expected input_ids of shape [B, L], got {tuple(input_ids.shape)}")
        B, L = input_ids.shape
        c = self.cfg
        if L > c["max_position_embeddings"]:
            raise ValueError(f"Sequence length must be less than max_position_embeddings (got {L} and {c['max_position_embeddings']})")
        out = torch.empty(B, c["projection_dim"], dtype=torch.float32, device=input_ids.device)
        if B == 0:
            return out
        with torch.cuda.device(input_ids.device):
            lib = self._ensure_packed()
            ids = input_ids.to(torch.int32).contiguous()
            eos = self.eos_positions(input_ids).contiguous()
            need = lib.vidil_text_encoder_workspace_bytes(self._handle, B, L)
            if self._ws is None or self._ws.numel() < need or self._ws.device != ids.device:
                self._ws = NativeEncoder._aligned(need, ids.device)
            st = lib.vidil_clip_text_forward(self._handle, ids.data_ptr(), eos.data_ptr(), B, L, out.data_ptr(),
                                             self._ws.data_ptr(), self._ws.numel(), torch.cuda.current_stream().cuda_stream)
            _lib.check(st, "vidil_clip_text_forward")
        return out


class VidilCLIPModel(nn.Module):
    """`model(**inputs)` replacement for the CLIPModel the reference builds (run_visual_tokenization.py:347):
    same keyword inputs, returns an object with `.image_embeds` / `.text_embeds` (the two fields the script reads)."""

    def __init__(self, hf_model, compute_dtype="fp16", native_text=True):
        super().__init__()
        self.vision = CLIPVisionB200.from_hf(hf_model, compute_dtype=compute_dtype)
        tc = hf_model.config.text_config
        # No library fallback on the product path: a text tower outside the native kernels' shapes (head_dim 64,
        # width a multiple of 128, <= 208 positions — every OpenAI CLIP checkpoint) is an error, not a detour
        # through transformers' eager code.
        if native_text and not (tc.hidden_size == 64 * tc.num_attention_heads and tc.hidden_size % 128 == 0
                                and tc.max_position_embeddings <= 208):
            raise RuntimeError(f"vidil_b200: unsupported CLIP text tower (hidden {tc.hidden_size}, heads "
                               f"{tc.num_attention_heads}, positions {tc.max_position_embeddings}); the native "
                               "text encoder needs head_dim 64, hidden % 128 == 0 and <= 208 positions")
        self.text = CLIPTextB200.from_hf(hf_model, compute_dtype=compute_dtype) if native_text else None

    @torch.no_grad()
    def forward(self, input_ids=None, pixel_values=None, attention_mask=None, **_unused):
        image_embeds = self.vision(pixel_values) if pixel_values is not None else None
        text_embeds = None
        if input_ids is not None:
            if self.text is None:
                raise RuntimeError("vidil_b200: this VidilCLIPModel was built with native_text=False (image tower only); "
                                   "there is no fallback text path")
            text_embeds = self.text(input_ids, attention_mask)
        return SimpleNamespace(image_embeds=image_embeds, text_embeds=text_embeds)
