"""Drop-in for the parts of the reference's `models/med.py` that the CapFilt path runs: `BertModel` as the multimodal
text encoder of BLIP_ITM (models/blip_itm.py:29,49-55) and `BertLMHeadModel` as the caption decoder of BLIP_Decoder
(models/blip.py:97,150-158).

The modules keep the reference's parameter tree (so BLIP checkpoints load through `load_state_dict` with the same keys:
`bert.embeddings.word_embeddings.weight`, `bert.encoder.layer.N.{attention,crossattention}.self.{query,key,value}.*`,
`...output.{dense,LayerNorm}.*`, `intermediate.dense.*`, `cls.predictions.*`) and the reference's call surface
(`text_encoder(input_ids, attention_mask=..., encoder_hidden_states=..., return_dict=True).last_hidden_state`,
`text_decoder.generate(input_ids=..., max_length=..., min_length=..., num_beams=..., eos_token_id=..., pad_token_id=...,
encoder_hidden_states=...)`).  The arithmetic is not here: weights are packed once into a native `vidil_med` handle and the
sm_100a kernels of libvidil_b200.so run the stack (include/vidil_b200.h: vidil_med_forward, vidil_med_generate).

Inference only, CUDA only — there is no CPU implementation and no fallback.
"""
from __future__ import annotations

import ctypes
import os
from types import SimpleNamespace

import torch
import torch.nn as nn

from . import _lib


class BertConfig(SimpleNamespace):
    """The fields of configs/med_config.json that the inference path reads (transformers' BertConfig carries many more)."""

    def __init__(self, vocab_size=30524, hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072,
                 max_position_embeddings=512, layer_norm_eps=1e-12, encoder_width=768, hidden_act="gelu", pad_token_id=0,
                 add_cross_attention=True, **ignored):
        super().__init__(vocab_size=vocab_size, hidden_size=hidden_size, num_hidden_layers=num_hidden_layers,
                         num_attention_heads=num_attention_heads, intermediate_size=intermediate_size,
                         max_position_embeddings=max_position_embeddings, layer_norm_eps=layer_norm_eps,
                         encoder_width=encoder_width, hidden_act=hidden_act, pad_token_id=pad_token_id,
                         add_cross_attention=add_cross_attention)
        if hidden_act != "gelu" or not add_cross_attention:
            raise ValueError("vidil_b200.med supports configs/med_config.json: hidden_act 'gelu', add_cross_attention true")

    @classmethod
    def from_json_file(cls, path):
        import json
        with open(path) as f:
            return cls(**json.load(f))


# ---- parameter holders with the reference's module tree (models/med.py:51-541) -----------------------------------------
class _Embeddings(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.word_embeddings = nn.Embedding(c.vocab_size, c.hidden_size, padding_idx=c.pad_token_id)
        self.position_embeddings = nn.Embedding(c.max_position_embeddings, c.hidden_size)
        self.LayerNorm = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)
        self.register_buffer("position_ids", torch.arange(c.max_position_embeddings).expand((1, -1)))


class _SelfAttention(nn.Module):
    def __init__(self, c, is_cross):
        super().__init__()
        kv_in = c.encoder_width if is_cross else c.hidden_size
        self.query = nn.Linear(c.hidden_size, c.hidden_size)
        self.key = nn.Linear(kv_in, c.hidden_size)
        self.value = nn.Linear(kv_in, c.hidden_size)


class _SelfOutput(nn.Module):
    def __init__(self, c, in_features=None):
        super().__init__()
        self.dense = nn.Linear(in_features or c.hidden_size, c.hidden_size)
        self.LayerNorm = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)


class _Attention(nn.Module):
    def __init__(self, c, is_cross=False):
        super().__init__()
        self.self = _SelfAttention(c, is_cross)
        self.output = _SelfOutput(c)


class _Intermediate(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.dense = nn.Linear(c.hidden_size, c.intermediate_size)


class _Layer(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.attention = _Attention(c)
        self.crossattention = _Attention(c, is_cross=True)
        self.intermediate = _Intermediate(c)
        self.output = _SelfOutput(c, in_features=c.intermediate_size)


class _Encoder(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.layer = nn.ModuleList([_Layer(c) for _ in range(c.num_hidden_layers)])


class _HeadTransform(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.dense = nn.Linear(c.hidden_size, c.hidden_size)
        self.LayerNorm = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)


class _LMPredictionHead(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.transform = _HeadTransform(c)
        self.decoder = nn.Linear(c.hidden_size, c.vocab_size, bias=False)
        self.bias = nn.Parameter(torch.zeros(c.vocab_size))


class _OnlyMLMHead(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.predictions = _LMPredictionHead(c)


def _init_bert(module, std=0.02):
    """BertPreTrainedModel._init_weights (med.py:553-563)."""
    if isinstance(module, (nn.Linear, nn.Embedding)):
        module.weight.data.normal_(mean=0.0, std=std)
    elif isinstance(module, nn.LayerNorm):
        module.bias.data.zero_()
        module.weight.data.fill_(1.0)
    if isinstance(module, nn.Linear) and module.bias is not None:
        module.bias.data.zero_()


# ---- native handle ---------------------------------------------------------------------------------------------------
class NativeMed:
    """Owns one `vidil_med*` and the torch-allocated workspace its calls run in."""

    def __init__(self, cfg: _lib.MedCfg):
        self.lib = _lib.load()
        self.cfg = cfg
        self.handle = ctypes.c_void_p()
        _lib.check(self.lib.vidil_med_create(ctypes.byref(cfg), ctypes.byref(self.handle)), "vidil_med_create")
        self._ws = None

    def __del__(self):
        try:
            if self.handle:
                self.lib.vidil_med_destroy(self.handle)
                self.handle = None
        except Exception:  # noqa: BLE001 - interpreter teardown
            pass

    def load(self, name: str, tensor: torch.Tensor) -> None:
        t = tensor.detach().to(dtype=torch.float32).contiguous()
        if not t.is_cuda:
            raise RuntimeError("vidil_b200: parameters must live on a CUDA device (no CPU path exists)")
        st = self.lib.vidil_med_load(self.handle, name.encode(), t.data_ptr(), t.numel(), torch.cuda.current_stream().cuda_stream)
        _lib.check(st, f"vidil_med_load({name})")

    def set_profiling(self, enable: bool) -> None:
        _lib.check(self.lib.vidil_med_set_profiling(self.handle, int(enable)), "vidil_med_set_profiling")

    def read_profile(self) -> dict:
        """{class: {ms, flops, bytes, launches}} of the kernels enqueued since the last read (synchronises)."""
        st = _lib.KernelStats()
        _lib.check(self.lib.vidil_med_read_profile(self.handle, ctypes.byref(st)), "vidil_med_read_profile")
        return {name: dict(ms=st.ms[i], flops=st.flops[i], bytes=st.bytes[i], launches=int(st.launches[i]))
                for i, name in enumerate(_lib.KCLASSES)}

    def workspace(self, need: int, device) -> torch.Tensor:
        if self._ws is None or self._ws.numel() < need or self._ws.device != device:
            buf = torch.empty(need + 1024, dtype=torch.uint8, device=device)
            off = (-buf.data_ptr()) % 1024
            self._ws = buf[off:off + need]
        return self._ws


def _i32(t: torch.Tensor, device) -> torch.Tensor:
    return t.to(device=device, dtype=torch.int32).contiguous()


class BertModel(nn.Module):
    """models/med.py:566-809 `BertModel(config, add_pooling_layer=False)` on the native path.  `cls_head`: an optional
    nn.Linear applied to token 0 inside the same native call (BLIP_ITM passes its itm_head)."""

    def __init__(self, config, add_pooling_layer=False, compute_dtype="bf16", lm_head=None):
        super().__init__()
        if add_pooling_layer:
            raise ValueError("the CapFilt path builds BertModel(add_pooling_layer=False) (blip_itm.py:29)")
        self.config = config
        self.compute_dtype = compute_dtype
        self.embeddings = _Embeddings(config)
        self.encoder = _Encoder(config)
        self.apply(_init_bert)
        self._native = None
        self._packed_sig = None
        self._lm_head = [lm_head] if lm_head is not None else None      # list: not a registered sub-module
        self._cls_head = None

    # -- packing ---------------------------------------------------------------------------------------------------------
    def attach_cls_head(self, linear: nn.Linear) -> None:
        if self._cls_head is None or self._cls_head[0] is not linear:
            self._cls_head = [linear]
            self._native = None

    def _params_for_signature(self):
        ps = list(self.parameters())
        if self._lm_head:
            ps += list(self._lm_head[0].parameters())
        if self._cls_head:
            ps += list(self._cls_head[0].parameters())
        return ps

    def _ensure_packed(self) -> NativeMed:
        c = self.config
        sig = tuple((p.data_ptr(), p._version) for p in self._params_for_signature())
        if self._native is None:
            cfg = _lib.MedCfg(vocab_size=c.vocab_size, max_positions=c.max_position_embeddings, hidden=c.hidden_size,
                              depth=c.num_hidden_layers, num_heads=c.num_attention_heads, mlp_dim=c.intermediate_size,
                              encoder_width=c.encoder_width, ln_eps=c.layer_norm_eps, lm_head=1 if self._lm_head else 0,
                              cls_out=self._cls_head[0].out_features if self._cls_head else 0,
                              dtype=_lib.DTYPES[self.compute_dtype], cta_group=0)
            self._native = NativeMed(cfg)
            self._packed_sig = None
        if sig != self._packed_sig:
            n = self._native
            e = self.embeddings
            n.load("word_embeddings", e.word_embeddings.weight)
            n.load("position_embeddings", e.position_embeddings.weight)
            n.load("emb_ln.weight", e.LayerNorm.weight)
            n.load("emb_ln.bias", e.LayerNorm.bias)
            for i, ly in enumerate(self.encoder.layer):
                p = f"layer.{i}."
                a, x = ly.attention, ly.crossattention
                n.load(p + "self.qkv.weight", torch.cat([a.self.query.weight, a.self.key.weight, a.self.value.weight], 0))
                n.load(p + "self.qkv.bias", torch.cat([a.self.query.bias, a.self.key.bias, a.self.value.bias], 0))
                n.load(p + "self.out.weight", a.output.dense.weight)
                n.load(p + "self.out.bias", a.output.dense.bias)
                n.load(p + "self.ln.weight", a.output.LayerNorm.weight)
                n.load(p + "self.ln.bias", a.output.LayerNorm.bias)
                n.load(p + "cross.q.weight", x.self.query.weight)
                n.load(p + "cross.q.bias", x.self.query.bias)
                n.load(p + "cross.kv.weight", torch.cat([x.self.key.weight, x.self.value.weight], 0))
                n.load(p + "cross.kv.bias", torch.cat([x.self.key.bias, x.self.value.bias], 0))
                n.load(p + "cross.out.weight", x.output.dense.weight)
                n.load(p + "cross.out.bias", x.output.dense.bias)
                n.load(p + "cross.ln.weight", x.output.LayerNorm.weight)
                n.load(p + "cross.ln.bias", x.output.LayerNorm.bias)
                n.load(p + "ffn.fc1.weight", ly.intermediate.dense.weight)
                n.load(p + "ffn.fc1.bias", ly.intermediate.dense.bias)
                n.load(p + "ffn.fc2.weight", ly.output.dense.weight)
                n.load(p + "ffn.fc2.bias", ly.output.dense.bias)
                n.load(p + "ffn.ln.weight", ly.output.LayerNorm.weight)
                n.load(p + "ffn.ln.bias", ly.output.LayerNorm.bias)
            if self._lm_head:
                h = self._lm_head[0].predictions
                n.load("head.dense.weight", h.transform.dense.weight)
                n.load("head.dense.bias", h.transform.dense.bias)
                n.load("head.ln.weight", h.transform.LayerNorm.weight)
                n.load("head.ln.bias", h.transform.LayerNorm.bias)
                n.load("head.decoder.weight", h.decoder.weight)
                n.load("head.decoder.bias", h.bias)
            if self._cls_head:
                n.load("cls.weight", self._cls_head[0].weight)
                n.load("cls.bias", self._cls_head[0].bias)
            _lib.check(n.lib.vidil_med_check_loaded(n.handle), "vidil_med_check_loaded")
            self._packed_sig = sig
        return self._native

    # -- whole-sequence forward ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def run(self, input_ids, attention_mask, encoder_hidden_states, frame_of_seq=None, causal=False, want_hidden=True,
            want_logits=False, want_cls=False, seqs_per_frame=0):
        """One vidil_med_forward call; returns (hidden | None, logits | None, cls | None).  frame_of_seq [n_seq]: the frame each
        sequence attends to; or seqs_per_frame > 0: sequences come frame-major (sequence i reads frame i // seqs_per_frame) and
        the cross-attention batches a frame's sequences into one query group."""
        enc = encoder_hidden_states
        if not enc.is_cuda:
            raise RuntimeError("vidil_b200: encoder_hidden_states must be on a CUDA device (no CPU path exists)")
        dev = enc.device
        with torch.cuda.device(dev):
            n = self._ensure_packed()
            c = self.config
            enc = enc.contiguous().float()
            F_, Nv = enc.shape[0], enc.shape[1]
            ids = _i32(input_ids, dev)
            mask = _i32(attention_mask, dev) if (attention_mask is not None and not causal) else None
            if mask is not None and not want_hidden and not want_logits:
                # Only token 0 is read (itm_head, blip_itm.py:56) and masked keys carry zero weight, so columns that are padding
                # in EVERY sequence (tokenizer padding='max_length', :46) cannot influence the result: drop them.
                keep = int((mask != 0).any(dim=0).nonzero().max().item()) + 1 if bool((mask != 0).any()) else 1
                if keep < ids.shape[1]:
                    ids, mask = ids[:, :keep].contiguous(), mask[:, :keep].contiguous()
            S, T = ids.shape
            fos = _i32(frame_of_seq, dev) if frame_of_seq is not None else None
            hidden = torch.empty(S, T, c.hidden_size, dtype=torch.float32, device=dev) if want_hidden else None
            logits = torch.empty(S, T, c.vocab_size, dtype=torch.float32, device=dev) if want_logits else None
            cls = torch.empty(S, self._cls_head[0].out_features, dtype=torch.float32, device=dev) if want_cls else None
            need = n.lib.vidil_med_forward_workspace_bytes(n.handle, S, T, F_, Nv)
            ws = n.workspace(need, dev)
            ptr = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
            st = n.lib.vidil_med_forward(n.handle, enc.data_ptr(), F_, Nv, ids.data_ptr(), ptr(mask), ptr(fos), int(seqs_per_frame), S, T,
                                         1 if causal else 0, ptr(hidden), ptr(logits), ptr(cls), ws.data_ptr(), ws.numel(),
                                         torch.cuda.current_stream().cuda_stream)
            _lib.check(st, "vidil_med_forward")
        return hidden, logits, cls

    def forward(self, input_ids=None, attention_mask=None, encoder_hidden_states=None, encoder_attention_mask=None,
                return_dict=True, is_decoder=False, mode="multimodal", **unsupported):
        """The call of blip_itm.py:49-54.  The image mask is all ones on this path (blip_itm.py:44) and is not read."""
        if mode != "multimodal" or encoder_hidden_states is None:
            raise NotImplementedError("only mode='multimodal' with encoder_hidden_states is on the CapFilt path")
        for k, v in unsupported.items():
            if v is not None and v is not False:
                raise NotImplementedError(f"BertModel.forward argument {k} is not on the CapFilt path")
        hidden, _, _ = self.run(input_ids, attention_mask, encoder_hidden_states, causal=bool(is_decoder))
        out = SimpleNamespace(last_hidden_state=hidden, pooler_output=None)
        return out if return_dict else (hidden, None)


class BertLMHeadModel(nn.Module):
    """models/med.py:811-955 on the native path: teacher-forced logits (`forward`) and beam-search `generate`."""

    def __init__(self, config, compute_dtype="bf16"):
        super().__init__()
        self.config = config
        self.cls = _OnlyMLMHead(config)
        self.cls.apply(_init_bert)
        self.bert = BertModel(config, add_pooling_layer=False, compute_dtype=compute_dtype, lm_head=self.cls)
        self._register_load_state_dict_pre_hook(self._fill_tied_decoder)

    @staticmethod
    def _fill_tied_decoder(state_dict, prefix, *args):
        # the reference ties decoder.weight to the word embeddings (med.py:533-535 comment); a checkpoint saved from a tied
        # model may carry only one of the two keys
        w, d = prefix + "bert.embeddings.word_embeddings.weight", prefix + "cls.predictions.decoder.weight"
        if d not in state_dict and w in state_dict:
            state_dict[d] = state_dict[w]
        state_dict.pop(prefix + "cls.predictions.decoder.bias", None)   # alias of cls.predictions.bias (med.py:538)

    @torch.no_grad()
    def forward(self, input_ids=None, attention_mask=None, encoder_hidden_states=None, encoder_attention_mask=None,
                return_dict=True, is_decoder=True, return_logits=False, **unsupported):
        """Logits for every position of fully given sequences (no padding mask: the generation path feeds all-ones)."""
        if attention_mask is not None and not bool(torch.all(attention_mask != 0)):
            raise NotImplementedError("decoder forward with padded sequences is not on the CapFilt path")
        _, logits, _ = self.bert.run(input_ids, None, encoder_hidden_states, causal=True, want_hidden=False, want_logits=True)
        if return_logits:
            return logits[:, :-1, :].contiguous()
        return SimpleNamespace(logits=logits, loss=None, past_key_values=None) if return_dict else (logits,)

    @torch.no_grad()
    def generate(self, input_ids=None, max_length=20, min_length=0, num_beams=1, eos_token_id=None, pad_token_id=0,
                 repetition_penalty=1.0, length_penalty=1.0, do_sample=False, top_p=1.0, top_k=50, num_return_sequences=1,
                 encoder_hidden_states=None, encoder_attention_mask=None, return_scores=False, generator=None, uniforms=None,
                 **unsupported):
        """transformers v4.15 `generate(...)` as blip.py:139-158 calls it: beam search (do_sample False) or nucleus sampling
        (do_sample True: repetition penalty, MinLength, top-k — the PretrainedConfig default 50 BLIP inherits — and top-p, then
        one draw per step; `uniforms` [max_length - prompt_len, B] in [0, 1) or, if None, torch.rand from `generator` / the
        global CPU generator).  `encoder_hidden_states` may be the per-frame tokens [B, N, E] or, as the reference passes them
        for beam search, already repeat_interleaved over the beams [B*num_beams, N, E] (blip.py:130) — then every num_beams-th
        row is used.  Returns int64 [B, L] padded with pad_token_id, L = min(longest hypothesis + 1, max_length), like
        BeamSearchScorer.finalize (sampling: the longest sequence, eos included)."""
        if do_sample:
            if num_beams != 1 or num_return_sequences != 1:
                raise NotImplementedError("sampling is built as blip.py:141-148 calls it: one beam, one returned sequence")
        elif repetition_penalty != 1.0:
            raise NotImplementedError("repetition_penalty != 1.0 is not on the beam-search path of run_video_CapFilt.py:102")
        if eos_token_id is None:
            raise ValueError("eos_token_id is required")
        enc = encoder_hidden_states
        B = input_ids.shape[0]
        if enc.shape[0] == B * num_beams and num_beams > 1:
            enc = enc[::num_beams]
        if enc.shape[0] != B:
            raise ValueError(f"{enc.shape[0]} image-token rows for {B} prompts")
        prompt = input_ids[0].to("cpu", torch.int32).contiguous()
        if not bool((input_ids.cpu() == prompt.to(input_ids.dtype)).all()):
            raise NotImplementedError("generate expects the same prompt for every frame (blip.py:133-137)")
        dev = enc.device
        with torch.cuda.device(dev):
            n = self.bert._ensure_packed()
            enc = enc.contiguous().float()
            Nv, Lp = enc.shape[1], prompt.numel()
            if do_sample:
                if uniforms is None:
                    uniforms = torch.rand(max_length - Lp, B, generator=generator)
                uni = uniforms.to(dev, torch.float32)
                if tuple(uni.shape) != (max_length - Lp, B):
                    raise ValueError(f"uniforms must be [max_length - prompt_len, B] = [{max_length - Lp}, {B}]")
            toks = torch.empty(B, max_length, dtype=torch.int32, device=dev)
            lens = torch.empty(B, dtype=torch.int32, device=dev)
            scores = torch.empty(B, dtype=torch.float32, device=dev)
            # the workspace holds the cross-attention K/V of every frame for every layer (12 x frames x tokens x 3 KB): bound it,
            # and run the frames in chunks when a batch would need more (captions do not depend on the batch they are in)
            budget = int(float(os.environ.get("VIDIL_MED_WORKSPACE_GB", "32")) * (1 << 30))
            chunk = B
            while chunk > 1 and n.lib.vidil_med_generate_workspace_bytes(n.handle, chunk, Nv, num_beams, max_length, Lp) > budget:
                chunk = (chunk + 1) // 2
            for b0 in range(0, B, chunk):
                nb = min(chunk, B - b0)
                need = n.lib.vidil_med_generate_workspace_bytes(n.handle, nb, Nv, num_beams, max_length, Lp)
                ws = n.workspace(need, dev)
                if do_sample:
                    u = uni[:, b0:b0 + nb].contiguous()
                    st = n.lib.vidil_med_sample(n.handle, enc[b0:b0 + nb].data_ptr(), nb, Nv, prompt.data_ptr(), Lp, max_length, min_length,
                                                eos_token_id, pad_token_id, int(top_k), float(top_p), float(repetition_penalty),
                                                u.data_ptr(), toks[b0:b0 + nb].data_ptr(), lens[b0:b0 + nb].data_ptr(),
                                                scores[b0:b0 + nb].data_ptr(), ws.data_ptr(), ws.numel(),
                                                torch.cuda.current_stream().cuda_stream)
                    _lib.check(st, "vidil_med_sample")
                    continue
                st = n.lib.vidil_med_generate(n.handle, enc[b0:b0 + nb].data_ptr(), nb, Nv, prompt.data_ptr(), Lp, num_beams, max_length,
                                              min_length, eos_token_id, pad_token_id, float(length_penalty), toks[b0:b0 + nb].data_ptr(),
                                              lens[b0:b0 + nb].data_ptr(), scores[b0:b0 + nb].data_ptr(), ws.data_ptr(), ws.numel(),
                                              torch.cuda.current_stream().cuda_stream)
                _lib.check(st, "vidil_med_generate")
            width = int(lens.max().item())          # finalize: sent_max_len = min(max(sent_lengths) + 1, max_length)
            out = toks[:, :width].long()
        return (out, scores, lens) if return_scores else out
