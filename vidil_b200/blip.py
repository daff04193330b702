"""Mirror of the reference's `models/blip.py` / `models/blip_itm.py` on the CapFilt path.

* `create_vit` (:298-326): the ViT factory every BLIP wrapper obtains its image tower from and calls as
  `self.visual_encoder(image)` (blip.py:94,106,128; blip_itm.py:27,43).  Re-pointing this one import is enough to put every
  reference wrapper's image tower on the native path.
* `BLIP_Decoder` (blip.py:80-167, `generate` with beam search) and `BLIP_ITM` (blip_itm.py:11-58, `match_head='itm'`), with
  the reference's constructor arguments and `state_dict` keys, re-implemented here on the native text stack
  (`vidil_b200.med`: vidil_med_generate / vidil_med_forward) — they are what `blip_decoder(...)` / `blip_itm(...)` return.
  Not covered (no shipped pipeline config uses them): nucleus sampling, the ITC head, the training forward.
"""
from __future__ import annotations

from .vision_transformer import VisionTransformer, interpolate_pos_embed  # noqa: F401  (re-exported)


def create_vit(vit, image_size, use_grad_checkpointing=False, ckpt_layer=0, drop_path_rate=0, pretrained_BLIP=None,
               num_frm=-1, compute_dtype="bf16", cache_identical_inputs=False):
    """Same arguments, return value and error as models/blip.py:298-326: ('base' | 'large') -> (module, width).

    use_grad_checkpointing / ckpt_layer / drop_path_rate only matter in training and are accepted and ignored,
    as `.eval()` + `torch.no_grad()` make them inert on the reference's inference path.
    cache_identical_inputs (extra, opt-in): return the previous output when called again with the same, unmodified
    frame tensor — the filterer of run_video_CapFilt.py:110-112 does that once per caption.
    """
    del use_grad_checkpointing, ckpt_layer, pretrained_BLIP, num_frm
    visual_encoder = None
    if vit == 'base':
        vision_width = 768
        visual_encoder = VisionTransformer(img_size=image_size, patch_size=16, embed_dim=vision_width, depth=12,
                                           num_heads=12, drop_path_rate=0 or drop_path_rate,
                                           compute_dtype=compute_dtype, cache_identical_inputs=cache_identical_inputs)
    elif vit == 'large':
        vision_width = 1024
        visual_encoder = VisionTransformer(img_size=image_size, patch_size=16, embed_dim=vision_width, depth=24,
                                           num_heads=16, drop_path_rate=0.1 or drop_path_rate,
                                           compute_dtype=compute_dtype, cache_identical_inputs=cache_identical_inputs)
    if visual_encoder is None:
        raise ValueError('cannot create vit:', vit)
    return visual_encoder, vision_width


# =====================================================================================================================
# CapFilt models (run_video_CapFilt.py:139-151): the captioner BLIP_Decoder (models/blip.py:77-167) and the filterer
# BLIP_ITM (models/blip_itm.py:10-66), same constructor arguments, state_dict keys and calls; both towers native.
# =====================================================================================================================
import os  # noqa: E402

import torch  # noqa: E402
import torch.nn as nn  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from .med import BertConfig, BertLMHeadModel, BertModel  # noqa: E402

# ids of bert-base-uncased + the two tokens init_tokenizer adds (models/blip.py:283-291)
PAD_ID, CLS_ID, SEP_ID, BOS_ID, ENC_ID = 0, 101, 102, 30522, 30523
_DEFAULT_MED_CONFIG = dict(vocab_size=30524, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                           intermediate_size=3072, max_position_embeddings=512, layer_norm_eps=1e-12)   # configs/med_config.json


def init_tokenizer():
    """models/blip.py:283-291.  Needs the bert-base-uncased vocabulary on disk (transformers cache); offline without it
    this raises exactly as the reference does — pass `tokenizer=` (any object with the BertTokenizer call surface) or work
    with token ids (`generate_ids`, `forward_ids`) instead."""
    from transformers import BertTokenizer
    tokenizer = BertTokenizer.from_pretrained('bert-base-uncased')
    tokenizer.add_special_tokens({'bos_token': '[DEC]'})
    tokenizer.add_special_tokens({'additional_special_tokens': ['[ENC]']})
    tokenizer.enc_token_id = tokenizer.additional_special_tokens_ids[0]
    return tokenizer


def _med_config(med_config, vision_width):
    if isinstance(med_config, BertConfig):
        c = BertConfig(**vars(med_config))
    elif isinstance(med_config, dict):
        c = BertConfig(**med_config)
    elif isinstance(med_config, str) and os.path.isfile(med_config):
        c = BertConfig.from_json_file(med_config)
    else:
        c = BertConfig(**_DEFAULT_MED_CONFIG)     # the shipped configs/med_config.json
    c.encoder_width = vision_width               # blip.py:96, blip_itm.py:28
    return c


def _tokenizer_or_none(tokenizer):
    if tokenizer is not None:
        return tokenizer
    try:
        return init_tokenizer()
    except Exception:  # noqa: BLE001 - no vocabulary on disk (offline)
        return None


class BLIP_Decoder(nn.Module):
    def __init__(self, med_config='configs/med_config.json', image_size=384, vit='base', vit_grad_ckpt=False, vit_ckpt_layer=0,
                 prompt='a picture of ', tokenizer=None, prompt_ids=None, compute_dtype="bf16"):
        """models/blip.py:78-101.  Extra: `tokenizer` (object with BertTokenizer's call surface) and `prompt_ids` (the
        tokenised prompt [CLS] a picture of [SEP] when no tokenizer is available)."""
        super().__init__()
        self.visual_encoder, vision_width = create_vit(vit, image_size, vit_grad_ckpt, vit_ckpt_layer, compute_dtype=compute_dtype)
        self.tokenizer = _tokenizer_or_none(tokenizer)
        self.text_decoder = BertLMHeadModel(config=_med_config(med_config, vision_width), compute_dtype=compute_dtype)
        self.prompt = prompt
        if self.tokenizer is not None:
            self._prompt_ids = list(self.tokenizer(self.prompt).input_ids)
        else:
            self._prompt_ids = list(prompt_ids) if prompt_ids is not None else [CLS_ID, 1037, 3861, 1997, SEP_ID]
        self.prompt_length = len(self._prompt_ids) - 1
        self.bos_token_id = getattr(self.tokenizer, "bos_token_id", None) or BOS_ID
        self.sep_token_id = getattr(self.tokenizer, "sep_token_id", None) or SEP_ID
        self.pad_token_id = getattr(self.tokenizer, "pad_token_id", None) or PAD_ID

    def forward(self, image, caption):
        raise NotImplementedError("the LM training loss (models/blip.py:104-125) is outside the inference hot path")

    @torch.no_grad()
    def generate_ids(self, image, num_beams=3, max_length=30, min_length=10, repetition_penalty=1.0, return_scores=False,
                     sample=False, top_p=0.9, generator=None):
        """blip.py:127-158 up to the token ids: ViT -> beam search (or, sample=True, nucleus sampling with the reference's
        repetition_penalty 1.1, blip.py:139-148); returns int64 [B, L] (prompt included)."""
        image_embeds = self.visual_encoder(image)                                              # :128
        input_ids = torch.tensor([self._prompt_ids], dtype=torch.long).repeat(image.size(0), 1)  # :133-134
        input_ids[:, 0] = self.bos_token_id                                                    # :136
        input_ids = input_ids[:, :-1]                                                          # :137
        if sample:
            return self.text_decoder.generate(input_ids=input_ids, max_length=max_length, min_length=min_length, do_sample=True,
                                              top_p=top_p, num_return_sequences=1, eos_token_id=self.sep_token_id,
                                              pad_token_id=self.pad_token_id, repetition_penalty=1.1,   # :141-148
                                              encoder_hidden_states=image_embeds, return_scores=return_scores, generator=generator)
        return self.text_decoder.generate(input_ids=input_ids, max_length=max_length, min_length=min_length, num_beams=num_beams,
                                          eos_token_id=self.sep_token_id, pad_token_id=self.pad_token_id,
                                          repetition_penalty=repetition_penalty, encoder_hidden_states=image_embeds,
                                          return_scores=return_scores)

    @torch.no_grad()
    def generate(self, image, sample=False, num_beams=3, max_length=30, min_length=10, top_p=0.9, repetition_penalty=1.0,
                 generator=None):
        """Same signature and result as models/blip.py:127-167 (list of caption strings with the prompt stripped); without a
        tokenizer the captions are lists of token ids after the prompt, special tokens removed.  sample=True draws with
        torch's global CPU generator (or `generator`): reproducible under torch.manual_seed like the reference, though not
        draw-for-draw identical to torch.multinomial."""
        outputs = self.generate_ids(image, num_beams, max_length, min_length, repetition_penalty, sample=sample, top_p=top_p,
                                    generator=generator).cpu()
        captions = []
        for output in outputs:
            if self.tokenizer is not None:
                caption = self.tokenizer.decode(output, skip_special_tokens=True)               # :163
                captions.append(caption[len(self.prompt):])                                    # :164
            else:
                special = {self.pad_token_id, self.sep_token_id, self.bos_token_id, CLS_ID, ENC_ID}
                captions.append([int(t) for t in output[self.prompt_length:] if int(t) not in special])
        return captions


class BLIP_ITM(nn.Module):
    def __init__(self, med_config='configs/med_config.json', image_size=384, vit='base', vit_grad_ckpt=False, vit_ckpt_layer=0,
                 embed_dim=256, tokenizer=None, compute_dtype="bf16", cache_identical_inputs=True):
        """models/blip_itm.py:11-38.  cache_identical_inputs: filter_captions (run_video_CapFilt.py:108-112) calls this module
        once per caption with the same frame tensor; the image tower then runs once."""
        super().__init__()
        self.visual_encoder, vision_width = create_vit(vit, image_size, vit_grad_ckpt, vit_ckpt_layer, compute_dtype=compute_dtype,
                                                       cache_identical_inputs=cache_identical_inputs)
        self.tokenizer = _tokenizer_or_none(tokenizer)
        self.text_encoder = BertModel(config=_med_config(med_config, vision_width), add_pooling_layer=False,
                                      compute_dtype=compute_dtype)
        text_width = self.text_encoder.config.hidden_size
        self.vision_proj = nn.Linear(vision_width, embed_dim)
        self.text_proj = nn.Linear(text_width, embed_dim)
        self.itm_head = nn.Linear(text_width, 2)
        self.text_encoder.attach_cls_head(self.itm_head)

    @torch.no_grad()
    def forward_ids(self, image, input_ids, attention_mask, frame_of_seq=None, seqs_per_frame=0):
        """ITM logits [n_seq, 2] for tokenised captions; with frame_of_seq (or frame-major pairs and seqs_per_frame) every
        (caption, frame) pair of a video goes through one native call instead of one call per caption."""
        image_embeds = self.visual_encoder(image)                                              # blip_itm.py:43
        _, _, cls = self.text_encoder.run(input_ids, attention_mask, image_embeds, frame_of_seq=frame_of_seq, causal=False,
                                          want_hidden=False, want_cls=True, seqs_per_frame=seqs_per_frame)
        return cls

    @torch.no_grad()
    def forward(self, image, caption, match_head='itm'):
        """models/blip_itm.py:41-66 (match_head 'itm' is what run_video_CapFilt.py:111 uses)."""
        if match_head != 'itm':
            raise NotImplementedError("match_head='itc' needs the text-only encoder mode, which is not on the CapFilt path")
        if self.tokenizer is None:
            raise RuntimeError("BLIP_ITM.forward(image, caption) needs a tokenizer (bert-base-uncased is not on disk); "
                               "pass tokenizer= or call forward_ids with token ids")
        text = self.tokenizer(caption, padding='max_length', truncation=True, max_length=35, return_tensors="pt")   # :46-47
        return self.forward_ids(image, text.input_ids, text.attention_mask)


def _is_url(url_or_filename):
    from urllib.parse import urlparse
    return urlparse(url_or_filename).scheme in ("http", "https")


def load_checkpoint(model, url_or_filename):
    """models/blip.py:332-354 (local files only: there is no network on the serving box)."""
    if _is_url(url_or_filename):
        raise RuntimeError('checkpoint download is not available; pass a local path')
    if not os.path.isfile(url_or_filename):
        raise RuntimeError('checkpoint url or path is invalid')
    checkpoint = torch.load(url_or_filename, map_location='cpu')
    state_dict = checkpoint['model']
    state_dict['visual_encoder.pos_embed'] = interpolate_pos_embed(state_dict['visual_encoder.pos_embed'], model.visual_encoder)
    own = model.state_dict()
    for key in own.keys():
        if key in state_dict.keys() and state_dict[key].shape != own[key].shape:
            del state_dict[key]
    msg = model.load_state_dict(state_dict, strict=False)
    print('load checkpoint from %s' % url_or_filename)
    return model, msg


def blip_decoder(pretrained='', **kwargs):
    model = BLIP_Decoder(**kwargs)
    if pretrained:
        model, msg = load_checkpoint(model, pretrained)
        assert (len(msg.missing_keys) == 0)
    return model


def blip_itm(pretrained='', **kwargs):
    model = BLIP_ITM(**kwargs)
    if pretrained:
        model, msg = load_checkpoint(model, pretrained)
        assert (len(msg.missing_keys) == 0)
    return model
