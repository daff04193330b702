"""Import the UNMODIFIED reference modules: from /root/reference in the build container, or from the staged copy
under oracle/_ref/ (git-ignored, made by oracle/stage_ref.py; it travels to the GPU box like a built .so).

TEST / BASELINE INFRASTRUCTURE: used by oracle/make_golden.py to create the committed fixtures, by tests that are
skipped when no reference tree is present, and by `bench.py --impl reference` / the `cpu_baseline` leg to time
the reference's own models/vit.py on the host cores.  Never imported by anything under vidil_b200/.

`models/vit.py` imports timm and fairscale, neither of which is installed.  The stand-ins
below carry no arithmetic except PatchEmbed, which restates timm's (Conv2d(k=s=patch) -> flatten(2) ->
transpose(1,2), attributes num_patches / grid_size / proj; timm is unpinned in docker/requirements.txt:19).
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import torch
import torch.nn as nn

STAGED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def _reference_root() -> str:
    env = os.environ.get("VIDIL_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isfile("/root/reference/models/vit.py"):
        return "/root/reference"
    return STAGED_ROOT  # oracle/stage_ref.py: the reference's own files, byte for byte, outside the git history


REFERENCE_ROOT = _reference_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "vit.py"))


class _PatchEmbed(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, norm_layer=None, flatten=True):
        super().__init__()
        img_size = (img_size, img_size) if isinstance(img_size, int) else tuple(img_size)
        patch_size = (patch_size, patch_size) if isinstance(patch_size, int) else tuple(patch_size)
        self.img_size, self.patch_size = img_size, patch_size
        self.grid_size = (img_size[0] // patch_size[0], img_size[1] // patch_size[1])
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.flatten = flatten
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.norm = norm_layer(embed_dim) if norm_layer else nn.Identity()

    def forward(self, x):
        x = self.proj(x)
        if self.flatten:
            x = x.flatten(2).transpose(1, 2)
        return self.norm(x)


class _DropPath(nn.Module):
    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        return x * mask / keep


def _mod(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_shims() -> None:
    if "timm" not in sys.modules:
        _mod("timm")
        _mod("timm.models")
        _mod("timm.models.vision_transformer", _cfg=lambda **kw: kw, PatchEmbed=_PatchEmbed)
        _mod("timm.models.registry", register_model=lambda f: f)
        _mod("timm.models.layers", trunc_normal_=nn.init.trunc_normal_, DropPath=_DropPath)
        _mod("timm.models.helpers", named_apply=lambda *a, **k: None, adapt_input_conv=lambda c, w: w)
        _mod("timm.models.hub", download_cached_file=lambda *a, **k: None)
    if "fairscale" not in sys.modules:
        _mod("fairscale")
        _mod("fairscale.nn")
        _mod("fairscale.nn.checkpoint")
        _mod("fairscale.nn.checkpoint.checkpoint_activations", checkpoint_wrapper=lambda m, *a, **k: m)


def uninstall_shims() -> None:
    """Remove the stand-in packages again (transformers probes `timm` with find_spec and chokes on a bare
    module object)."""
    for name in [n for n in sys.modules if n == "timm" or n.startswith("timm.") or n == "fairscale" or
                 n.startswith("fairscale.")]:
        if getattr(sys.modules[name], "__file__", None) is None:
            del sys.modules[name]


def import_reference_vit():
    """Returns the reference's models.vit module (VisionTransformer, interpolate_pos_embed, ...)."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    install_shims()
    spec = importlib.util.spec_from_file_location("vidil_reference_vit", os.path.join(REFERENCE_ROOT, "models", "vit.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def build_reference_vit(vit: str, image_size: int, state_dict: dict | None = None):
    """The reference's create_vit (models/blip.py:298-326) restated at the call level only: it is
    `VisionTransformer(img_size, patch_size=16, embed_dim, depth, num_heads, drop_path_rate=...)` of the
    reference's own class.  ('tiny' is our CPU-sized extra.)"""
    from .weights import VIT_CONFIGS
    mod = import_reference_vit()
    D, depth, heads = VIT_CONFIGS[vit]
    m = mod.VisionTransformer(img_size=image_size, patch_size=16, embed_dim=D, depth=depth, num_heads=heads,
                              drop_path_rate=0.1 if vit == "large" else 0.0)
    if state_dict is not None:
        missing, unexpected = m.load_state_dict(state_dict, strict=True)
        assert not missing and not unexpected
    return m.eval()


def import_reference_med():
    """Returns the reference's models.med module (BertConfig, BertModel, BertLMHeadModel) under the installed
    transformers: med.py imports three helpers from `transformers.modeling_utils` that later releases moved to
    `transformers.pytorch_utils`, and calls two PreTrainedModel methods that changed; the aliases below
    carry no arithmetic."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    import transformers.modeling_utils as mu
    import transformers.pytorch_utils as pu
    for name in ("apply_chunking_to_forward", "prune_linear_layer"):
        if not hasattr(mu, name):
            setattr(mu, name, getattr(pu, name))
    if not hasattr(mu, "find_pruneable_heads_and_indices"):
        mu.find_pruneable_heads_and_indices = lambda *a, **k: (set(), None)
    if "vidil_reference_med" in sys.modules:
        return sys.modules["vidil_reference_med"]
    spec = importlib.util.spec_from_file_location("vidil_reference_med", os.path.join(REFERENCE_ROOT, "models", "med.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["vidil_reference_med"] = mod
    spec.loader.exec_module(mod)
    mod.BertPreTrainedModel.init_weights = lambda self: self.apply(self._init_weights)
    mod.BertPreTrainedModel.get_head_mask = lambda self, head_mask, n, *a: [None] * n
    return mod


def build_reference_med(name: str, kind: str, state_dict: dict):
    """kind 'decoder': the reference's BertLMHeadModel as BLIP_Decoder builds it (blip.py:95-97); kind 'itm': BertModel
    (add_pooling_layer=False) + the itm_head Linear as BLIP_ITM builds them (blip_itm.py:29-37).  Returns (model, head)."""
    from .weights import MED_CONFIGS
    med = import_reference_med()
    c = MED_CONFIGS[name]
    cfg = med.BertConfig(vocab_size=c["vocab_size"], hidden_size=c["hidden_size"], num_hidden_layers=c["num_hidden_layers"],
                         num_attention_heads=c["num_attention_heads"], intermediate_size=c["intermediate_size"],
                         max_position_embeddings=c["max_position_embeddings"], layer_norm_eps=c["layer_norm_eps"],
                         hidden_act="gelu", pad_token_id=0, type_vocab_size=2)
    cfg.encoder_width = c["encoder_width"]
    cfg.add_cross_attention = True
    if kind == "decoder":
        m = med.BertLMHeadModel(config=cfg)
        sd = {k[len("text_decoder."):]: v for k, v in state_dict.items()}
        sd["cls.predictions.decoder.bias"] = sd["cls.predictions.bias"]
        missing, unexpected = m.load_state_dict(sd, strict=False)
        assert not unexpected and all("position_ids" in k for k in missing), (missing, unexpected)
        # the output matrix is a separate tensor in the synthetic dict (BLIP checkpoints store both keys)
        assert not torch.equal(m.cls.predictions.decoder.weight, m.bert.embeddings.word_embeddings.weight)
        return m.eval(), None
    m = med.BertModel(config=cfg, add_pooling_layer=False)
    sd = {k[len("text_encoder."):]: v for k, v in state_dict.items() if k.startswith("text_encoder.")}
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all("position_ids" in k for k in missing), (missing, unexpected)
    head = nn.Linear(c["hidden_size"], 2)
    head.load_state_dict({"weight": state_dict["itm_head.weight"], "bias": state_dict["itm_head.bias"]})
    return m.eval(), head.eval()


def extract_reference_function(rel_path: str, first_line: int, last_line: int, name: str):
    """exec() a nested function of the reference straight from its source lines (nothing is copied into this
    repository) — used for `aggregate_frame_tokens`, which is a closure inside predict_video."""
    import textwrap
    from collections import defaultdict
    with open(os.path.join(REFERENCE_ROOT, rel_path)) as f:
        lines = f.readlines()[first_line - 1:last_line]
    ns = {"defaultdict": defaultdict}
    exec(textwrap.dedent("".join(lines)), ns)  # noqa: S102 - trusted local reference source
    return ns[name]
