"""Mirror of the ViT factory in the reference's `models/blip.py` (create_vit, :298-326).

The BLIP wrappers (BLIP_Decoder, BLIP_ITM, ...) stay the reference's: they obtain their image tower from
`create_vit` and call it as `self.visual_encoder(image)` (blip.py:94,106,128; blip_itm.py:27,43), so
re-pointing this one import makes every one of those call sites run on the native path.
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
