"""CPU restatement of the reference's per-frame pre-processing — parity oracle.

TEST INFRASTRUCTURE (see oracle/vit_oracle.py for the import rule).

Reference: run_video_CapFilt.py:128-137 `process_frame` = torchvision Compose([ToPILImage(), Resize((S, S), BICUBIC),
ToTensor(), Normalize(mean, std)]).  The arithmetic of the resize lives in Pillow (un-vendored, unpinned; 12.2.0 installed
here): src/libImaging/Resample.c — `precompute_coeffs` (double-precision bicubic weights with the support widened by the
down-scaling factor, i.e. antialiased), `normalize_coeffs_8bpc` (weights rounded to 22-bit fixed point),
`ImagingResampleHorizontal_8bpc` then `ImagingResampleVertical_8bpc` (int32 accumulation from 1 << 21, >> 22, clip to
[0, 255]; the horizontal result is rounded to uint8 before the vertical pass).  ToTensor is uint8 -> float32 / 255,
Normalize is (x - mean) / std in float32.  Pinned bit-for-bit against PIL/torchvision itself: tests/golden/preprocess.npz
(oracle/make_golden.py) and, where PIL is importable, live in tests/test_oracle_golden.py.
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2
MEAN = (0.48145466, 0.4578275, 0.40821073)   # run_video_CapFilt.py:133
STD = (0.26862954, 0.26130258, 0.27577711)


def _bicubic(x: float) -> float:
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size: int, out_size: int):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the full-image box: (bounds [out,2] int32, kk [out,ksize] int32)."""
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = sum(w)
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(0.5 + v * (1 << PRECISION_BITS)) if v >= 0 else int(-0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _resample_axis(img: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    """One 8-bit pass along `axis` of an [H, W, C] uint8 image."""
    bounds, kk = precompute_coeffs(img.shape[axis], out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], dtype=np.uint8)
    for xx in range(out_size):
        xmin, n = bounds[xx]
        acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(kk[xx, :n].astype(np.int64), src[xmin:xmin + n], axes=(0, 0))
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_bicubic_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """PIL Image.resize((out_w, out_h), BICUBIC) of an [H, W, C] uint8 array: horizontal pass, then vertical."""
    tmp = _resample_axis(img, out_w, axis=1) if img.shape[1] != out_w else img
    return _resample_axis(tmp, out_h, axis=0) if img.shape[0] != out_h else tmp


def process_frame(frame: np.ndarray, image_size: int, mean=MEAN, std=STD) -> np.ndarray:
    """run_video_CapFilt.py:128-137: [H, W, 3] uint8 -> [3, S, S] float32."""
    r = resize_bicubic_u8(np.ascontiguousarray(frame), image_size, image_size)
    x = r.astype(np.float32) / np.float32(255.0)                      # ToTensor
    x = (x - np.asarray(mean, dtype=np.float32)) / np.asarray(std, dtype=np.float32)   # Normalize, float32 sub then div
    return np.ascontiguousarray(x.transpose(2, 0, 1))


CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)   # transformers.image_utils.OPENAI_CLIP_MEAN
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)   # OPENAI_CLIP_STD


def clip_resize_geometry(h: int, w: int, size: int):
    """transformers' get_resize_output_image_size(image, size=shortest_edge, default_to_square=False): the shorter side
    becomes `size`, the longer one int(size * long / short).  Returns (new_h, new_w, top, left) with the centre-crop
    origin of image_transforms.center_crop: (new - size) // 2 on each axis."""
    short, long_ = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long_ / short)
    nh, nw = (new_long, new_short) if w <= h else (new_short, new_long)
    return nh, nw, (nh - size) // 2, (nw - size) // 2


def clip_process_frame(frame: np.ndarray, size: int = 224, mean=CLIP_MEAN, std=CLIP_STD) -> np.ndarray:
    """run_visual_tokenization.py:138-140 `processor(images=frames, return_tensors="pt")` for one frame = transformers'
    CLIPImageProcessor with the PIL backend (the only one that existed when the reference was written; un-vendored,
    unpinned — 5.5.0 installed here, class CLIPImageProcessorPil): PIL bicubic resize of the shortest edge to `size`,
    centre crop size x size, rescale = float64(u8) * (1/255) cast to float32, normalise (x - mean) / std in float32.
    [H, W, 3] uint8 -> [3, size, size] float32.  Pinned bit-for-bit against the processor itself:
    tests/golden/clip_preprocess.npz (oracle/make_golden.py) and live in tests/test_oracle_golden.py."""
    h, w = frame.shape[:2]
    nh, nw, top, left = clip_resize_geometry(h, w, size)
    r = resize_bicubic_u8(np.ascontiguousarray(frame), nh, nw)[top:top + size, left:left + size]
    x = (r.astype(np.float64) * (1 / 255)).astype(np.float32)
    x = (x - np.asarray(mean, dtype=np.float32)) / np.asarray(std, dtype=np.float32)
    return np.ascontiguousarray(x.transpose(2, 0, 1))
