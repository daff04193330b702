"""Drop-in for the reference's per-frame pre-processing: run_video_CapFilt.py:128-137 `process_frame` (BLIP side) and
run_visual_tokenization.py:138-140 `processor(images=frames, return_tensors="pt")` (CLIP side: transformers'
CLIPImageProcessor — shortest edge to 224 bicubic, centre crop, rescale, normalise).

The reference converts every decoded frame to a PIL image on the host, resizes it with PIL's bicubic filter, converts to
a float tensor, normalises and copies it to the GPU — one frame at a time (:161).  Here the decoded uint8 frames go to the
GPU as they are (a third to a quarter of the bytes of the float tensor) and `vidil_preprocess_frames` produces the same
numbers there, bit for bit, for a whole batch at once.  No CPU path.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib

MEAN = (0.48145466, 0.4578275, 0.40821073)   # run_video_CapFilt.py:133
STD = (0.26862954, 0.26130258, 0.27577711)


@torch.no_grad()
def process_frames(frames_u8: torch.Tensor, image_size: int, mean=MEAN, std=STD, out: torch.Tensor | None = None,
                   recipe: str = "blip") -> torch.Tensor:
    """frames_u8: uint8 [B, H, W, 3] on a CUDA device -> float32 [B, 3, S, S].  recipe "blip": identical to stacking the
    reference's process_frame over the frames (both sides resized to S); recipe "clip": identical to transformers'
    CLIPImageProcessor (shortest edge to S, centre crop S x S) — see clip_process_frames."""
    if recipe not in ("blip", "clip"):
        raise ValueError(f"unknown pre-processing recipe {recipe!r}")
    if not frames_u8.is_cuda:
        raise RuntimeError("vidil_b200: frames must be on a CUDA device (no CPU path exists)")
    if frames_u8.dtype != torch.uint8 or frames_u8.dim() != 4 or frames_u8.shape[-1] != 3:
        raise RuntimeError(f"expected uint8 frames of shape [B, H, W, 3], got {frames_u8.dtype} {tuple(frames_u8.shape)}")
    lib = _lib.load()
    x = frames_u8.contiguous()
    B, H, W, _ = x.shape
    S = int(image_size)
    if out is None:
        out = torch.empty(B, 3, S, S, dtype=torch.float32, device=x.device)
    if B == 0:
        return out
    with torch.cuda.device(x.device):
        size_fn = lib.vidil_preprocess_workspace_bytes if recipe == "blip" else lib.vidil_clip_preprocess_workspace_bytes
        run_fn = lib.vidil_preprocess_frames if recipe == "blip" else lib.vidil_clip_preprocess_frames
        need = size_fn(B, H, W, S)
        ws = torch.empty(need + 1024, dtype=torch.uint8, device=x.device)
        off = (-ws.data_ptr()) % 1024
        m = (ctypes.c_float * 3)(*[float(v) for v in mean])
        s = (ctypes.c_float * 3)(*[float(v) for v in std])
        st = run_fn(x.data_ptr(), B, H, W, S, m, s, out.data_ptr(), ws.data_ptr() + off, need,
                    torch.cuda.current_stream().cuda_stream)
        _lib.check(st, "vidil_preprocess_frames" if recipe == "blip" else "vidil_clip_preprocess_frames")
    return out


def clip_process_frames(frames_u8: torch.Tensor, size: int = 224, mean=MEAN, std=STD, out: torch.Tensor | None = None) -> torch.Tensor:
    """uint8 [B, H, W, 3] on a CUDA device -> `pixel_values` float32 [B, 3, size, size], bit-identical to
    `CLIPImageProcessor()(images=frames, return_tensors="pt")["pixel_values"]` (PIL backend) — the call at
    run_visual_tokenization.py:138-140.  OPENAI_CLIP_MEAN / STD are the same constants as the BLIP side's."""
    return process_frames(frames_u8, size, mean, std, out, recipe="clip")


class VidilCLIPProcessor:
    """`processor` drop-in for run_visual_tokenization.py (CLIPProcessor.from_pretrained(...), :348): called as
    `processor(images=frames, return_tensors="pt")` (:138) or `processor(text=..., return_tensors="pt", padding=True,
    truncation=True)` (:88).  Images — PIL images, numpy [H, W, 3] uint8 arrays or a uint8 tensor [B, H, W, 3] — are stacked
    (per frame geometry), uploaded as uint8 and pre-processed on the GPU; the returned
    `pixel_values` already live on `device` (the caller's `.to(device)` is then a no-op).  Text goes to the tokenizer the
    reference would have used, which is host-side string work outside the hot path."""

    def __init__(self, device, tokenizer=None, size: int = 224):
        self.device = torch.device(device)
        self.tokenizer = tokenizer
        self.size = size

    def __call__(self, text=None, images=None, return_tensors="pt", **kwargs):
        if text is not None:
            if self.tokenizer is None:
                raise RuntimeError("VidilCLIPProcessor was built without a tokenizer: it pre-processes images only")
            return self.tokenizer(text, return_tensors=return_tensors, **kwargs)
        if images is None:
            raise ValueError("VidilCLIPProcessor needs `images` or `text`")
        if isinstance(images, torch.Tensor):
            batch = images if images.dim() == 4 else images[None]
        else:
            if not isinstance(images, (list, tuple)):
                images = [images]
            arrs = [np.asarray(im) for im in images]                        # PIL -> [H, W, 3] uint8
            shapes = {a.shape for a in arrs}
            if len(shapes) > 1:
                # frames of several videos with different geometries in one call (predict_video batches across videos):
                # one device call per geometry, rows returned in the caller's order
                out = torch.empty(len(arrs), 3, self.size, self.size, dtype=torch.float32, device=self.device)
                for shp in shapes:
                    rows = [i for i, a in enumerate(arrs) if a.shape == shp]
                    sub = self(images=torch.from_numpy(np.ascontiguousarray(np.stack([arrs[i] for i in rows]))))["pixel_values"]
                    out[torch.tensor(rows, device=self.device)] = sub
                return {"pixel_values": out}
            batch = torch.from_numpy(np.ascontiguousarray(np.stack(arrs)))
        if batch.dtype != torch.uint8 or batch.shape[-1] != 3:
            raise RuntimeError(f"expected uint8 RGB frames [B, H, W, 3], got {batch.dtype} {tuple(batch.shape)}")
        return {"pixel_values": clip_process_frames(batch.to(self.device, non_blocking=True), self.size)}


def process_frame(frame, config, device):
    """Same arguments and result as run_video_CapFilt.py:128-137: one decoded frame (numpy or tensor, [H, W, 3] uint8)
    -> normalised float tensor [3, S, S] on `device`."""
    t = torch.from_numpy(np.ascontiguousarray(frame)) if isinstance(frame, np.ndarray) else frame
    return process_frames(t.to(device)[None], config["image_size"])[0]


@torch.no_grad()
def encode_u8_stream(visual_encoder, batches_u8, image_size: int, outs=None, recipe: str = "blip", half_tokens: bool = False):
    """Decoded frames in, tokens out, nothing else on the host: `batches_u8` is an iterable of pinned CPU uint8 tensors
    [B, H, W, 3]; each is copied to the device on a side stream, resized / normalised there (process_frames), encoded by
    `visual_encoder` (a vidil_b200 VisionTransformer) and its [B, N+1, D] tokens copied back to pinned host memory on the
    side stream (half_tokens: in the encoder's 16-bit operand type, half the bytes).  Batch k+1's upload and batch k-1's
    download overlap batch k's kernels.  Yields one CPU tensor per batch,
    in order; with `outs` (two pinned tensors) the yielded tensor is only valid until two batches later."""
    dev = next(visual_encoder.parameters()).device
    if dev.type != "cuda":
        raise RuntimeError("vidil_b200: the module must be moved to a CUDA device first")
    with torch.cuda.device(dev):
        main = torch.cuda.current_stream()
        side_in, side_out = torch.cuda.Stream(), torch.cuda.Stream()
        pending = []   # (event, host_out, keep-alive tensors)
        dev_u8 = [None, None]
        freed = [None, None]   # event: the kernels that read dev_u8[slot] have run
        k = 0
        for frames in batches_u8:
            if frames.is_cuda or frames.dtype != torch.uint8:
                raise RuntimeError("encode_u8_stream takes host uint8 tensors [B, H, W, 3]")
            slot = k & 1
            if len(pending) == 2:
                ev, o, _ = pending.pop(0)
                ev.synchronize()
                yield o
            with torch.cuda.stream(side_in):
                if freed[slot] is not None:
                    side_in.wait_event(freed[slot])
                if dev_u8[slot] is None or dev_u8[slot].shape != frames.shape:
                    dev_u8[slot] = torch.empty(frames.shape, dtype=torch.uint8, device=dev)
                dev_u8[slot].copy_(frames, non_blocking=True)
                uploaded = torch.cuda.Event()
                uploaded.record(side_in)
            main.wait_event(uploaded)
            x = process_frames(dev_u8[slot], image_size, recipe=recipe)
            freed[slot] = torch.cuda.Event()
            freed[slot].record(main)
            tokens = visual_encoder.forward_tokens16(x) if half_tokens else visual_encoder(x)
            done = torch.cuda.Event()
            done.record(main)
            host = outs[slot] if outs is not None else torch.empty(tokens.shape, dtype=tokens.dtype, pin_memory=True)
            with torch.cuda.stream(side_out):
                side_out.wait_event(done)
                host.copy_(tokens, non_blocking=True)
                copied = torch.cuda.Event()
                copied.record(side_out)
            tokens.record_stream(side_out)
            pending.append((copied, host, (frames, x)))
            k += 1
        for ev, o, _ in pending:
            ev.synchronize()
            yield o
