"""One ITM pass (multimodal encoder + itm_head) over synthetic (caption, frame) pairs, for ncu launch lists and event timing.

    python tools/itm_profile.py [--frames 1024] [--tokens 197] [--width 1024] [--caption-len 20] [--reps 3]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vidil_b200.med import BertConfig, BertModel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=1024)
    ap.add_argument("--tokens", type=int, default=197)
    ap.add_argument("--width", type=int, default=1024)
    ap.add_argument("--caption-len", type=int, default=20)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    m = BertModel(BertConfig(encoder_width=a.width), compute_dtype="bf16")
    head = torch.nn.Linear(768, 2)
    with torch.no_grad():
        for n, p in m.named_parameters():
            p.normal_(0.0, 0.02)
            if "LayerNorm" in n and n.endswith("weight"):
                p.add_(1.0)
    m, head = m.to(dev).eval(), head.to(dev)
    m.attach_cls_head(head)
    Fv = 8
    enc = torch.randn(a.frames, a.tokens, a.width, device=dev)
    n_pairs = a.frames * Fv                                  # every frame's caption against the 8 frames of its video
    ids = torch.randint(1000, 30000, (n_pairs, 35), device=dev)
    mask = (torch.arange(35, device=dev)[None] < a.caption_len).int().repeat(n_pairs, 1)
    # frame-major pairs: each frame against the 8 captions of its video (seqs_per_frame = 8)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for i in range(a.reps):
        ev[0].record()
        _, _, cls = m.run(ids, mask, enc, want_hidden=False, want_cls=True, seqs_per_frame=Fv)
        ev[1].record()
        torch.cuda.synchronize()
        print(f"rep {i}: ITM {ev[0].elapsed_time(ev[1]):.2f} ms for {n_pairs} pairs x {a.caption_len} tokens")


if __name__ == "__main__":
    main()
