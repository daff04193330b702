"""One beam-search captioning pass over synthetic image tokens (no ViT), for ncu launch lists and event timing.

    python tools/med_profile.py [--frames 1024] [--tokens 197] [--width 1024] [--reps 3]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vidil_b200.med import BertConfig, BertLMHeadModel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=1024)
    ap.add_argument("--tokens", type=int, default=197)
    ap.add_argument("--width", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--max-length", type=int, default=20)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    m = BertLMHeadModel(BertConfig(encoder_width=a.width), compute_dtype="bf16")
    with torch.no_grad():
        for n, p in m.named_parameters():
            p.normal_(0.0, 0.02)
            if "LayerNorm" in n and n.endswith("weight"):
                p.add_(1.0)
        m.cls.predictions.decoder.weight.normal_(0.0, 0.1)
    m = m.to(dev).eval()
    enc = torch.randn(a.frames, a.tokens, a.width, device=dev)
    prompt = torch.tensor([[30522, 1037, 3861, 1997]]).repeat(a.frames, 1)
    kw = dict(input_ids=prompt, max_length=a.max_length, min_length=5, num_beams=3, eos_token_id=102, pad_token_id=0,
              encoder_hidden_states=enc, return_scores=True)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for i in range(a.reps):
        ev[0].record()
        out, sc, lens = m.generate(**kw)
        ev[1].record()
        torch.cuda.synchronize()
        print(f"rep {i}: generate {ev[0].elapsed_time(ev[1]):.2f} ms for {a.frames} frames, mean caption length {lens.float().mean():.1f}")


if __name__ == "__main__":
    main()
