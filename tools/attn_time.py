"""Developer tool: the ViT / CLIP attention kernels alone at the bench shapes (run under ncu for per-kernel time and
instruction counts: `ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum -k regex:attention_t python tools/attn_time.py`),
plus a CUDA-event time of the whole op-level call (which includes the fp32 <-> 16-bit casts)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vidil_b200 import ops  # noqa: E402

shapes = [(256, 197, 16, "bf16"), (256, 197, 16, "fp16"), (256, 257, 16, "bf16"), (128, 577, 12, "bf16")]
if len(sys.argv) > 1:
    shapes = shapes[:int(sys.argv[1])]
for B, N, H, dt in shapes:
    qkv = torch.randn(B, N, 3 * H * 64, device="cuda")
    for _ in range(2):
        out = ops.attention(qkv, H, dtype=dt)
    torch.cuda.synchronize()
    q, k, v = qkv.view(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)[:, :8].double()
    ref = (torch.softmax(q @ k.transpose(-1, -2) * 0.125, -1) @ v).transpose(1, 2).reshape(8, N, H * 64)
    err = (out[:8].double() - ref).abs()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.attention(qkv, H, dtype=dt)
    e1.record()
    torch.cuda.synchronize()
    print(f"B={B} N={N} H={H} {dt}: op call {e0.elapsed_time(e1) / 5 * 1e3:.0f} us (with casts); max err {err.max():.3e} mean {err.mean():.3e}",
          flush=True)
