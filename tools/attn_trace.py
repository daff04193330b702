"""Developer tool: timeline (SM cycles) of CTA 0 of the tcgen05 attention kernel at the headline shape."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vidil_b200 import _lib, ops  # noqa: E402

B, N, H = 256, int(sys.argv[1]) if len(sys.argv) > 1 else 197, 16
qkv = torch.randn(B, N, 3 * H * 64, device="cuda")
lib = _lib.load()
ops.attention(qkv, H)  # warm
trace = torch.zeros(7 * 16 * 8, dtype=torch.int64, device="cuda")
lib.vidil_debug_set_attention_trace(trace.data_ptr())
ops.attention(qkv, H)
torch.cuda.synchronize()
lib.vidil_debug_set_attention_trace(None)
t = trace.view(7, 16, 8).cpu()
t0 = int(t[t > 0].min())
names = {0: ["top", "qk_empty ok", "v_empty ok"],
         1: ["top", "qk_full ok", "s_free ok", "S issued", "v_full ok", "p_full ok", "PV issued"],
         3: ["top", "s_full ok", "pass1 done", "token ok", "P written", "o_full ok", "O drained"]}
names[2], names[4] = names[1], names[3]
names[5] = names[6] = [f"c{i}" for i in range(8)]   # even chunks: after the TMEM-load wait; odd chunks: before it
if N == 257:
    names[3] = names[4] = ["top", "sx done", "s_full ok", "pass1 done", "token ok", "P written", "o_full ok", "O drained"]
    names[5] = ["top", "qk_full ok", "K released", "v_full ok", "V released"]
roles = ["producer", "mma L0", "mma L1", "softmax L0", "softmax L1", "chunks L0" if N != 257 else "tail", "chunks L1"]
for it in range(4, 9):
    print(f"--- item iteration {it}")
    for r in range(7):
        row = t[r, it]
        evs = [(names[r][e], int(row[e]) - t0) for e in range(len(names[r])) if row[e] > 0]
        print(f"  {roles[r]:11s} " + "  ".join(f"{n}:{c}" for n, c in evs))
