"""Developer tool: per-frame timeline of the similarity selection kernel (VIDIL_SEL_TRACE, csrc/topk.cu)."""
import os
import sys
import tempfile

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
path = os.path.join(tempfile.gettempdir(), "sel_trace.txt")
os.environ["VIDIL_SEL_TRACE"] = path
from vidil_b200 import ops  # noqa: E402

Fr, T, Dm, k = 2048, 10000, 768, 5
img = torch.nn.functional.normalize(torch.randn(Fr, Dm, generator=torch.Generator().manual_seed(0)), dim=-1).cuda()
bank = torch.nn.functional.normalize(torch.randn(T, Dm, generator=torch.Generator().manual_seed(1)), dim=-1).cuda()
for _ in range(3):
    ops.sim_topk(img, bank, k)
torch.cuda.synchronize()
t = np.loadtxt(path, dtype=np.int64)
t0 = t[:, 1].min()
start, pool, sk, end = (t[:, i] - t0 for i in (1, 2, 3, 4))
print(f"kernel span {end.max() / 1e3:.1f} us; warp start p50 {np.median(start) / 1e3:.1f} max {start.max() / 1e3:.1f} us")
print(f"per warp: pool wait p50 {np.median(pool - start) / 1e3:.2f} us, s_k p50 {np.median(sk - pool) / 1e3:.2f}, rest p50 {np.median(end - sk) / 1e3:.2f} "
      f"p99 {np.percentile(end - sk, 99) / 1e3:.2f} max {(end - sk).max() / 1e3:.2f}; lifetime p50 {np.median(end - start) / 1e3:.2f} max {(end - start).max() / 1e3:.2f}")
order = np.argsort(-end)[:12]
for f in order:
    print(f"frame {t[f, 0]:5d} sm {t[f, 5]:3d} cta {t[f, 8]:3d} start {start[f] / 1e3:6.2f} pool {pool[f] / 1e3:6.2f} s_k {sk[f] / 1e3:6.2f} end {end[f] / 1e3:6.2f} us  "
          f"batches {t[f, 6]} groups {t[f, 7]}")
for nb in sorted(set(t[:, 6])):
    sel = t[:, 6] == nb
    print(f"batches {nb:2d} (groups {sorted(set(t[sel, 7]))}): {sel.sum():5d} frames, lifetime p50 {np.median((end - start)[sel]) / 1e3:.2f} us")
