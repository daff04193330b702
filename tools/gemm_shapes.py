"""Developer tool: run the headline GEMM shapes with each epilogue once (time them with ncu -k regex:gemm_tcgen05)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vidil_b200 import _lib, ops  # noqa: E402

M = 50432
cases = [("fc1-shape store", 4096, 1024, _lib.EPI_STORE), ("fc1-shape gelu", 4096, 1024, _lib.EPI_GELU),
         ("fc1-shape quickgelu", 4096, 1024, _lib.EPI_QUICKGELU), ("fc1-shape f32 store", 4096, 1024, _lib.EPI_STORE_F32),
         ("qkv", 3072, 1024, _lib.EPI_STORE), ("proj resid", 1024, 1024, _lib.EPI_RESID),
         ("proj-shape store", 1024, 1024, _lib.EPI_STORE), ("fc2 resid", 1024, 4096, _lib.EPI_RESID),
         ("fc2-shape store", 1024, 4096, _lib.EPI_STORE), ("fc1-shape gelu nobias", 4096, 1024, _lib.EPI_GELU)]
for name, N, K, epi in cases:
    a = torch.randn(M, K, device="cuda")
    w = torch.randn(N, K, device="cuda") * K ** -0.5
    b = None if "nobias" in name else torch.randn(N, device="cuda")
    out = torch.zeros(M, N, device="cuda") if epi == _lib.EPI_RESID else None
    ops.linear(a, w, b, epilogue=epi, out=out)
    torch.cuda.synchronize()
    print(name)
    del a, w, out
