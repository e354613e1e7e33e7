"""Tiny driver for ncu: a few Chamfer forward/backward launches at BASELINE config C2."""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
hp = importlib.import_module("3d-point-clouds-autocomplete_b200")
B, N, M = 32, 2048, 2048
g = torch.Generator().manual_seed(0)
a = (torch.rand(B, N, 3, generator=g) - 0.5).cuda()
b = (torch.rand(B, M, 3, generator=g) - 0.5).cuda()
one = torch.ones((), device="cuda")
fused = len(sys.argv) > 2 and sys.argv[2] == "step"
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    if fused:  # ring kernel + fused tail
        loss = hp.chamfer_step(a, b, one)[0]
        continue
    loss, d1, i1, d2, i2, inv = hp.chamfer_forward(a, b, want_inverse=True)
    hp.chamfer_backward(a, b, i1, i2, one, inv)
torch.cuda.synchronize()
print("done", float(loss))
