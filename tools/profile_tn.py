"""Tiny driver for ncu: TargetNetwork forward + backward launches at BASELINE config C4 (B=64 x 2048 points)."""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
hp = importlib.import_module("3d-point-clouds-autocomplete_b200")
hp.target_network_set_mode(os.environ.get("HP_TN_MODE", "tf32x3"))
B, N, LOC = 64, 2048, [32, 64, 128, 64]
g = torch.Generator().manual_seed(0)
w = (torch.randn(B, 19011, generator=g) * 0.15).cuda().requires_grad_(True)
x = (torch.randn(B, N, 3, generator=g) * 0.6).cuda()
go = torch.randn(B, N, 3, generator=g).cuda()
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 5):
    w.grad = None
    hp.target_network_forward(w, x, LOC, True).backward(go)
torch.cuda.synchronize()
print("done")
