"""ncu driver: Chamfer forward/backward launches at a chosen batch (default 37: 1184 ring work items = 8 per SM)."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
hp = importlib.import_module("3d-point-clouds-autocomplete_b200")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 37
g = torch.Generator().manual_seed(0)
a = (torch.rand(B, 2048, 3, generator=g) - 0.5).cuda()
b = (torch.rand(B, 2048, 3, generator=g) - 0.5).cuda()
one = torch.ones((), device="cuda")
for _ in range(4):
    loss, d1, i1, d2, i2, inv = hp.chamfer_forward(a, b, want_inverse=True)
    hp.chamfer_backward(a, b, i1, i2, one, inv)
torch.cuda.synchronize()
