"""Stage-by-stage timing of the C4 hot path (B=64 x 2048 points) with CUDA events, warm L2."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
hp = importlib.import_module("3d-point-clouds-autocomplete_b200")
LOC = [32, 64, 128, 64]
B, N = 64, 2048
dev = "cuda:0"
g = torch.Generator().manual_seed(0)
w = (torch.randn(B, 19011, generator=g) * 0.15).to(dev)
x = (torch.rand(B, N, 3, generator=g) - 0.5).to(dev)
gt = (torch.rand(B, N, 3, generator=g) - 0.5).to(dev)
coef = torch.full((), 0.05, device=dev)


def t(fn, reps=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3, out


us, rec = t(lambda: hp.target_network_forward(w, x, LOC, True, False)); print(f"tn fwd        {us:8.1f} us")
us, f = t(lambda: hp.chamfer_forward(gt, rec, want_inverse=True)); print(f"chamfer fwd   {us:8.1f} us")
loss, d1, i1, d2, i2, inv = f
us, gg = t(lambda: hp.chamfer_backward(gt, rec, i1, i2, coef, inv)); print(f"chamfer bwd   {us:8.1f} us")
us, _ = t(lambda: hp.target_network_backward(w, x, gg[1], LOC, True, False)); print(f"tn bwd        {us:8.1f} us")
gr = torch.randn(B, N, 3, generator=g).to(dev)
us, _ = t(lambda: hp.target_network_backward(w, x, gr, LOC, True, False)); print(f"tn bwd (randn grad_out) {us:8.1f} us")
hpg = hp.HotPathStepGraph(B, N, LOC, True, dev)
us, _ = t(hpg.replay); print(f"hot path graph {us:8.1f} us")
