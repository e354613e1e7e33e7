"""Timing of the fused TargetNetwork fwd / bwd at BASELINE config C4 (B=64 x 2048 pts) and C1 (B=32),
with the reference's per-sample torch loop (model/full_model.py:70-74 op sequence) on the GPU beside it."""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
hp = importlib.import_module("3d-point-clouds-autocomplete_b200")
LOC = [32, 64, 128, 64]
N = 2048
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
hp.target_network_set_mode(os.environ.get("HP_TN_MODE", "tf32x3"))
print("mode", os.environ.get("HP_TN_MODE", "tf32x3"))


def timeit(fn, reps=reps, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def torch_loop(w, x):
    """The reference's op sequence: per sample 5x (mm, + bias, relu)."""
    dims = [3] + LOC + [3]
    outs = []
    for s in range(w.size(0)):
        h, off = x[s], 0
        for l in range(5):
            i, o = dims[l], dims[l + 1]
            Wl = w[s, off:off + i * o].view(o, i)
            off += i * o
            h = torch.mm(h, Wl.t()) + w[s, off:off + o]
            off += o
            if l < 4:
                h = torch.relu(h)
        outs.append(h)
    return torch.stack(outs)


peak = max(hp._native.measure_peak(0, 8192), hp._native.measure_peak(1, 8192))
for B in (64, 32, 1):
    g = torch.Generator().manual_seed(B)
    w = (torch.randn(B, 19011, generator=g) * 0.15).cuda().requires_grad_(True)
    x = (torch.randn(B, N, 3, generator=g) * 0.6).cuda()
    go = torch.randn(B, N, 3, generator=g).cuda()
    t_f = timeit(lambda: hp.target_network_forward(w.detach(), x, LOC, True))
    y = hp.target_network_forward(w, x, LOC, True)

    def bwd():
        w.grad = None
        y.backward(go, retain_graph=True)

    t_b = timeit(bwd)
    f_flop, b_flop = 37440.0 * B * N, (74688.0 + 37440.0) * B * N
    print(f"B={B}: fwd {t_f * 1e3:.1f} us = {f_flop / t_f / 1e9:.1f} TFLOP/s ({f_flop / (t_f * 1e-3) / peak:.3f} of FP32 peak {peak / 1e12:.1f}); "
          f"bwd {t_b * 1e3:.1f} us = {b_flop / t_b / 1e9:.1f} TFLOP/s executed incl. recompute ({b_flop / (t_b * 1e-3) / peak:.3f}); "
          f"algorithmic bwd (74688 FLOP/pt) {74688.0 * B * N / (t_b * 1e-3) / peak:.3f}")
    if B != 1:
        wr = w.detach().clone().requires_grad_(True)
        t_rf = timeit(lambda: torch_loop(wr.detach(), x), reps=3, warm=1)

        def ref_fb():
            wr.grad = None
            torch_loop(wr, x).backward(go)

        t_rfb = timeit(ref_fb, reps=3, warm=1)
        print(f"      reference-style per-sample torch loop on this GPU: fwd {t_rf:.2f} ms, fwd+bwd {t_rfb:.2f} ms "
              f"(ours fwd+bwd {t_f + t_b:.3f} ms)")
