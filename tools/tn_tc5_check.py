"""tcgen05 TargetNetwork forward (mode "tf32x3") against the mma.sync path and the fp32 kernels: error and time."""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
hp = importlib.import_module("3d-point-clouds-autocomplete_b200")
LOC = [32, 64, 128, 64]


def run(mode, w, x, cf=False):
    hp.target_network_set_mode(mode)
    y = hp.target_network_forward(w, x, LOC, True, channels_first=cf)
    torch.cuda.synchronize()
    return y


for (b, n) in [(1, 128), (2, 300), (3, 2048), (64, 2048), (160, 130)]:
    g = torch.Generator().manual_seed(b * 1000 + n)
    w = (torch.randn(b, 19011, generator=g) * 0.15).cuda()
    x = (torch.randn(b, n, 3, generator=g) * 0.6).cuda()
    ref = run("fp32", w, x)
    y3 = run("mma.sync", w, x)
    y5 = run("tf32x3", w, x)
    sc = float(ref.abs().max())
    print(f"b={b} n={n}: tcgen05 vs fp32 {float((y5 - ref).abs().max()) / sc:.2e}   mma.sync vs fp32 {float((y3 - ref).abs().max()) / sc:.2e}   "
          f"nan {bool(torch.isnan(y5).any())}", flush=True)
    y5c = run("tf32x3", w, x, cf=True)
    assert torch.equal(y5c.permute(0, 2, 1), y5)
b, n = 64, 2048
g = torch.Generator().manual_seed(1)
w = (torch.randn(b, 19011, generator=g) * 0.15).cuda()
x = (torch.randn(b, n, 3, generator=g) * 0.6).cuda()
for mode in ("tf32x3", "mma.sync", "fp32"):
    hp.target_network_set_mode(mode)
    for _ in range(3):
        hp.target_network_forward(w, x, LOC, True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        hp.target_network_forward(w, x, LOC, True)
    e1.record()
    torch.cuda.synchronize()
    print(f"{mode}: fwd {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
