"""Per-layer gradient error of the fused TargetNetwork vs the float64 oracle (debug aid)."""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
hp = importlib.import_module("3d-point-clouds-autocomplete_b200")
from oracle import oracle as O
LOC = [32, 64, 128, 64]
for (b, n, use_bias) in [(4, 2048, False), (4, 2048, True), (5, 300, False)]:
    g = torch.Generator().manual_seed(b * 100 + n)
    W = O.target_network_num_weights(LOC, use_bias)
    w = torch.randn(b, W, generator=g) * 0.15
    x = torch.randn(b, n, 3, generator=g) * 0.6
    go = torch.randn(b, n, 3, generator=g)
    wd = w.cuda().requires_grad_(True)
    y = hp.target_network_forward(wd, x.cuda(), LOC, use_bias)
    (y * go.cuda()).sum().backward()
    ogw, _ = O.target_network_backward_f64(w.numpy(), x.numpy(), go.numpy(), LOC, use_bias)
    got = wd.grad.cpu().numpy().astype(np.float64)
    dims = [3] + LOC + [3]
    off = 0
    print(f"b={b} n={n} bias={use_bias}: overall max abs err {np.abs(got-ogw).max():.3e}, ref max {np.abs(ogw).max():.3e}")
    for l in range(5):
        i, o = dims[l], dims[l + 1]
        for name, cnt in (("W", i * o), ("b", o if use_bias else 0)):
            if cnt == 0:
                continue
            e = np.abs(got[:, off:off + cnt] - ogw[:, off:off + cnt])
            s = np.unravel_index(e.argmax(), e.shape)
            print(f"   layer {l+1} {name}: max err {e.max():.3e} (ref max {np.abs(ogw[:, off:off+cnt]).max():.3e}) at sample {s[0]} elem {s[1]}; per-sample max {e.max(axis=1)}")
            off += cnt
