"""How far the TargetNetwork kernels are from the float64 oracle, per mode (max-abs error / max-abs value, the tests' metric;
bar: 1e-5).  usage: python tools/tn_error_margins.py"""
import importlib
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
hp = importlib.import_module("3d-point-clouds-autocomplete_b200")
from oracle import oracle as O  # noqa: E402
from test_target_network_gpu import _inputs, _rel_err, FAST  # noqa: E402


def f64_forward(w, x, loc, use_bias):
    w, h = np.asarray(w, np.float64), np.asarray(x, np.float64)
    dims = [3] + list(loc) + [3]
    off = 0
    for l in range(len(dims) - 1):
        i, o = dims[l], dims[l + 1]
        W = w[:, off:off + i * o].reshape(-1, o, i)
        off += i * o
        h = np.einsum("bni,boi->bno", h, W)
        if use_bias:
            h = h + w[:, off:off + o][:, None, :]
            off += o
        if l < len(dims) - 2:
            h = np.maximum(h, 0.0)
    return h


for mode in ("tf32x3", "mma.sync", "fp32"):
    hp.target_network_set_mode(mode)
    worst = [0.0, 0.0, 0.0]
    for (b, n, seed) in [(4, 2048, 1), (8, 2048, 2), (70, 256, 3), (160, 130, 4), (2, 4096, 5), (3, 2048, 6)]:
        w, x, go = _inputs(b, n, FAST, True, seed=seed)
        wd = w.cuda().requires_grad_(True)
        xd = x.cuda().requires_grad_(True)
        y = hp.target_network_forward(wd, xd, FAST, True)
        (y * go.cuda()).sum().backward()
        y64 = f64_forward(w.numpy(), x.numpy(), FAST, True)
        gw, gx = O.target_network_backward_f64(w.numpy(), x.numpy(), go.numpy(), FAST, True)
        e = (_rel_err(y.detach().cpu().numpy(), y64), _rel_err(wd.grad.cpu().numpy(), gw), _rel_err(xd.grad.cpu().numpy(), gx))
        worst = [max(a, c) for a, c in zip(worst, e)]
        print(f"{mode} b={b} n={n}: y {e[0]:.2e}  grad_w {e[1]:.2e}  grad_x {e[2]:.2e}")
    print(f"{mode} WORST: y {worst[0]:.2e}  grad_w {worst[1]:.2e}  grad_x {worst[2]:.2e}   (bar 1e-5)")
