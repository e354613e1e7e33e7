"""Per-phase timeline of the TargetNetwork backward kernel (CTA 0: chain warp 0, helper warp 8), from a -DHP_TM_TRACE variant build:
    bash tools/build_variant.sh trace -DHP_TM_TRACE;  HP_B200_LIB=$PWD/3d-point-clouds-autocomplete_b200/lib/variants/libhp_b200_trace.so python tools/tn_trace.py"""
import ctypes
import importlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
hp = importlib.import_module("3d-point-clouds-autocomplete_b200")
B, N, LOC = 64, 2048, [32, 64, 128, 64]
g = torch.Generator().manual_seed(0)
w = (torch.randn(B, 19011, generator=g) * 0.15).cuda().requires_grad_(True)
x = (torch.randn(B, N, 3, generator=g) * 0.6).cuda()
go = torch.randn(B, N, 3, generator=g).cuda()
for _ in range(3):
    w.grad = None
    hp.target_network_forward(w, x, LOC, True).backward(go)
torch.cuda.synchronize()
lib = hp._native.load()
buf = np.zeros((2, 64, 16), np.uint64)
lib.hp_debug_tn_trace.argtypes = [ctypes.c_void_p]
rc = lib.hp_debug_tn_trace(buf.ctypes.data)
assert rc == 0, rc
t0 = int(buf[0, 0, 0])
names_c = ["tile top", "past TILE_EMPTY+stage", "x/dY loaded", "fwd done (A4_FULL)", "z4 ready", "A4_EMPTY passed", "dgrad4 done", "A3_EMPTY passed",
           "dgrad3 done", "A2_EMPTY passed", "dgrad2 done", "A1_EMPTY passed", "Z1 stored"]
names_h = ["wait A4_FULL", "got A4", "dW5 done", "got Z4", "wgrad4 done", "got Z3", "wgrad3 done", "got Z2", "wgrad2 done", "got Z1", "small done"]
for tile in range(7):
    c = [(int(v) - t0) / 1e3 if v else float("nan") for v in buf[0, tile, :13]]
    h = [(int(v) - t0) / 1e3 if v else float("nan") for v in buf[1, tile, :11]]
    print(f"tile {tile} chain :", "  ".join(f"{n} {v:.1f}" for n, v in zip(names_c, c)))
    print(f"tile {tile} helper:", "  ".join(f"{n} {v:.1f}" for n, v in zip(names_h, h)))

cta = np.zeros((256, 4), np.uint64)
lib.hp_debug_tn_cta.argtypes = [ctypes.c_void_p]
assert lib.hp_debug_tn_cta(cta.ctypes.data) == 0
c = cta[:148].astype(np.int64)
k0 = c[:, 0].min()
rel = (c - k0) / 1e3
print("per CTA (us from the first CTA's entry): entry min/max %.1f/%.1f; first tile %.1f/%.1f; last tile done %.1f/%.1f; exit %.1f/%.1f" % (
    rel[:, 0].min(), rel[:, 0].max(), rel[:, 1].min(), rel[:, 1].max(), rel[:, 2].min(), rel[:, 2].max(), rel[:, 3].min(), rel[:, 3].max()))
d = rel[:, 2] - rel[:, 1]
print("tile-loop duration per CTA: min %.1f median %.1f max %.1f; flush+fold: median %.1f max %.1f" % (d.min(), np.median(d), d.max(), np.median(rel[:, 3] - rel[:, 2]), (rel[:, 3] - rel[:, 2]).max()))
