"""Phase timeline of the tcgen05 TargetNetwork forward kernel (CTA 0: epilogue warp 0 of slot 0, the MMA warp) from a -DHP_TM_TRACE build."""
import ctypes
import importlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
hp = importlib.import_module("3d-point-clouds-autocomplete_b200")
hp.target_network_set_mode("tf32x3")
B, N, LOC = 64, 2048, [32, 64, 128, 64]
g = torch.Generator().manual_seed(0)
w = (torch.randn(B, 19011, generator=g) * 0.15).cuda()
x = (torch.randn(B, N, 3, generator=g) * 0.6).cuda()
for _ in range(3):
    hp.target_network_forward(w, x, LOC, True)
torch.cuda.synchronize()
lib = hp._native.load()
buf = np.zeros((2, 64, 16), np.uint64)
lib.hp_debug_tn_trace.argtypes = [ctypes.c_void_p]
assert lib.hp_debug_tn_trace(buf.ctypes.data) == 0
t0 = int(buf[0, 0, 0])
ne = ["start", "A1 stored", "D2 ready", "A2 stored", "D3 ready", "A3a stored", "A3b stored", "D4 ready", "L5 done"]
for tile in range(0, 7):
    e = [(int(v) - t0) / 1e3 if v else float("nan") for v in buf[0, tile, :9]]
    m = [(int(v) - t0) / 1e3 if v else float("nan") for v in buf[1, tile, :12]]
    if tile % 2 == 0:
        print(f"tile {tile} epilogue(slot 0):", "  ".join(f"{n} {v:.2f}" for n, v in zip(ne, e)))
    print(f"tile {tile} mma warp      :", "  ".join(f"s{i // 3}{'wWi'[i % 3]} {v:.2f}" for i, v in enumerate(m)))

cta = np.zeros((256, 4), np.uint64)
lib.hp_debug_tn_cta.argtypes = [ctypes.c_void_p]
assert lib.hp_debug_tn_cta(cta.ctypes.data) == 0
c = cta[:148].astype(np.int64)
rel = (c - c[:, 0].min()) / 1e3
print("per CTA (us): entry %.1f..%.1f; first segment staged %.1f..%.1f (median %.1f); tiles done (thread 0) %.1f..%.1f (median %.1f); exit %.1f..%.1f" % (
    rel[:, 0].min(), rel[:, 0].max(), rel[:, 1].min(), rel[:, 1].max(), np.median(rel[:, 1]), rel[:, 2].min(), rel[:, 2].max(), np.median(rel[:, 2]),
    rel[:, 3].min(), rel[:, 3].max()))
