"""Quick EMD timing at BASELINE config C3 (B=32, 2048x2048) + reference extension beside it."""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
hp = importlib.import_module("3d-point-clouds-autocomplete_b200")
from oracle import oracle as O  # noqa: E402  (comparator only)

B, N, M = 32, 2048, 2048
g = torch.Generator().manual_seed(0)
a = (torch.rand(B, N, 3, generator=g) - 0.5).cuda()
b = (torch.rand(B, M, 3, generator=g) - 0.5).cuda()


def timeit(fn, reps=5, warm=2):
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


peak_mufu = hp._native.measure_peak(2, 8192)
print(f"MUFU.EX2 peak measured: {peak_mufu / 1e12:.3f} T ex2/s")
t_fused = timeit(lambda: hp.emd_cost_pairs(a, b))
ex2 = 27 * B * N * M
print(f"fused emd_cost_pairs: {t_fused:.3f} ms  -> {ex2 / t_fused / 1e9:.2f} T ex2/s algorithmic ({ex2 / (t_fused * 1e-3) / peak_mufu:.3f} of MUFU peak; 36/27 executed)")
t_fast = timeit(lambda: hp.emd_cost_pairs(a, b, fast=True))
print(f"opt-in fast emd_cost_pairs (one ex2 shared by P3 and the next P1): {t_fast:.3f} ms ({ex2 / (t_fast * 1e-3) / peak_mufu:.3f} of MUFU peak)")
t_am = timeit(lambda: hp.ApproxMatch(a, b))
match, _ = hp.ApproxMatch(a, b)
t_mc = timeit(lambda: hp.MatchCost(a, b, match))
t_mg = timeit(lambda: hp.MatchCostGrad(a, b, match))
print(f"ApproxMatch {t_am:.3f} ms, MatchCost {t_mc:.3f} ms, MatchCostGrad {t_mg:.3f} ms")
big = 512
I, J = torch.meshgrid(torch.arange(B), torch.arange(B), indexing="ij")
ia = I.reshape(-1)[:big].to(torch.int32).cuda().contiguous()
ib = J.reshape(-1)[:big].to(torch.int32).cuda().contiguous()
t_big = timeit(lambda: hp.emd_cost_pairs(a, b, ia, ib), reps=2, warm=1)
print(f"fused, {big} pairs (no column split): {t_big:.3f} ms -> {27 * big * N * M / (t_big * 1e-3) / peak_mufu:.3f} of MUFU peak")
ext = O.load_reference_ext()
if ext is not None:
    t_ref = timeit(lambda: ext.MatchCost(a, b, ext.ApproxMatch(a, b)[0]), reps=2, warm=1)
    print(f"reference extension ApproxMatch+MatchCost: {t_ref:.3f} ms")
