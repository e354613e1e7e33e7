"""Timing of the all-pairs CD matrix (BASELINE config C5 shapes) on one GPU."""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
hp = importlib.import_module("3d-point-clouds-autocomplete_b200")
NR = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
NS = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
P = 2048
g = torch.Generator().manual_seed(0)
ref = (torch.rand(NR, P, 3, generator=g) - 0.5).cuda()
smp = (torch.rand(NS, P, 3, generator=g) - 0.5).cuda()
peak = max(hp._native.measure_peak(0, 8192), hp._native.measure_peak(1, 8192))
for _ in range(2):
    hp.pairwise_cd(ref[:64], smp[:64])
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
cd = hp.pairwise_cd(ref, smp)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
pairs = NR * NS * P * P
print(f"pairwise_cd {NR}x{NS}x{P}^2: {ms:.2f} ms -> {pairs / ms / 1e9:.2f} T unordered pairs/s; "
      f"algorithmic 16 FLOP/unordered pair = {16 * pairs / ms / 1e9:.1f} TFLOP/s = {16 * pairs / (ms * 1e-3) / peak:.3f} of measured FP32 peak {peak / 1e12:.1f}")
# the CD half of compute_all_metrics with 1-NNA (ref-vs-sample matrix + upper triangles of the two self-distance matrices)
hp.compute_all_metrics(smp[:64], ref[:64], with_emd=False, one_nn=True)
torch.cuda.synchronize()
e0.record()
res = hp.compute_all_metrics(smp, ref, with_emd=False, one_nn=True)
vals = {k: float(v) for k, v in res.items()}
e1.record()
torch.cuda.synchronize()
print(f"compute_all_metrics CD + 1-NNA {NS} vs {NR}: {e0.elapsed_time(e1):.1f} ms", {k: round(v, 5) for k, v in vals.items()})
