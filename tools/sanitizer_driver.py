"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck / initcheck)."""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
hp = importlib.import_module("3d-point-clouds-autocomplete_b200")
dev = "cuda:0"
g = torch.Generator().manual_seed(0)
for (b, n, m) in [(2, 300, 257), (1, 1100, 130)]:
    a = (torch.rand(b, n, 3, generator=g) - 0.5).to(dev).requires_grad_(True)
    c = (torch.rand(b, m, 3, generator=g) - 0.5).to(dev).requires_grad_(True)
    loss = hp.ChamferLoss()(c, a)                      # ring forward + unpack (inverse maps) + gather backward
    loss.backward()
    d1, d2 = hp.nn_distance(a, c)
    (d1.sum() + 2 * d2.sum()).backward()
    hp.chamfer_step(a.detach(), c.detach(), torch.tensor(0.5, device=dev))  # ring forward + sectioned tail (tickets, PDL): slot path
    skew = (c.detach() * 0.02).contiguous()                                  # skewed assignment: ranking / prefix path of the tail
    hp.chamfer_step(a.detach(), skew, torch.tensor(0.5, device=dev))
    e1, i1, e2, i2 = hp.NNDistance(a.detach(), c.detach())
    hp.NNDistanceGrad(a.detach(), c.detach(), i1, i2, torch.ones_like(e1), torch.ones_like(e2))  # sorting backward
w = (torch.randn(3, 19011, generator=g) * 0.15).to(dev).requires_grad_(True)
x = (torch.randn(3, 200, 3, generator=g) * 0.6).to(dev).requires_grad_(True)
y = hp.target_network_forward(w, x, [32, 64, 128, 64], True)
y.sum().backward()
w2 = (torch.randn(2, hp.target_network_num_weights([16, 8], False), generator=g) * 0.3).to(dev).requires_grad_(True)
hp.target_network_forward(w2, x[:2].detach(), [16, 8], False).sum().backward()   # generic path
p = (torch.rand(2, 256, 3, generator=g) - 0.5).to(dev)
q = (torch.rand(2, 200, 3, generator=g) - 0.5).to(dev)
match, _ = hp.ApproxMatch(p, q)
hp.MatchCost(p, q, match)
hp.MatchCostGrad(p, q, match)
hp.emd_cost_pairs(p, p.flip(0))
hp.emd_cost_pairs(p, p.flip(0), fast=True)                                      # fused P3 + P1 sweep
ia = (torch.arange(640, dtype=torch.int32) % 2).to(dev)                           # 640 pairs: the passes fill the GPU, so the auction's
hp.emd_cost_pairs(p, q, ia, (1 - ia).contiguous())                                # chain of kernels launches programmatically dependent
hp.batch_pairwise_dist(p, q)                                                     # a2: the expansion-form matrix
s1 = (torch.rand(6, 128, 3, generator=g) - 0.5).to(dev)
s2 = (torch.rand(5, 128, 3, generator=g) - 0.5).to(dev)
hp.compute_all_metrics(s1, s2, with_emd=True, one_nn=True)
torch.cuda.synchronize()
print("sanitizer driver done")
