import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
hp = importlib.import_module("3d-point-clouds-autocomplete_b200")
from oracle import oracle as O
ext = O.load_reference_ext()
for (b, n, m) in [(2, 2048, 2048), (4, 1024, 1024), (32, 2048, 2048)]:
    g = torch.Generator().manual_seed(n + m)
    a = (torch.rand(b, n, 3, generator=g) - 0.5).cuda(); c = (torch.rand(b, m, 3, generator=g) - 0.5).cuda()
    rmatch, _ = ext.ApproxMatch(a, c); rcost = ext.MatchCost(a, c, rmatch); rg1, rg2 = ext.MatchCostGrad(a, c, rmatch)
    match, _ = hp.ApproxMatch(a, c); cost = hp.MatchCost(a, c, match); g1, g2 = hp.MatchCostGrad(a, c, match)
    torch.cuda.synchronize()
    d = (match - rmatch).abs()
    rel = d / rmatch.abs().clamp_min(1e-30)
    bad = (d > 2e-6 + 5e-4 * rmatch.abs())
    print(f"b={b} n={n} m={m}: match max abs {d.max().item():.3e} (max ref {rmatch.max().item():.3e}); violating {int(bad.sum())} of {match.numel()}; "
          f"cost rel {((cost-rcost).abs()/rcost.abs()).max().item():.2e}; g1 rel {((g1-rg1).abs().max()/rg1.abs().max()).item():.2e} g2 rel {((g2-rg2).abs().max()/rg2.abs().max()).item():.2e}")
