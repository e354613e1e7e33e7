"""How far are ApproxMatch / MatchCostGrad from the reference extension, shape by shape (tolerance question of tests/test_emd_gpu.py)."""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
hp = importlib.import_module("3d-point-clouds-autocomplete_b200")
from oracle import oracle as O
ext = O.load_reference_ext()
DEV = "cuda:0"
def rel(x, y):
    x, y = x.double().cpu().numpy(), y.double().cpu().numpy()
    return float(np.abs(x - y).max() / max(np.abs(y).max(), 1e-30))
for (b, n, m, seed) in [(4, 1024, 1024, 2048), (2, 2048, 2048, 4096), (33, 256, 256, 512), (3, 500, 250, 750), (2, 100, 333, 433), (3, 300, 300, 12), (8, 2048, 2048, 1), (16, 512, 512, 2)]:
    g = torch.Generator().manual_seed(seed)
    a = (torch.rand(b, n, 3, generator=g) - 0.5).to(DEV); c = (torch.rand(b, m, 3, generator=g) - 0.5).to(DEV)
    rm, _ = ext.ApproxMatch(a, c); rg1, rg2 = ext.MatchCostGrad(a, c, rm); rc = ext.MatchCost(a, c, rm)
    rm2, _ = ext.ApproxMatch(a, c)
    m_, _ = hp.ApproxMatch(a, c); g1, g2 = hp.MatchCostGrad(a, c, m_); cc = hp.MatchCost(a, c, m_)
    big = rm > 1e-3 * rm.max()
    print(f"b={b} n={n} m={m}: match max-normalised {rel(m_, rm):.2e}, element-wise on entries > 1e-3 max {float(((m_ - rm).abs() / rm.abs())[big].max()):.2e}; "
          f"bit-identical entries {float((m_ == rm).float().mean()):.4f}; reference run-to-run identical {bool(torch.equal(rm, rm2))}; "
          f"cost rel {float(((cc - rc).abs() / rc.abs()).max()):.2e}; grad1 {rel(g1, rg1):.2e} grad2 {rel(g2, rg2):.2e}")
