"""TargetNetwork tensor-core kernels against the fp32 kernels over awkward shapes (many tiles per sample, many samples per CTA,
ragged tails, no bias, channels-first); max-normalised differences, bar 1e-5.  Inputs come from the tests' generator, which replaces
points whose pre-activation sits within fp32 rounding of a ReLU kink (there the gate -- and with it the point's whole gradient -- is
ill-defined: with plain random inputs about one point in 10^5 flips and moves single gradient entries by 1e-2)."""
import importlib, os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'tests'))
from test_target_network_gpu import _inputs  # noqa: E402
hp = importlib.import_module("3d-point-clouds-autocomplete_b200")
LOC = [32, 64, 128, 64]
worst = 0.0
for (b, n, bias, cf) in [(2, 20000, True, False), (300, 64, True, True), (1, 4096, False, False), (33, 1000, True, True), (148, 128, True, False),
                         (149, 129, False, True), (7, 16, True, False), (1, 15, True, False), (64, 2048, True, True), (3, 33333, True, False)]:
    w, x, _go = _inputs(b, n, LOC, bias, seed=b * 7 + n)
    w, x = w.cuda(), x.cuda()
    g = torch.Generator().manual_seed(b * 7 + n)
    go = torch.randn((b, 3, n) if cf else (b, n, 3), generator=g).cuda()
    res = {}
    for mode in ("fp32", "tf32x3", "mma.sync"):
        hp.target_network_set_mode(mode)
        wd, xd = w.clone().requires_grad_(True), x.clone().requires_grad_(True)
        y = hp.target_network_forward(wd, xd, LOC, bias, channels_first=cf)
        (y * go).sum().backward()
        torch.cuda.synchronize()
        res[mode] = (y.detach(), wd.grad, xd.grad)
    hp.target_network_set_mode("tf32x3")
    line = f"b={b} n={n} bias={bias} cf={cf}:"
    for mode in ("tf32x3", "mma.sync"):
        errs = [float((res[mode][i] - res["fp32"][i]).abs().max() / res["fp32"][i].abs().max()) for i in range(3)]
        worst = max(worst, *errs)
        line += f"  {mode} y {errs[0]:.1e} gw {errs[1]:.1e} gx {errs[2]:.1e}"
    print(line, flush=True)
print("WORST", worst, "(bar 1e-5)")
