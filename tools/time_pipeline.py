"""Throughput of ChamferHostPipeline (pinned host clouds -> H2D -> fwd+bwd -> D2H) at B=32, 2048x2048."""
import importlib, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
hp = importlib.import_module("3d-point-clouds-autocomplete_b200")
dev = torch.device("cuda", 0)
B, N, M = 32, 2048, 2048
g = torch.Generator().manual_seed(0)
a = (torch.rand(B, N, 3, generator=g) - 0.5).pin_memory()
b = (torch.rand(B, M, 3, generator=g) - 0.5).pin_memory()
for depth in (4, 5, 6, 8):
    pipe = hp.ChamferHostPipeline(B, N, M, dev, depth=depth)
    for _ in range(10):
        pipe.submit(a, b)
    pipe.drain()
    n = 400
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    pipe.fork_from()
    for _ in range(n):
        last = pipe.submit(a, b)
    t_submit = time.perf_counter() - t0
    pipe.join_into()
    e1.record()
    loss, ga, gb = pipe.result(last)
    pipe.drain()
    ref = hp.chamfer_step(a.to(dev), b.to(dev), torch.ones((), device=dev))
    ok = torch.equal(ga, ref[5].cpu()) and torch.equal(loss, ref[0].cpu())
    print(f"depth {depth}: {e0.elapsed_time(e1) / n * 1e3:.1f} us/step on the GPU timeline, host submit {t_submit / n * 1e6:.1f} us/step, results ok: {ok}")
