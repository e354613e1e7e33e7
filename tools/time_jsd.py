"""JSD between two C5-sized sets (1000 clouds x 2048 points, 28^3 grid clipped to the sphere) on one GPU."""
import importlib, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
hp = importlib.import_module("3d-point-clouds-autocomplete_b200")
g = torch.Generator().manual_seed(0)
a = ((torch.rand(1000, 2048, 3, generator=g) - 0.5) * 0.55).cuda()
b = ((torch.rand(1000, 2048, 3, generator=g) - 0.5) * 0.5).cuda()
hp.metrics.jsd_between_point_cloud_sets(a[:10], b[:10])
torch.cuda.synchronize()
t0 = time.perf_counter()
v = hp.metrics.jsd_between_point_cloud_sets(a, b)
torch.cuda.synchronize()
print(f"jsd_between_point_cloud_sets 1000x2048 vs 1000x2048, resolution 28: {1e3 * (time.perf_counter() - t0):.1f} ms, JSD = {v:.6f}")
