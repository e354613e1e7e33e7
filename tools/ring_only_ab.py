"""A/B timing of the ring kernel alone (bench library switches come from the environment)."""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
hp = importlib.import_module("3d-point-clouds-autocomplete_b200")
nat = hp._native; lib = nat.load_bench()
B, N, M = 32, 2048, 2048
dev = torch.device("cuda:0"); g = torch.Generator().manual_seed(0)
a = (torch.rand(B, N, 3, generator=g) - 0.5).to(dev); b = (torch.rand(B, M, 3, generator=g) - 0.5).to(dev)
ws = torch.zeros(lib.hp_chamfer_workspace_bytes(B, N, M), dtype=torch.uint8, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def run():
    nat.check_bench(lib.hp_measure_chamfer_ring_only(B, N, a.data_ptr(), M, b.data_ptr(), ws.data_ptr(), ws.numel(), st), "ring")
for _ in range(5): run()
cold, warm = [], []
for _ in range(200):
    flush.fill_(1); e0.record(); run(); e1.record(); torch.cuda.synchronize(); cold.append(e0.elapsed_time(e1) * 1e3)
for _ in range(60):
    e0.record(); run(); e1.record(); torch.cuda.synchronize(); warm.append(e0.elapsed_time(e1) * 1e3)
env = {k: os.environ.get(k, "-") for k in ("HP_NO_CARVEOUT", "HP_RING_ONLY_TICKETS", "HP_RING_WARPS", "HP_RING_VARIANT")}
print(f"ring alone: cold mean {np.mean(cold):.2f} median {np.median(cold):.1f} min {min(cold):.1f} | warm median {np.median(warm):.1f} min {min(warm):.1f} us  {env}")
