"""Per-CTA timeline of the Chamfer step (bench library, hp_measure_set_trace): when do the tail CTAs start, how long do they
wait for their cloud's ticket, how much of the tail is exposed behind the ring kernel.  Usage: python tools/chamfer_timeline.py [B N M]"""
import ctypes
import importlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
hp = importlib.import_module("3d-point-clouds-autocomplete_b200")
nat = hp._native
lib = nat.load_bench()
B, N, M = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (32, 2048, 2048)
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
a = (torch.rand(B, N, 3, generator=g) - 0.5).to(dev)
b = (torch.rand(B, M, 3, generator=g) - 0.5).to(dev)
ws = torch.zeros(lib.hp_chamfer_workspace_bytes(B, N, M), dtype=torch.uint8, device=dev)
d1 = torch.empty(B, N, device=dev); i1 = torch.empty(B, N, dtype=torch.int32, device=dev)
d2 = torch.empty(B, M, device=dev); i2 = torch.empty(B, M, dtype=torch.int32, device=dev)
loss = torch.empty(1, device=dev); g1 = torch.empty(B, N, 3, device=dev); g2 = torch.empty(B, M, 3, device=dev)
one = torch.ones(1, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
RW = int(os.environ.get('HP_RING_WARPS', '4')) * 256  # rows per ring CTA (bench library switch)
ring_ctas = B * ((N + RW - 1) // RW) * ((M + 127) // 128)
tail_ctas = B * ((N + 255) // 256 + (M + 255) // 256)
trace = torch.zeros(2 * ring_ctas + 10 * tail_ctas, dtype=torch.int64, device=dev)


def step():
    nat.check_bench(lib.hp_chamfer_step(B, N, a.data_ptr(), M, b.data_ptr(), one.data_ptr(), d1.data_ptr(), i1.data_ptr(), d2.data_ptr(),
                                        i2.data_ptr(), loss.data_ptr(), g1.data_ptr(), g2.data_ptr(), ws.data_ptr(), ws.numel(), st), "step")


for _ in range(5):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for _ in range(200):
    flush.fill_(1)
    e0.record(); step(); e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
print(f"step (eager launches, L2 flushed): mean {np.mean(ts):.2f} median {np.median(ts):.1f} us  min {min(ts):.1f} us   HP_NO_PDL={os.environ.get('HP_NO_PDL', '0')} "
      f"HP_RING_WARPS={os.environ.get('HP_RING_WARPS', '4')} HP_TAIL_TICKETS={os.environ.get('HP_TAIL_TICKETS', 'default')} HP_RING_VARIANT={os.environ.get('HP_RING_VARIANT', '0')}")
if os.environ.get("HP_TIMELINE_BRIEF"):
    sys.exit(0)
nat.check_bench(lib.hp_measure_set_trace(trace.data_ptr()), "trace")
for rep in range(2):
    trace.zero_()
    flush.fill_(1)
    step()
    torch.cuda.synchronize()
    t = trace.cpu().numpy().astype(np.float64)
    r = t[: 2 * ring_ctas].reshape(-1, 2)
    q10 = t[2 * ring_ctas:].reshape(-1, 10)
    t0 = r[:, 0].min()
    q10 = (q10 - t0) / 1e3
    q = q10[:, [0, 1, 8]]
    r = (r - t0) / 1e3
    names = ["start", "ticket", "keys+loss sums", "ranking", "prefix+scan", "placement", "gather+stores", "service flags", "end"]
    late = q10[:, 1] > r[:, 1].max() - 1.0   # tail CTAs whose ticket came with the last ring CTAs: the exposed ones
    print("tail phases (us after the ticket; median over all CTAs | over the last-cloud CTAs): " + ", ".join(
        f"{names[i]} {np.median(q10[:, i] - q10[:, 1]):.1f}|{np.median(q10[late, i] - q10[late, 1]):.1f}" for i in (2, 3, 4, 5, 6, 7, 8)
        if np.median(q10[:, i]) > 0))
    print(f"--- rep {rep}: ring CTAs {ring_ctas}, tail CTAs {tail_ctas} (times in us from the first ring CTA's start)")
    print(f"ring: last start {r[:,0].max():.1f}  first end {r[:,1].min():.1f}  last end {r[:,1].max():.1f}  CTA duration median {np.median(r[:,1]-r[:,0]):.1f} max {(r[:,1]-r[:,0]).max():.1f}")
    print(f"tail: first start {q[:,0].min():.1f}  last start {q[:,0].max():.1f}  last end {q[:,2].max():.1f}")
    print(f"tail: wait for ticket median {np.median(q[:,1]-q[:,0]):.1f} max {(q[:,1]-q[:,0]).max():.1f}; work after ticket median {np.median(q[:,2]-q[:,1]):.1f} max {(q[:,2]-q[:,1]).max():.1f}")
    per_cloud = tail_ctas // B
    for c in (0, B // 4, B // 2, 3 * B // 4, B - 1):
        qc = q[c * per_cloud:(c + 1) * per_cloud]
        rc = r[c * (ring_ctas // B):(c + 1) * (ring_ctas // B)]
        print(f"  cloud {c:3d}: ring done {rc[:,1].max():6.1f} | tail start {qc[:,0].min():6.1f}..{qc[:,0].max():6.1f}  ticket {qc[:,1].min():6.1f}..{qc[:,1].max():6.1f}  end {qc[:,2].min():6.1f}..{qc[:,2].max():6.1f}")
    started_before = (q[:, 0] < r[:, 1].max()).sum()
    print(f"tail CTAs started before the ring kernel's last CTA ended: {started_before} / {tail_ctas}; exposed tail = {q[:,2].max() - r[:,1].max():.1f} us")
nat.check_bench(lib.hp_measure_set_trace(None), "trace off")
# the ring kernel alone: per-CTA timeline (start time vs duration), then event timing
nat.check_bench(lib.hp_measure_set_trace(trace.data_ptr()), "trace")
ring_ws0 = torch.zeros_like(ws)
trace.zero_()
flush.fill_(1)
nat.check_bench(lib.hp_measure_chamfer_ring_only(B, N, a.data_ptr(), M, b.data_ptr(), ring_ws0.data_ptr(), ring_ws0.numel(), st), "ring")
torch.cuda.synchronize()
t = trace.cpu().numpy().astype(np.float64)
r = t[: 2 * ring_ctas].reshape(-1, 2)
r = (r - r[:, 0].min()) / 1e3
order = np.argsort(r[:, 0])
dur = (r[:, 1] - r[:, 0])[order]
print(f"ring alone: last start {r[:,0].max():.1f}  last end {r[:,1].max():.1f}; CTA duration by start order: first 592 median {np.median(dur[:592]):.1f}, "
      f"next {len(dur)-592} median {np.median(dur[592:]):.1f}, last 100 median {np.median(dur[-100:]):.1f} min {dur.min():.1f} max {dur.max():.1f}")
ends = np.sort(r[:, 1])
print("ring alone: CTAs still running at t = " + ", ".join(f"{tt}us:{int((r[:,0] <= tt).sum() - (r[:,1] <= tt).sum())}" for tt in (5, 15, 25, 30, 35, 40, 45, 48)))
nat.check_bench(lib.hp_measure_set_trace(None), "trace off")
# the ring kernel alone, with and without the ticket arrival
ring_ws = torch.zeros_like(ws)
ts = []
for _ in range(30):
    flush.fill_(1)
    e0.record()
    nat.check_bench(lib.hp_measure_chamfer_ring_only(B, N, a.data_ptr(), M, b.data_ptr(), ring_ws.data_ptr(), ring_ws.numel(), st), "ring")
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
print(f"ring kernel alone: median {np.median(ts):.1f} us  min {min(ts):.1f} us  HP_RING_VARIANT={os.environ.get('HP_RING_VARIANT', '0')}")
