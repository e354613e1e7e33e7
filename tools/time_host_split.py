"""ChamferStepGraph.run_from_host_loss_only at C2 (B=32, 2048 x 2048) for several batch splits: dependent steps (the loss is read
on the host after every step), CUDA events, L2 flushed between steps.  Also the two H2D copies alone and the device-resident step."""
import importlib
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
hp = importlib.import_module("3d-point-clouds-autocomplete_b200")
dev = torch.device("cuda:0")
B, N, M = 32, 2048, 2048
g = torch.Generator().manual_seed(0)
a_h = (torch.rand(B, N, 3, generator=g) - 0.5).pin_memory()
b_h = (torch.rand(B, M, 3, generator=g) - 0.5).pin_memory()
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=60, warm=5):
    """bench.py's pattern: everything is enqueued back to back (the CPU enqueues the timed work while the flush runs, so no launch
    gap is timed); steps are serialised by the stream."""
    for _ in range(warm):
        fn()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    torch.cuda.synchronize()
    for e0, e1 in evs:
        flush_buf.fill_(1)
        e0.record()
        fn()
        e1.record()
    torch.cuda.synchronize()
    out = [e0.elapsed_time(e1) * 1e3 for e0, e1 in evs]
    return statistics.mean(out), statistics.median(out), min(out)


da, db = torch.empty(B, N, 3, device=dev), torch.empty(B, M, 3, device=dev)


def copies():
    da.copy_(a_h, non_blocking=True)
    db.copy_(b_h, non_blocking=True)


cg, _, _ = hp.graphs._capture(copies, dev)
print("two H2D copies alone (graph): mean %.1f median %.1f min %.1f us" % timed(cg.replay))
specs = [1, 2, 3, 4, [0, 12, 32], [0, 14, 32], [0, 8, 20, 32], [0, 8, 16, 32]]
if len(sys.argv) > 1:
    specs = [eval(s) for s in sys.argv[1:]]
ref = None
for spec in specs:
    st = hp.ChamferStepGraph(B, N, M, dev, with_host_io=True, split_host_io=spec)
    st.xyz1_host.copy_(a_h)
    st.xyz2_host.copy_(b_h)
    t = timed(st.run_from_host_loss_only)
    g1, g2 = st.grad_outputs_on_device()
    if ref is None:
        ref = (g1.clone(), g2.clone(), float(st.loss_host))
        print("device-resident step (graph): mean %.1f median %.1f min %.1f us" % timed(st.replay))
    ok = torch.equal(g1, ref[0]) and torch.equal(g2, ref[1]) and abs(float(st.loss_host) - ref[2]) <= 2e-6 * abs(ref[2])
    print(f"split {spec!s:18} parts {st.host_io_parts}: mean %.1f median %.1f min %.1f us   same results: {ok}" % t, flush=True)
    del st
