"""TargetNetwork forward / forward+backward per mode, timed as CUDA-graph replays (no launch overhead), B=64 x 2048."""
import importlib
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
hp = importlib.import_module("3d-point-clouds-autocomplete_b200")
LOC = [32, 64, 128, 64]
B, N = 64, 2048
g = torch.Generator().manual_seed(1)
w = (torch.randn(B, 19011, generator=g) * 0.15).cuda()
x = (torch.randn(B, N, 3, generator=g) * 0.6).cuda()
modes = sys.argv[1:] or ["tf32x3", "fp32"]
for mode in modes:
    hp.target_network_set_mode(mode)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            y = hp.target_network_forward(w, x, LOC, True)
        s.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            y = hp.target_network_forward(w, x, LOC, True)
    ts = []
    for _ in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gr.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    tng = hp.TargetNetworkStepGraph(B, N, LOC, True, "cuda:0", channels_first=True)
    tng.weights.copy_(w)
    tb = []
    for _ in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        tng.replay()
        e1.record()
        torch.cuda.synchronize()
        tb.append(e0.elapsed_time(e1) * 1e3)
    flop = (37440.0 + 74688.0) * B * N
    print(f"{mode}: forward graph replay {statistics.median(ts):.1f} us; fwd+bwd graph replay {statistics.median(tb):.1f} us = "
          f"{flop / statistics.median(tb) / 1e6:.1f} TFLOP/s algorithmic")
