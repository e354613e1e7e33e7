"""ChamferStepGraph.run_from_host (H2D + step + D2H serialised in one graph) vs the device-resident step, CUDA events."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
hp = importlib.import_module("3d-point-clouds-autocomplete_b200")
dev = torch.device("cuda", 0)
step = hp.ChamferStepGraph(32, 2048, 2048, dev, with_host_io=True)


def t(fn, reps=50):
    for _ in range(5):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


print(f"HP_NO_PDL={os.environ.get('HP_NO_PDL', '0')}: replay {t(step.replay):.1f} us   run_from_host {t(step.run_from_host):.1f} us")
