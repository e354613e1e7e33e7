"""Print hp_measure_peak for every kind (FFMA, FFMA2, EX2, Chamfer inner-loop microbenchmarks, mma.sync TF32)."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
hp = importlib.import_module("3d-point-clouds-autocomplete_b200")
names = {0: "FFMA chain (FLOP/s)", 1: "FFMA2 chain (FLOP/s)", 2: "MUFU.EX2 (op/s)", 3: "chamfer loop packed 2q (alg FLOP/s, 8/pair)",
         4: "chamfer loop scalar 4q", 5: "chamfer loop packed 4q", 6: "mma.sync tf32 (FLOP/s)", 7: "ring-pattern FMA-pipe only (packed op/s per lane)",
         8: "ring-pattern + min3", 9: "ring-pattern + min3 + setp/sel", 10: "kind 9, 3 warps/scheduler", 11: "kind 9, 2 warps/scheduler", 12: "kind 9, 1 warp/scheduler"}
for k in range(int(sys.argv[2]) if len(sys.argv) > 2 else 0, int(sys.argv[1]) if len(sys.argv) > 1 else 7):
    try:
        r = hp._native.measure_peak(k, 4096 if k < 3 else 64)
        print(f"kind {k}: {r:.4e}  {names.get(k, '')}")
    except Exception as e:
        print("kind", k, "failed", e)
