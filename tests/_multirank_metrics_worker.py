"""torchrun worker of tests/test_metrics_multirank_gpu.py: compute_all_metrics sharded over the ranks must equal the
single-process result (a group of one, same kernels) BIT FOR BIT, for every result key.

    torchrun --nproc-per-node G tests/_multirank_metrics_worker.py <backend> <case>

backend nccl: one rank per GPU (NVLink collectives).  backend gloo: every rank on cuda:0 (a one-GPU box): the kernels still
run on the GPU for every shard, only the few-KB (min, argmin) vectors travel through host memory.
case "mid":  150 x 131 clouds x 1024 points, CD + EMD + 1-NNA (all 12 keys; unequal set sizes, ragged last row block)
case "c5cd": 1000 x 1000 clouds x 2048 points, CD + 1-NNA-CD (BASELINE config C5, CD half at full size)
"""
import importlib
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
hp = importlib.import_module("3d-point-clouds-autocomplete_b200")
backend, case = sys.argv[1], sys.argv[2]
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", local if backend == "nccl" else 0)
torch.cuda.set_device(dev)
if backend == "nccl":
    dist.init_process_group("nccl", device_id=dev)
else:
    dist.init_process_group("gloo")
g = torch.Generator().manual_seed(7)
if case == "mid":
    smp = (torch.rand(150, 1024, 3, generator=g) - 0.5).to(dev)
    ref = (torch.rand(131, 1024, 3, generator=g) - 0.5).to(dev)
    kw = dict(with_emd=True, one_nn=True)
else:
    smp = (torch.rand(1000, 2048, 3, generator=g) - 0.5).to(dev)
    ref = (torch.rand(1000, 2048, 3, generator=g) - 0.5).to(dev)
    kw = dict(with_emd=False, one_nn=True)
sharded = hp.compute_all_metrics(smp, ref, **kw)
solo_group = None
for r in range(world):  # every rank must take part in every new_group call
    grp = dist.new_group([r], backend=backend)
    if r == rank:
        solo_group = grp
solo = hp.compute_all_metrics(smp, ref, group=solo_group, **kw)
bad = [k for k in solo if not torch.equal(solo[k].cpu(), sharded[k].cpu())]
ok = torch.tensor([0 if bad or set(solo) != set(sharded) else 1])
if backend == "nccl":
    ok = ok.to(dev)
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
if rank == 0:
    print("RESULT " + json.dumps({"world": world, "backend": backend, "case": case, "keys": len(solo), "mismatching": bad,
                                  "all_ranks_ok": bool(ok.item()), "values": {k: float(v) for k, v in sharded.items()}}), flush=True)
dist.destroy_process_group()
sys.exit(0 if ok.item() else 1)
