"""torchrun driver: compute_all_metrics sharded over the ranks must equal the single-process result bit for bit."""
import importlib
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
hp = importlib.import_module("3d-point-clouds-autocomplete_b200")
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
g = torch.Generator().manual_seed(7)
smp = (torch.rand(150, 1024, 3, generator=g) - 0.5).to(dev)
ref = (torch.rand(131, 1024, 3, generator=g) - 0.5).to(dev)
sharded = hp.compute_all_metrics(smp, ref, with_emd=True, one_nn=True)
# single-process result on every rank: a group of one
solo_group = None
for r in range(world):
    grp = dist.new_group([r])
    if r == rank:
        solo_group = grp
solo = hp.compute_all_metrics(smp, ref, with_emd=True, one_nn=True, group=solo_group)
bad = [k for k in solo if not torch.equal(solo[k].cpu(), sharded[k].cpu())]
ok = torch.tensor([0 if bad else 1], device=dev)
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
if rank == 0:
    print("keys:", len(solo), "mismatching:", bad, "all ranks ok:", bool(ok.item()))
    print({k: round(float(v), 6) for k, v in sharded.items()})
dist.destroy_process_group()
sys.exit(0 if ok.item() else 1)
