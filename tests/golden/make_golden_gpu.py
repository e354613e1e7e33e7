"""Generate tests/golden/gpu_reference_ext.npz by RUNNING the reference's unmodified CUDA
extension (oracle/_ref, built by oracle/build_ref.sh) on a B200:

    gpurun -- python tests/golden/make_golden_gpu.py        # writes gpurun_out/gpu_reference_ext.npz
    cp gpurun_out/gpu_reference_ext.npz tests/golden/

Seeded inputs, small shapes.  Nothing here is used by the product path.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
from oracle import oracle as O  # noqa: E402

ext = O.load_reference_ext()
assert ext is not None, "oracle/_ref is missing: run oracle/build_ref.sh in the build container first"
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1856)
out = {}


def put(prefix, **kw):
    for k, v in kw.items():
        out[f"{prefix}_{k}"] = v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)


# nearest neighbour: uniform (N != M, non-multiples of anything), lattice with ties, multi-tile (M > 512)
cases = {
    "nn_uni": ((torch.rand(3, 300, 3, generator=g) - 0.5), (torch.rand(3, 257, 3, generator=g) - 0.5)),
    "nn_lat": ((torch.randint(-8, 9, (2, 700, 3), generator=g).float() / 16),
               (torch.randint(-8, 9, (2, 1100, 3), generator=g).float() / 16)),
    "nn_one": ((torch.rand(2, 1, 3, generator=g) - 0.5), (torch.rand(2, 5, 3, generator=g) - 0.5)),
}
for name, (a, b) in cases.items():
    a, b = a.to(dev).contiguous(), b.to(dev).contiguous()
    d1, i1, d2, i2 = ext.NNDistance(a, b)
    g1 = torch.randn(d1.shape, generator=g).to(dev)
    g2 = torch.randn(d2.shape, generator=g).to(dev)
    ga, gb = ext.NNDistanceGrad(a, b, i1, i2, g1, g2)
    torch.cuda.synchronize()
    put(name, a=a, b=b, d1=d1, i1=i1, d2=d2, i2=i2, g1=g1, g2=g2, ga=ga, gb=gb)

# approximate EMD: equal sizes, n > m (integer ratio 2) and n < m
emd_cases = {
    "emd_eq": ((torch.rand(2, 160, 3, generator=g) - 0.5), (torch.rand(2, 160, 3, generator=g) - 0.5)),
    "emd_nm": ((torch.rand(2, 192, 3, generator=g) - 0.5), (torch.rand(2, 96, 3, generator=g) - 0.5)),
    "emd_mn": ((torch.rand(1, 70, 3, generator=g) - 0.5), (torch.rand(1, 150, 3, generator=g) - 0.5)),
}
for name, (a, b) in emd_cases.items():
    a, b = a.to(dev).contiguous(), b.to(dev).contiguous()
    match, temp = ext.ApproxMatch(a, b)
    cost = ext.MatchCost(a, b, match)
    g1, g2 = ext.MatchCostGrad(a, b, match)
    torch.cuda.synchronize()
    put(name, a=a, b=b, match=match, cost=cost, g1=g1, g2=g2)

os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
path = os.path.join(REPO, "gpurun_out", "gpu_reference_ext.npz")
np.savez_compressed(path, **out)
print("wrote", path, "bytes", os.path.getsize(path))
