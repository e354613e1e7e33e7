"""Completion-evaluation metrics on the B200 kernels (the callers next to the hot path).

Mirrors the reference's ``utils/evaluation``:
  * ``mmd.minimum_mathing_distance``  (utils/evaluation/mmd.py:23-47)   -- one fused all-pairs launch instead of a
    Python double loop with a ``.item()`` sync per chunk;
  * ``total_mutual_diff.process`` / ``process_one_tmd`` (total_mutual_diff.py:14-24,50-62) and the KD-tree Chamfer
    it calls (utils/evaluation/chamfer.py:8-32);
  * ``completeness.directed_hausdorff`` / ``completeness`` (completeness.py:14-38,47-50), which materialise a
    [B,3,N,M] tensor resp. build a CPU KD-tree per cloud.
All of them are nearest-neighbour reductions over point pairs, i.e. the Chamfer kernels with a different epilogue.
Inputs may be numpy arrays (like the reference scripts load them) or tensors; results are host scalars/lists.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch

from .chamfer import NNDistance
from .metrics import pairwise_cd


def _dev(t, device) -> torch.Tensor:
    if isinstance(t, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(t, dtype=np.float32))
    return t.to(device=device, dtype=torch.float32).contiguous()


def minimum_mathing_distance(sample_pcs, ref_pcs, batch_size, device=None, first_of_chunk_only: bool = True
                             ) -> Tuple[float, List[float]]:
    """MMD-CD of reconstructions vs ground truth, ``(mmd, matched_dists)`` like utils/evaluation/mmd.py:23-47.

    The reference hands ``nn_distance`` a [1,N,3] reference and a [Bc,N,3] chunk; its backend takes the batch size from
    the first argument (structural_loss.cpp:86), so only the FIRST sample of every chunk is ever compared.
    ``first_of_chunk_only=True`` (default) reproduces that result exactly; ``False`` takes the minimum over all
    samples, which is what the metric means."""
    device = torch.device("cuda" if device is None else device)
    ref, smp = _dev(ref_pcs, device), _dev(sample_pcs, device)
    if ref.shape[1:] != smp.shape[1:]:
        raise ValueError('Incompatible size of point-clouds.')
    if first_of_chunk_only:
        smp = smp[::int(batch_size)].contiguous()
    cd = pairwise_cd(ref, smp)  # [n_ref, n_smp]: mean_i min_j + mean_j min_i, utils/evaluation/mmd.py:39
    matched = cd.min(dim=1).values
    return float(matched.mean().item()), [float(v) for v in matched.cpu().tolist()]


def total_mutual_difference(gen_pcs, device=None) -> Tuple[float, List[float]]:
    """TMD of ``gen_pcs`` [S, K, N, 3] (K completions of each of S partial shapes): per shape
    ``sum_{j<k} CD(pc_j, pc_k) * 2 / (K - 1)`` (total_mutual_diff.py:55-62), CD = mean squared NN distance both ways
    (utils/evaluation/chamfer.py:24-32).  Returns ``(mean over shapes, per-shape values)``."""
    device = torch.device("cuda" if device is None else device)
    g = _dev(gen_pcs, device)
    S, K, N, _ = g.shape
    out = []
    for s in range(S):
        cd = pairwise_cd(g[s], g[s])  # [K, K], symmetric
        iu = torch.triu_indices(K, K, offset=1, device=cd.device)
        out.append(cd[iu[0], iu[1]].sum() * 2.0 / (K - 1))
    vals = torch.stack(out) if out else torch.zeros(0, device=device)
    return (float(vals.mean().item()) if S else float("nan")), [float(v) for v in vals.cpu().tolist()]


def directed_hausdorff(point_cloud1: torch.Tensor, point_cloud2: torch.Tensor, reduce_mean: bool = True):
    """completeness.py:14-38: ``max_i min_j |a_i - b_j|`` for (B,3,N) -> (B,3,M) clouds, without the [B,3,N,M] tensor."""
    a = point_cloud1.transpose(1, 2).contiguous().float()
    b = point_cloud2.transpose(1, 2).contiguous().float()
    d1, _i1, _d2, _i2 = NNDistance(a, b)
    h = d1.max(dim=1).values.sqrt()  # sqrt is monotone: max of sqrt(min d^2) = sqrt(max min d^2)
    return h.mean() if reduce_mean else h


def unidirectional_hausdorff(existing_pcs, gen_pcs, device=None) -> float:
    """UHD (completeness.py:53-88): existing [S, 3, Ne] partial inputs vs gen_pcs [S, K, 3, N] completions; mean over
    shapes of the mean over completions of the directed Hausdorff distance partial -> completion."""
    device = torch.device("cuda" if device is None else device)
    ex, g = _dev(existing_pcs, device), _dev(gen_pcs, device)
    S, K = g.shape[0], g.shape[1]
    ex_rep = ex.unsqueeze(1).expand(S, K, ex.shape[1], ex.shape[2]).reshape(S * K, ex.shape[1], ex.shape[2])
    h = directed_hausdorff(ex_rep, g.reshape(S * K, g.shape[2], g.shape[3]), reduce_mean=False)
    return float(h.view(S, K).mean(dim=1).mean().item())


def completeness(query_points, ref_points, thres: float = 0.03, device=None) -> float:
    """completeness.py:47-50: share of query points whose nearest reference point is closer than ``thres``."""
    device = torch.device("cuda" if device is None else device)
    q, r = _dev(query_points, device).unsqueeze(0), _dev(ref_points, device).unsqueeze(0)
    d1, _i1, _d2, _i2 = NNDistance(q, r)
    return float((d1.sqrt() < thres).float().mean().item())
