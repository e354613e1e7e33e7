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


# --------------------------------------------------------------------------------------
# evaluate_generativity (core/experiments.py:63-104): the generation loop, batched
# --------------------------------------------------------------------------------------
def generate_completions(full_model, existing: torch.Tensor, n_samples: int, epoch: int, device, mean: float = 0.0,
                         std: float = 0.005, n_points: int = 2048, keep_lowest: int = 1024, max_batch: int = 256,
                         return_uncut: bool = False):
    """The inner loop of ``evaluate_generativity`` (core/experiments.py:79-91) for ONE partial shape ``existing`` [1, N, 3] (or
    already transposed [1, 3, N]): ``n_samples`` completions from fresh noise, each cut to the ``keep_lowest`` points with the
    smallest second coordinate (``pc.T[pc[1].argsort()[:1024]]``, :87) -> ``[n_samples, keep_lowest, 3]`` on ``device``.

    The reference runs ``n_samples`` batch-1 forwards: per sample one CPU ``normal_`` draw of the noise, the encoder of the SAME
    ``existing`` again, the hypernetwork, ~15 TargetNetwork launches, one H2D copy, one D2H copy and a numpy argsort.  Here the
    host draws happen in the reference's order (noise_j, then the TargetNetwork input cloud of sample j: identical global-RNG
    consumption), the encoder runs once, and hypernetwork + fused TargetNetwork + the cut run on the whole batch.
    ``full_model`` is the reference's (or the drop-in) FullModel in eval mode; needs a generative mode (HyperPocket).
    ``return_uncut=True`` additionally returns the full reconstructions [n_samples, n_points, 3] (before the cut)."""
    from .target_network import generate_points, target_network_forward

    device = torch.device(device)
    if existing.size(-1) == 3:  # model/full_model.py:56-57 (the reference transposes its argument in place)
        existing.transpose_(existing.dim() - 2, existing.dim() - 1)
    noise_size = full_model.get_noise_size()
    tcfg = full_model.target_network_config
    loc, use_bias = list(tcfg["layer_out_channels"]), bool(tcfg["use_bias"])
    noises = torch.empty(n_samples, noise_size)
    points = torch.empty(n_samples, n_points, 3)
    if device.type == "cuda":
        noises, points = noises.pin_memory(), points.pin_memory()
    for j in range(n_samples):  # host RNG in the reference's order: fixed_noise (:82), then generate_points (full_model.py:72)
        noises[j] = torch.zeros(1, noise_size).normal_(mean=mean, std=std)[0]
        points[j] = generate_points(full_model.point_generator_config, epoch, (n_points, 3))
    out, full = [], []
    with torch.no_grad():
        real_mu = full_model.real_encoder(existing.to(device))           # HyperPocket.get_latent, eval branch (:108-113)
        for j0 in range(0, n_samples, max_batch):
            nz = noises[j0:j0 + max_batch].to(device, non_blocking=True)
            latent = torch.cat([nz, real_mu.expand(nz.size(0), -1)], 1)
            weights = full_model.hyper_network(latent)
            rec = target_network_forward(weights, points[j0:j0 + max_batch].to(device, non_blocking=True), loc, use_bias, False)
            order = torch.argsort(rec[:, :, 1], dim=1, stable=True)[:, :keep_lowest]     # [b, keep]: ascending second coordinate
            out.append(torch.gather(rec, 1, order.unsqueeze(-1).expand(-1, -1, 3)))
            if return_uncut:
                full.append(rec)
    if return_uncut:
        return torch.cat(out).contiguous(), torch.cat(full).contiguous()
    return torch.cat(out).contiguous()


def evaluate_generativity_for_shape(full_model, existing: torch.Tensor, cat_gt: torch.Tensor, epoch: int, device, batch_size=None,
                                    mean: float = 0.0, std: float = 0.005, group=None, with_jsd: bool = True) -> dict:
    """One iteration of the outer loop of ``evaluate_generativity`` (core/experiments.py:76-99): ``len(cat_gt)`` completions of
    ``existing`` against the ground-truth missing parts ``cat_gt`` [G, 1024, 3] -> the dict the reference accumulates per
    category (compute_all_metrics keys as floats + 'jsd')."""
    from .chamfer import ChamferLoss
    from .metrics import compute_all_metrics, jsd_between_point_cloud_sets

    cat_gt = cat_gt.to(device).contiguous()
    recs = generate_completions(full_model, existing, cat_gt.size(0), epoch, device, mean, std, keep_lowest=cat_gt.size(1))
    res = {k: float(v.item()) for k, v in compute_all_metrics(recs, cat_gt, batch_size, ChamferLoss(), group=group).items()}
    if with_jsd:
        res["jsd"] = float(jsd_between_point_cloud_sets(recs, cat_gt))
    return res
