"""Approximate earth mover's distance (soft auction) on the B200 kernels.

Mirrors, for this path, the reference's operator interface:
  * backend functions ``ApproxMatch`` / ``MatchCost`` / ``MatchCostGrad``
    (utils/pytorch_structural_losses/structural_loss.cpp:26-73),
  * the autograd op ``match_cost`` (utils/pytorch_structural_losses/match_cost.py:5-48),
  * ``approx_match`` (named by the north star; the reference only has the backend symbol).
"""
from __future__ import annotations

from typing import Optional

import torch
from torch.autograd import Function

from . import _native
from ._glue import check_points, check_same_device, on_device_of


def _shapes(set_d, set_q, who):
    check_points(set_d, "set_d")
    check_points(set_q, "set_q")
    check_same_device(set_d, set_q)
    b = set_d.size(0)  # batch of the first argument only, like structural_loss.cpp:28,45,60
    if set_q.size(0) < b:
        raise RuntimeError(f"{who}: batch mismatch ({b} vs {set_q.size(0)})")
    if b > 0 and (set_d.size(1) == 0 or set_q.size(1) == 0):
        raise RuntimeError(f"{who}: empty point set (the reference divides n/m)")
    return b, set_d.size(1), set_q.size(1)


def ApproxMatch(set_d: torch.Tensor, set_q: torch.Tensor):
    """-> [match [B, M, N], temp [B, 2(N+M)]]  (structural_loss.cpp:26-41)."""
    b, n, m = _shapes(set_d, set_q, "ApproxMatch")
    dev = set_d.device
    match = torch.empty((b, m, n), dtype=torch.float32, device=dev)
    temp = torch.empty((b, (n + m) * 2), dtype=torch.float32, device=dev)
    lib = _native.load()
    with on_device_of(set_d) as stream:
        nbytes = lib.hp_approxmatch_workspace_bytes(b, n, m)
        ws = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=dev)  # per call: freed (stream-ordered) on return
        rc = lib.hp_approxmatch_ws(b, n, m, set_d.data_ptr(), set_q.data_ptr(), match.data_ptr(), temp.data_ptr(),
                                   ws.data_ptr(), ws.numel(), stream)
    _native.check(rc, "hp_approxmatch_ws")
    return [match, temp]


def _check_match(match, b, m, n):
    if (not match.is_cuda) or match.dtype != torch.float32 or tuple(match.shape) != (b, m, n) or not match.is_contiguous():
        raise RuntimeError(f"match must be a contiguous CUDA float32 tensor of shape {(b, m, n)}, got "
                           f"{match.dtype} {tuple(match.shape)}")


def MatchCost(set_d, set_q, match):
    """-> cost [B]  (structural_loss.cpp:43-56)."""
    b, n, m = _shapes(set_d, set_q, "MatchCost")
    _check_match(match, b, m, n)
    out = torch.empty((b,), dtype=torch.float32, device=set_d.device)
    with on_device_of(set_d) as stream:
        rc = _native.load().hp_matchcost(b, n, m, set_d.data_ptr(), set_q.data_ptr(), match.data_ptr(),
                                         out.data_ptr(), stream)
    _native.check(rc, "hp_matchcost")
    return out


def MatchCostGrad(set_d, set_q, match):
    """-> [grad1 [B,N,3], grad2 [B,M,3]]  (structural_loss.cpp:58-73)."""
    b, n, m = _shapes(set_d, set_q, "MatchCostGrad")
    _check_match(match, b, m, n)
    dev = set_d.device
    grad1 = torch.empty((b, n, 3), dtype=torch.float32, device=dev)
    grad2 = torch.empty((b, m, 3), dtype=torch.float32, device=dev)
    with on_device_of(set_d) as stream:
        rc = _native.load().hp_matchcostgrad(b, n, m, set_d.data_ptr(), set_q.data_ptr(), match.data_ptr(),
                                             grad1.data_ptr(), grad2.data_ptr(), stream)
    _native.check(rc, "hp_matchcostgrad")
    return [grad1, grad2]


def approx_match(seta: torch.Tensor, setb: torch.Tensor) -> torch.Tensor:
    """The soft assignment matrix match[b, l, k] (l over setb, k over seta); no gradient."""
    with torch.no_grad():
        return ApproxMatch(seta.contiguous(), setb.contiguous())[0]


_emd_ws = {}


def emd_cost_pairs(first: torch.Tensor, second: torch.Tensor, ia: Optional[torch.Tensor] = None,
                   ib: Optional[torch.Tensor] = None, fast: bool = False) -> torch.Tensor:
    """Match-free fused EMD cost: cost[p] = match_cost(first[ia[p]], second[ib[p]]) (forward only).

    No [pairs, M, N] matrix is ever written: the auction's per-level weights are folded into the cost
    inside pass 3 (utils/metrics.py only ever uses match_cost under no_grad).  ``fast=True`` takes hp_emd_cost_pairs_fast
    (one ex2 shared between two passes, but up to 2.1e-5 off the reference's cost -- outside the 1e-5 bar -- and, since the exact
    path leaves exhausted points out, no longer faster either: a measured variant, not a recommendation)."""
    check_points(first, "first")
    check_points(second, "second")
    check_same_device(first, second)
    n, m = first.size(1), second.size(1)
    dev = first.device
    if ia is None and ib is None:
        if first.size(0) != second.size(0):
            raise RuntimeError("emd_cost_pairs: identity pairing needs equal cloud counts")
        pairs = first.size(0)
    else:
        if ia is None or ib is None or ia.numel() != ib.numel():
            raise RuntimeError("emd_cost_pairs: ia and ib must both be given, same length")
        for t in (ia, ib):
            if t.dtype != torch.int32 or not t.is_cuda or not t.is_contiguous():
                raise RuntimeError("emd_cost_pairs: index lists must be contiguous CUDA int32 tensors")
        pairs = ia.numel()
    cost = torch.empty((pairs,), dtype=torch.float32, device=dev)
    if pairs == 0:
        return cost
    if n == 0 or m == 0:
        raise RuntimeError("emd_cost_pairs: empty point set")
    lib = _native.load()
    with on_device_of(first) as stream:
        nbytes = lib.hp_emd_cost_workspace_bytes(pairs, n, m)
        key = (dev.index, int(stream))
        ws = _emd_ws.get(key)
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
            _emd_ws[key] = ws
        fn = lib.hp_emd_cost_pairs_fast if fast else lib.hp_emd_cost_pairs
        rc = fn(pairs, n, m, first.data_ptr(), ia.data_ptr() if ia is not None else None,
                second.data_ptr(), ib.data_ptr() if ib is not None else None, cost.data_ptr(),
                ws.data_ptr(), ws.numel(), stream)
    _native.check(rc, "hp_emd_cost_pairs")
    return cost


class MatchCostFunction(Function):
    """match_cost(seta, setb) -> cost [B]  (match_cost.py:9-27).

    When no input needs a gradient the fused match-free kernel is used (that is the only way the
    reference repository calls it, utils/metrics.py:74 under no_grad).  Otherwise the match matrix is
    built, kept on ctx like match_cost.py:20, and the backward is MatchCostGrad scaled by the upstream
    gradient (match_cost.py:44-46); no gradient flows through the assignment itself."""

    @staticmethod
    def forward(ctx, seta, setb):
        needs_grad = any(ctx.needs_input_grad)
        if not needs_grad and seta.size(0) == setb.size(0):
            check_points(seta, "seta")
            check_points(setb, "setb")
            return emd_cost_pairs(seta, setb)
        match, _temp = ApproxMatch(seta, setb)
        ctx.save_for_backward(seta, setb)
        ctx.match = match
        return MatchCost(seta, setb, match)

    @staticmethod
    def backward(ctx, grad_output):
        seta, setb = ctx.saved_tensors
        grada, gradb = MatchCostGrad(seta, setb, ctx.match)
        g = grad_output.reshape(-1, 1, 1)
        grada, gradb = grada * g, gradb * g
        if setb.size(0) != seta.size(0):
            full = torch.zeros_like(setb)
            full[:seta.size(0)] = gradb
            gradb = full
        return grada, gradb


match_cost = MatchCostFunction.apply
