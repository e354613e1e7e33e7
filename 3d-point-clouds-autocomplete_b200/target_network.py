"""Per-sample TargetNetwork MLP on the fused B200 kernels.

Mirrors, for this path, the reference's interface:
  * ``TargetNetwork(config, weights).forward(x)`` (model/target_network.py:5-45): one sample, its flat
    weight vector sliced into (W, b) pairs, ``mm`` + bias + ReLU per layer;
  * the per-sample loop of ``FullModel.forward`` (model/full_model.py:67-74) as ONE batched autograd op,
    ``target_network_forward(weights[B,W], points, layer_out_channels, use_bias, channels_first)``;
  * ``generate_points`` / ``generate_points_batched`` (utils/points.py:8-36): host-side input sampling,
    kept on the torch CPU RNG in the reference's draw order.
"""
from __future__ import annotations

import ctypes
import os
from typing import Sequence

import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function

from . import _native
from ._glue import on_device_of, zeroed_workspace


def target_network_dims(layer_out_channels: Sequence[int]):
    return [3] + [int(c) for c in layer_out_channels] + [3]


def _c_dims(layer_out_channels):
    dims = target_network_dims(layer_out_channels)
    return (ctypes.c_int * len(dims))(*dims), len(dims) - 1


def target_network_set_mode(mode: str) -> None:
    """Arithmetic of the tuned 32,64,128,64 kernels, process-wide: "tf32x3" (default: error-compensated 3xTF32 on the tensor
    cores, within 3e-6 of the reference's fp32 torch.mm chain; forward on tcgen05 with the activations in tensor memory, backward on
    mma.sync), "mma.sync" (the same arithmetic with the forward on mma.sync as well) or "fp32" (FFMA chains on the CUDA cores)."""
    modes = {"tf32x3": 0, "fp32": 1, "mma.sync": 2}
    if mode not in modes:
        raise ValueError(f"mode must be one of {sorted(modes)}")
    _native.check(_native.load().hp_target_network_set_mode(modes[mode]), "hp_target_network_set_mode")


if os.environ.get("HP_B200_TN_MODE"):  # process-wide default from the environment ("tf32x3" | "fp32")
    target_network_set_mode(os.environ["HP_B200_TN_MODE"])


def target_network_num_weights(layer_out_channels: Sequence[int], use_bias: bool = True) -> int:
    """Length of one sample's flat weight vector (19011 for 32,64,128,64 with bias)."""
    cd, nl = _c_dims(layer_out_channels)
    n = _native.load().hp_target_network_num_weights(nl, cd, int(bool(use_bias)))
    if n < 0:
        raise RuntimeError(f"invalid layer_out_channels {list(layer_out_channels)}")
    return int(n)


def _check(weights: torch.Tensor, points: torch.Tensor, layer_out_channels, use_bias):
    if not (isinstance(weights, torch.Tensor) and isinstance(points, torch.Tensor)):
        raise RuntimeError("weights and points must be torch tensors")
    if not (weights.is_cuda and points.is_cuda):
        raise RuntimeError("TargetNetwork (B200) needs CUDA tensors; there is no CPU fallback")
    if weights.device != points.device:
        raise RuntimeError(f"weights and points must be on the same device ({weights.device} vs {points.device})")
    if weights.dtype != torch.float32 or points.dtype != torch.float32:
        raise RuntimeError(f"weights and points must be float32, got {weights.dtype} / {points.dtype}")
    if weights.dim() != 2:
        raise RuntimeError(f"weights must have shape [batch, num_weights], got {tuple(weights.shape)}")
    W = target_network_num_weights(layer_out_channels, use_bias)
    if weights.size(1) != W:
        # the reference asserts split_index == len(weights) (target_network.py:29)
        raise AssertionError(f"weight vector has {weights.size(1)} entries, the network needs {W}")
    if points.dim() not in (2, 3) or points.size(-1) != 3:
        raise RuntimeError(f"points must have shape [batch, n, 3] or [n, 3], got {tuple(points.shape)}")
    if points.dim() == 3 and points.size(0) != weights.size(0):
        raise RuntimeError(f"batch mismatch: {weights.size(0)} weight vectors vs {points.size(0)} point clouds")


def _forward_impl(weights, points, layer_out_channels, use_bias, channels_first):
    _check(weights, points, layer_out_channels, use_bias)
    weights, points = weights.contiguous(), points.contiguous()
    b, n = weights.size(0), points.size(-2)
    shared = points.dim() == 2
    cd, nl = _c_dims(layer_out_channels)
    out = torch.empty((b, 3, n) if channels_first else (b, n, 3), dtype=torch.float32, device=weights.device)
    with on_device_of(weights) as stream:
        rc = _native.load().hp_target_network_forward(b, n, nl, cd, int(bool(use_bias)), weights.data_ptr(),
                                                      points.data_ptr(), 0 if shared else 3 * n, out.data_ptr(),
                                                      int(bool(channels_first)), stream)
    _native.check(rc, "hp_target_network_forward")
    return out, weights, points


def target_network_backward(weights: torch.Tensor, points: torch.Tensor, grad_out: torch.Tensor,
                            layer_out_channels: Sequence[int], use_bias: bool = True, channels_first: bool = False,
                            need_grad_points: bool = False):
    """(grad_weights [B, W], grad_points or None) of sum(out * grad_out); the forward is recomputed in the kernel."""
    _check(weights, points, layer_out_channels, use_bias)
    weights, points = weights.contiguous(), points.contiguous()
    b, n = weights.size(0), points.size(-2)
    shared = points.dim() == 2
    cd, nl = _c_dims(layer_out_channels)
    dev = weights.device
    grad_out = grad_out.contiguous()
    if grad_out.dtype != torch.float32 or tuple(grad_out.shape) != ((b, 3, n) if channels_first else (b, n, 3)):
        raise RuntimeError(f"grad_out must be float32 of shape {(b, 3, n) if channels_first else (b, n, 3)}, "
                           f"got {grad_out.dtype} {tuple(grad_out.shape)}")
    gw = torch.empty_like(weights)
    pts = points
    if need_grad_points and shared:  # a shared cloud's gradient is the sum over samples: expand, then reduce
        pts = points.unsqueeze(0).expand(b, n, 3).contiguous()
    gp = torch.empty((b, n, 3), dtype=torch.float32, device=dev) if need_grad_points else None
    lib = _native.load()
    with on_device_of(weights) as stream:
        nbytes = lib.hp_target_network_backward_workspace_bytes(b, n, nl, cd, int(bool(use_bias)))
        ws = zeroed_workspace(dev, stream, nbytes, "target_network")
        rc = lib.hp_target_network_backward(b, n, nl, cd, int(bool(use_bias)), weights.data_ptr(), pts.data_ptr(),
                                            0 if (shared and not need_grad_points) else 3 * n, grad_out.data_ptr(),
                                            int(bool(channels_first)), gw.data_ptr(),
                                            gp.data_ptr() if need_grad_points else None, ws.data_ptr(), ws.numel(), stream)
    _native.check(rc, "hp_target_network_backward")
    if need_grad_points and shared:
        gp = gp.sum(dim=0)
    return gw, gp


class _TargetNetworkFunction(Function):
    @staticmethod
    def forward(ctx, weights, points, layer_out_channels, use_bias, channels_first):
        out, weights, points = _forward_impl(weights, points, layer_out_channels, use_bias, channels_first)
        ctx.save_for_backward(weights, points)
        ctx.cfg = (tuple(int(c) for c in layer_out_channels), bool(use_bias), bool(channels_first))
        return out

    @staticmethod
    def backward(ctx, grad_out):
        weights, points = ctx.saved_tensors
        loc, use_bias, channels_first = ctx.cfg
        gw, gp = target_network_backward(weights, points, grad_out, loc, use_bias, channels_first,
                                         need_grad_points=ctx.needs_input_grad[1])
        return gw, gp, None, None, None


def target_network_forward(weights: torch.Tensor, points: torch.Tensor, layer_out_channels: Sequence[int],
                           use_bias: bool = True, channels_first: bool = False) -> torch.Tensor:
    """All samples of a batch in one launch.

    weights [B, W] (row s = the hypernetwork's output for sample s, hyper_network.py:41-43),
    points [B, N, 3] or one shared cloud [N, 3]  ->  [B, N, 3], or [B, 3, N] with ``channels_first=True``
    (what ``FullModel.forward`` returns, full_model.py:68,74).  Differentiable w.r.t. weights and points.
    """
    return _TargetNetworkFunction.apply(weights, points, tuple(layer_out_channels), bool(use_bias), bool(channels_first))


class TargetNetwork(nn.Module):
    """Drop-in for model/target_network.py:5-45: ``TargetNetwork(config, weights)(x)`` with x [N, 3].

    ``config`` needs ``use_bias`` and ``layer_out_channels``; ``weights`` is one sample's flat vector.
    ``self.layers`` / ``self.output`` expose the same (weight, bias) views as the reference.  The whole
    MLP runs as one fused kernel instead of five ``torch.mm`` + bias + ReLU launches."""

    def __init__(self, config, weights):
        super().__init__()
        self.use_bias = config['use_bias']
        out_ch = list(config['layer_out_channels'])
        self.layer_out_channels = out_ch
        self._weights = weights
        dims = target_network_dims(out_ch)
        split_index = 0
        views = []
        for l in range(len(dims) - 1):
            i, o = dims[l], dims[l + 1]
            layer_data = {"weight": weights[split_index:split_index + i * o].view(o, i)}
            split_index += i * o
            if self.use_bias:
                layer_data["bias"] = weights[split_index:split_index + o]
                split_index += o
            views.append(layer_data)
        self.layers = {str(l + 1): v for l, v in enumerate(views[:-1])}
        self.output = views[-1]
        self.activation = torch.nn.ReLU()
        assert split_index == len(weights)

    def forward(self, x):
        return target_network_forward(self._weights.unsqueeze(0), x, self.layer_out_channels, self.use_bias)[0]


# --------------------------------------------------------------------------------------
# input sampling (host side; utils/points.py:8-36)
# --------------------------------------------------------------------------------------
def generate_points_from_uniform_distribution(size, low=-1, high=1):
    """Rejection sampling of size[0] points uniform in the unit ball from the global torch CPU RNG:
    draws 3*size[0] candidates in the cube per attempt and keeps the first size[0] inside the ball
    (utils/points.py:8-13) -- same RNG consumption, so seeded runs reproduce the reference's inputs."""
    while True:
        points = torch.zeros([size[0] * 3, *size[1:]]).uniform_(low, high)
        points = points[torch.norm(points, dim=1) < 1]
        if points.shape[0] >= size[0]:
            return points[:size[0]]


def generate_points(config, epoch, size, normalize_points=None):
    """utils/points.py:16-36: with 'progressive' normalisation, points closer to the origin than
    coef = linspace(0, 1, max_epoch)[epoch-1] (1 after max_epoch) are projected onto the sphere of radius coef."""
    norm_cfg = config['target_network_input']['normalization']
    if normalize_points is None:
        normalize_points = norm_cfg['enable']
    points = generate_points_from_uniform_distribution(size=size)
    if normalize_points and norm_cfg['type'] == 'progressive':
        max_epoch = norm_cfg['epoch']
        coef = np.linspace(0, 1, max_epoch)[epoch - 1] if epoch <= max_epoch else 1
        inside = np.linalg.norm(points, axis=1) < coef
        sel = points[inside]
        radius = torch.from_numpy(np.linalg.norm(sel, axis=1)).float()
        points[inside] = coef * (sel.T / radius).T
    return points


def generate_points_batched(config, epoch, batch: int, size, normalize_points=None, pin: bool = True) -> torch.Tensor:
    """The B per-sample input clouds of one FullModel.forward (full_model.py:70-74), drawn sequentially in
    the reference's order, stacked into ONE [B, N, 3] (pinned) host tensor so a single H2D copy replaces B."""
    out = torch.empty((batch, size[0], *size[1:]), dtype=torch.float32)
    if pin and torch.cuda.is_available():
        out = out.pin_memory()
    for j in range(batch):
        out[j] = generate_points(config, epoch, size, normalize_points)
    return out


def reconstruct_batch(target_network_config, point_generator_config, weights: torch.Tensor, n_points: int, epoch: int,
                      device=None, points: torch.Tensor = None) -> torch.Tensor:
    """The body of the per-sample loop of ``FullModel.forward`` (model/full_model.py:67-74) for the whole batch:

        for j, w in enumerate(target_networks_weights):
            reconstruction[j] = TargetNetwork(cfg, w)(generate_points(...).to(device)).T

    becomes: draw the B input clouds on the host in the reference's order (same global CPU RNG consumption,
    SURVEY Q7), ONE pinned H2D copy, one fused kernel that writes the trainer's [B, 3, N] layout directly.
    ``weights`` [B, W] is the hypernetwork output (gradients flow back to it); ``points`` overrides the sampling."""
    device = weights.device if device is None else torch.device(device)
    b = weights.size(0)
    if points is None:
        points = generate_points_batched(point_generator_config, epoch, b, (n_points, 3), pin=weights.is_cuda)
    points = points.to(device, non_blocking=True)
    return target_network_forward(weights, points, list(target_network_config['layer_out_channels']),
                                  bool(target_network_config['use_bias']), True)
