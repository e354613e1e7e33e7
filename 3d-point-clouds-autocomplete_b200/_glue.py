"""Tensor-side helpers shared by the host wrappers (validation, streams, workspaces)."""
from __future__ import annotations

import contextlib
import threading
from typing import Dict, Tuple

import torch


def check_points(t: torch.Tensor, name: str, require_contiguous: bool = True) -> None:
    """The reference's CHECK_INPUT (structural_loss.cpp:7-9) raises RuntimeError for non-CUDA
    or non-contiguous tensors; dtype / trailing-dim problems surface there as obscure
    `data<float>()` errors or silent garbage.  Here all four are explicit RuntimeErrors."""
    if not isinstance(t, torch.Tensor):
        raise RuntimeError(f"{name} must be a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32, got {t.dtype}")
    if t.dim() != 3 or t.size(-1) != 3:
        raise RuntimeError(f"{name} must have shape [batch, points, 3], got {tuple(t.shape)}")
    if require_contiguous and not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")


def check_same_device(*tensors: torch.Tensor) -> None:
    dev = tensors[0].device
    for t in tensors[1:]:
        if t.device != dev:
            raise RuntimeError(f"all tensors must be on the same device ({dev} vs {t.device})")


@contextlib.contextmanager
def on_device_of(t: torch.Tensor):
    """Device guard (the reference has none, structural_loss.cpp:26-128) + current stream handle."""
    idx = t.device.index if t.device.index is not None else torch.cuda.current_device()
    if idx != torch.cuda.current_device():
        with torch.cuda.device(idx):
            yield torch.cuda.current_stream(idx).cuda_stream
    else:
        yield torch.cuda.current_stream(idx).cuda_stream


_workspaces: Dict[Tuple[int, int, str], torch.Tensor] = {}
_owners = threading.local()


@contextlib.contextmanager
def workspace_owner(store: dict):
    """While active (on this thread), ``zeroed_workspace`` hands out buffers that live in ``store`` instead of the shared
    per-stream cache.  CUDA-graph captures run under it with a dict the graph object keeps: the raw pointers baked into
    the graph stay valid for the graph's lifetime, and two graphs never share (and race on) one workspace, whatever stream
    handle PyTorch recycles for their capture."""
    stack = getattr(_owners, "stack", None)
    if stack is None:
        stack = _owners.stack = []
    stack.append(store)
    try:
        yield store
    finally:
        stack.pop()


def zeroed_workspace(device: torch.device, stream: int, nbytes: int, tag: str) -> torch.Tensor:
    """A scratch buffer that kernels keep zero-restored between calls: per (device, stream, tag) from a shared cache, or
    per (device, tag) from the active ``workspace_owner`` store."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    stack = getattr(_owners, "stack", None)
    if stack:
        store, key = stack[-1], (idx, tag)
        ws = store.get(key)
        if ws is None or ws.numel() < nbytes:
            if ws is not None:
                store.setdefault("_retired", []).append(ws)  # an earlier capture may still point at it
            ws = torch.zeros(max(int(nbytes), 4096), dtype=torch.uint8, device=device)
            store[key] = ws
        return ws
    key = (idx, int(stream), tag)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(max(int(nbytes), 4096), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def check_with_workspace(rc: int, what: str, ws: torch.Tensor) -> None:
    """_native.check for calls whose kernels leave state in a zero-restored workspace: if the call failed part-way (e.g.
    the ring kernel ran but the tail launch was refused), stale keys would poison every later call, so the workspace is
    cleared (stream-ordered) before the error is raised."""
    if rc != 0:
        try:
            ws.zero_()
        except Exception:
            pass
    from . import _native

    _native.check(rc, what)
