"""CPU: the product path may not import, call, link or execute anything under oracle/, nor
fall back to a CPU implementation."""
import os
import re

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(REPO, "3d-point-clouds-autocomplete_b200")


def _sources():
    for root, _dirs, files in os.walk(PKG):
        if os.path.basename(root) in ("build", "lib", "__pycache__"):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                yield os.path.join(root, f)


def test_package_never_references_the_oracle():
    bad = []
    for p in _sources():
        text = open(p).read()
        if re.search(r"\boracle\b", text) and "hp_oracle" in text or re.search(r"(from|import)\s+oracle", text):
            bad.append(p)
        if "libhp_oracle" in text or "oracle/_ref" in text or "StructuralLossesBackend.cpython" in text:
            bad.append(p)
    assert not bad, bad


def test_missing_library_is_loud(hp, monkeypatch):
    import pytest

    monkeypatch.setattr(hp._native, "_lib", None)
    monkeypatch.setattr(hp._native, "LIB_PATH", "/nonexistent/libhp_b200.so")
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        hp._native.load()


def test_cpu_tensors_are_rejected(hp):
    import pytest
    import torch

    a = torch.zeros(1, 4, 3)
    with pytest.raises(RuntimeError, match="CUDA"):
        hp.NNDistance(a, a)
    with pytest.raises(RuntimeError, match="CUDA"):
        hp.nn_distance(a, a)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        hp.ChamferLoss()(a, a)
