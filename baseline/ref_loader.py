"""Access to the UNMODIFIED reference files staged under baseline/_ref (tools/stage_reference.py copies them byte for byte from
/root/reference in the build container; the directory is git-ignored and travels to the GPU box with the gpurun snapshot).

Measurement and test infrastructure only: bench.py's reference arm / comparators and tests/ use it; nothing under the product
package imports it.  Modules are loaded BY FILE PATH under private names, so the reference's generic top-level package names
(`model`, `utils`, `losses`) never shadow anything; only `reference_tree()` puts the staged tree on sys.path, behind the
product's drop-in directory, which is exactly how INTEGRATION.md tells a user to switch the reference over.
"""
from __future__ import annotations

import importlib
import importlib.util
import json
import os
import sys
from types import SimpleNamespace
from typing import Optional

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(REPO, "baseline", "_ref")
DROPIN = os.path.join(REPO, "3d-point-clouds-autocomplete_b200", "dropin")


def ref_root() -> Optional[str]:
    return REF if os.path.isfile(os.path.join(REF, "MANIFEST.json")) else None


def load_file(rel: str, name: str):
    """Load baseline/_ref/<rel> as module `name` (no sys.path change; the file must not import its siblings)."""
    root = ref_root()
    if root is None:
        raise FileNotFoundError("baseline/_ref is not staged: run `python tools/stage_reference.py` where /root/reference exists")
    spec = importlib.util.spec_from_file_location(name, os.path.join(root, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def reference_chamfer_loss():
    """The reference's own pure-torch ChamferLoss class (losses/champfer_loss.py:5-35)."""
    return load_file("losses/champfer_loss.py", "_hp_ref_champfer_loss").ChamferLoss


def full_model_config(sample: str = "config_completion.json.sample") -> dict:
    return json.load(open(os.path.join(REF, "settings", sample)))["full_model"]


def reference_tree() -> SimpleNamespace:
    """sys.path = [drop-in, baseline/_ref, ...]: `model.full_model.FullModel` is then the product's drop-in (a subclass of the
    reference's class with the per-sample TargetNetwork loop batched), everything else of `model.*` / `utils.points` is the
    reference's own file.  Returns both FullModel classes and the reference's TargetNetwork and ChamferLoss."""
    if ref_root() is None:
        raise FileNotFoundError("baseline/_ref is not staged")
    for p in (REF, DROPIN):
        if p in sys.path:
            sys.path.remove(p)
    sys.path[:0] = [DROPIN, REF]
    ours = importlib.import_module("model.full_model")
    ref_fm = sys.modules["model._reference_full_model"]
    ref_tn = load_file("model/target_network.py", "_hp_ref_target_network")
    ref_fm.TargetNetwork = ref_tn.TargetNetwork  # the reference class must drive the reference's own TargetNetwork, not the drop-in
    return SimpleNamespace(OurFullModel=ours.FullModel, RefFullModel=ref_fm.FullModel, RefTargetNetwork=ref_tn.TargetNetwork,
                           RefChamferLoss=reference_chamfer_loss())


def weights_init(m) -> None:
    """The initialisation the reference applies to every module before training (core/setup.py:63-77: Xavier-uniform with the
    ReLU gain on Conv / Linear weights, zero biases; BatchNorm does not occur in FullModel).  core/setup.py itself cannot be
    imported on the GPU box (it pulls utils/util.py -> matplotlib), hence this restatement for the benchmark's random init."""
    import torch

    name = m.__class__.__name__
    if name.find("Conv") != -1 or name.find("Linear") != -1:
        torch.nn.init.xavier_uniform_(m.weight, torch.nn.init.calculate_gain("relu"))
        if m.bias is not None:
            torch.nn.init.constant_(m.bias, 0)
