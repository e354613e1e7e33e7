#!/usr/bin/env python
"""Stage the UNMODIFIED reference files that bench.py's reference arm and the C1 / C4 comparisons import into
git-ignored ``baseline/_ref/``.  ``/root/reference`` exists only in the build container; ``baseline/_ref``
travels to the GPU box with the gpurun snapshot (it is git-ignored, not gpurun-ignored).  Nothing is copied into tracked
paths, nothing is edited: the files are byte-identical copies and ``MANIFEST.json`` records their sha256.

    python tools/stage_reference.py            # copy (idempotent)
    python tools/stage_reference.py --check    # exit 1 if a staged file differs from its source

``pip install`` of the reference is not applicable: it has no setup.py / pyproject.toml at its root (it is a script tree run
with PYTHONPATH=.), so the "install" is this copy; its CUDA extension is built separately into oracle/_ref by
oracle/build_ref.sh.  Only the pure-torch host files are staged -- the ones that import on a box without h5py / trimesh /
matplotlib / ray: model/*, losses/*, utils/points.py, utils/__init__.py, utils/metrics.py, the
structural-loss Python wrappers, core/epoch_loops.py (read, not imported) and the settings samples that pin the benchmark shapes.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("HP_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(REPO, "baseline", "_ref")

FILES = [
    "model/encoder.py", "model/full_model.py", "model/hyper_network.py", "model/target_network.py",
    "losses/champfer_loss.py",
    "utils/__init__.py", "utils/points.py", "utils/metrics.py",
    "utils/pytorch_structural_losses/nn_distance.py", "utils/pytorch_structural_losses/match_cost.py",
    "utils/evaluation/mmd.py",
    "core/epoch_loops.py",
    "settings/config_3depn_airplane.json.sample", "settings/config_completion.json.sample",
]


def _sha(path: str) -> str:
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def stage(check: bool = False) -> int:
    if not os.path.isdir(SRC):
        present = os.path.isfile(os.path.join(DST, "MANIFEST.json"))
        print(f"[stage_reference] {SRC} not present (GPU box?): using the staged copy" if present
              else f"[stage_reference] neither {SRC} nor a staged copy exists: the reference arm falls back to the oracle port")
        return 0
    manifest, bad = {}, []
    for rel in FILES:
        s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
        if not os.path.isfile(s):
            bad.append(rel + " (missing in the reference)")
            continue
        if check:
            if not os.path.isfile(d) or _sha(d) != _sha(s):
                bad.append(rel)
        else:
            os.makedirs(os.path.dirname(d), exist_ok=True)
            shutil.copyfile(s, d)
        manifest[rel] = _sha(s)
    if not check:
        with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
            json.dump({"source": SRC, "files": manifest}, f, indent=1, sort_keys=True)
    if bad:
        print("[stage_reference] problems:", bad, file=sys.stderr)
        return 1
    print(f"[stage_reference] {'checked' if check else 'staged'} {len(manifest)} files -> {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(stage(check="--check" in sys.argv))
