"""GPU, multi-process: the sharded compute_all_metrics (row blocks of the cloud-distance matrices per rank, only
(min, argmin) vectors gathered) equals the single-process result bit for bit on real kernels.  One rank per GPU over NCCL where
the box has several GPUs (2, and 8 when visible); on a one-GPU box two ranks share cuda:0 and exchange the vectors over gloo."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(REPO, "tests", "_multirank_metrics_worker.py")


def _run(nproc: int, backend: str, case: str, port: int):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), WORKER, backend, case]
    p = subprocess.run(cmd, cwd=REPO, capture_output=True, text=True, timeout=1500)
    lines = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
    assert p.returncode == 0 and lines, f"rc={p.returncode}\n{p.stdout[-2000:]}\n{p.stderr[-3000:]}"
    res = json.loads(lines[-1][len("RESULT "):])
    assert res["all_ranks_ok"] and res["mismatching"] == [] and res["world"] == nproc, res
    return res


def _layouts():
    n = torch.cuda.device_count()
    out = [(2, "nccl")] if n >= 2 else [(2, "gloo")]
    if n >= 8:
        out.append((8, "nccl"))
    return out


def test_sharded_equals_single_process_all_twelve_keys(hp):
    """150 x 131 clouds x 1024 points, CD + EMD + 1-NNA: every one of the 12 result keys identical to the group-of-one run."""
    for i, (nproc, backend) in enumerate(_layouts()):
        res = _run(nproc, backend, "mid", 29611 + i)
        assert res["keys"] == 12, res
        assert 0.0 < res["values"]["cov(Coverage)-CD"] <= 1.0 and 0.0 <= res["values"]["1-NN-EMD-acc"] <= 1.0


def test_sharded_equals_single_process_c5_cd_full_size(hp):
    """BASELINE config C5, CD half: 1000 x 1000 clouds x 2048 points with 1-NNA-CD (2.0 * 10^6 cloud pairs per run)."""
    nproc, backend = _layouts()[-1]
    res = _run(nproc, backend, "c5cd", 29631)
    assert res["keys"] == 6, res
