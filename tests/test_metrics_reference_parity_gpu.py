"""GPU: compute_all_metrics against the REFERENCE'S OWN ARITHMETIC at a realistic shape -- its utils/metrics.py (staged
unmodified in baseline/_ref), its pure-torch expansion-form ChamferLoss and its CUDA extension's ApproxMatch + MatchCost,
composed exactly like utils/metrics.py:121-158,194-238 (tests/_reference_metrics_worker.py, a subprocess so that the
reference's top-level `utils` / `losses` packages cannot shadow anything here).

What can differ: this library's CD entries use the direct form, the reference's the expansion form (entries agree
to ~2e-6 absolute), and its EMD is a different launch decomposition (1e-5 on the cost).  COV and 1-NNA are INDEX work on
those matrices, so near-tied clouds can flip an argmin; the test counts the flips instead of hiding them."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference(tmp_path, smp, ref, bs):
    if not os.path.isfile(os.path.join(REPO, "baseline", "_ref", "utils", "metrics.py")):
        pytest.skip("baseline/_ref not staged (python tools/stage_reference.py in the build container)")
    from oracle import oracle as O

    if O.load_reference_ext() is None:
        pytest.skip("oracle/_ref not built")
    fin, fout = str(tmp_path / "in.npz"), str(tmp_path / "out.npz")
    np.savez(fin, smp=smp.numpy(), ref=ref.numpy())
    p = subprocess.run([sys.executable, os.path.join(REPO, "tests", "_reference_metrics_worker.py"), fin, fout, str(bs)],
                       capture_output=True, text=True, timeout=1500, cwd=REPO)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-3000:]
    return np.load(fout)


@pytest.mark.parametrize("n_smp,n_ref,npts,bs", [(64, 64, 2048, 32), (40, 56, 1024, 25)])
def test_compute_all_metrics_vs_reference_arithmetic(hp, tmp_path, n_smp, n_ref, npts, bs):
    g = torch.Generator().manual_seed(n_smp * 1000 + npts)
    # clouds of different spread, so that nearest clouds are well separated for most rows but not all
    smp = (torch.rand(n_smp, npts, 3, generator=g) - 0.5) * (0.6 + 0.4 * torch.rand(n_smp, 1, 1, generator=g))
    ref = (torch.rand(n_ref, npts, 3, generator=g) - 0.5) * (0.6 + 0.4 * torch.rand(n_ref, 1, 1, generator=g))
    R = _reference(tmp_path, smp, ref, bs)
    sd, rd = smp.to(DEV), ref.to(DEV)
    ours = hp.compute_all_metrics(sd, rd, bs, hp.ChamferLoss(), one_nn=True)
    ref_keys = sorted(k[len("metric:"):] for k in R.files if k.startswith("metric:"))
    assert sorted(ours) == ref_keys, (sorted(ours), ref_keys)

    # the matrices themselves
    M_cd, M_emd = hp.metrics._pairwise_EMD_CD_(rd, sd)
    np.testing.assert_allclose(M_cd.cpu().numpy(), R["M_rs_cd"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(M_emd.cpu().numpy(), R["M_rs_emd"], rtol=1e-5)
    # the opt-in shortcut (one ex2 shared by two passes) is measured, not trusted: a few 1e-5 off, documented in csrc/emd.cu
    fast = hp.pairwise_emd(rd, sd, fast=True).cpu().numpy()
    rel = lambda x: float(np.max(np.abs(x - R["M_rs_emd"]) / np.abs(R["M_rs_emd"])))  # noqa: E731
    print("EMDREL " + json.dumps({"shape": [n_smp, n_ref, npts], "default_max_rel": rel(M_emd.cpu().numpy()), "fast_max_rel": rel(fast)}))
    assert rel(fast) < 1e-4
    M_rr_emd = hp.pairwise_emd(rd, rd).cpu().numpy()
    M_ss_cd = hp.pairwise_cd(sd, sd).cpu().numpy()
    np.testing.assert_allclose(M_rr_emd, R["M_rr_emd"], rtol=1e-5)
    np.testing.assert_allclose(M_ss_cd, R["M_ss_cd"], rtol=1e-5, atol=2e-6)

    # value metrics: 1e-5 relative
    for k in ("mmd(Fidelity)-CD", "mmd_smp-CD", "mmd(Fidelity)-EMD", "mmd_smp-EMD"):
        assert float(ours[k]) == pytest.approx(float(R["metric:" + k]), rel=1e-5), k

    # index metrics: count the flips between the two arithmetic forms
    flips = {}
    for tag, mine, theirs in (("CD", M_cd.cpu().numpy(), R["M_rs_cd"]), ("EMD", M_emd.cpu().numpy(), R["M_rs_emd"])):
        a, b = mine.argmin(axis=0), theirs.argmin(axis=0)  # mmd_cov(M.t()): argmin over ref for every sample
        flips["cov_argmin_" + tag] = int((a != b).sum())
        cov_theirs = float(R["metric:cov(Coverage)-" + tag])
        # a flipped argmin changes the unique count by at most one
        assert abs(float(ours["cov(Coverage)-" + tag]) - cov_theirs) <= (flips["cov_argmin_" + tag] + 1e-9) / n_ref, (tag, flips)
    for tag in ("CD", "EMD"):
        for k in ("acc", "acc_t", "acc_f"):
            mine, theirs = float(ours[f"1-NN-{tag}-{k}"]), float(R[f"metric:1-NN-{tag}-{k}"])
            flips[f"1nn_{tag}_{k}_diff_predictions"] = round(abs(mine - theirs) * (n_smp + n_ref if k == "acc" else (n_ref if k == "acc_t" else n_smp)))
    print("FLIPS " + json.dumps({"shape": [n_smp, n_ref, npts], **flips}))
    # near ties are rare at these shapes: a handful of flips at most, none is the expected outcome
    assert flips["cov_argmin_CD"] <= 1 and flips["cov_argmin_EMD"] <= 1, flips
    assert all(v <= 1 for k, v in flips.items() if k.startswith("1nn_")), flips
