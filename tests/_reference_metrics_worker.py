"""Subprocess of tests/test_metrics_reference_parity_gpu.py: runs the REFERENCE's own utils/metrics.py (staged unmodified under
baseline/_ref by tools/stage_reference.py) with its own pure-torch ChamferLoss and its own CUDA extension (oracle/_ref) on the
GPU, exactly as compute_all_metrics composes them (utils/metrics.py:121-158, 194-238), plus its knn on the three matrices
(the 1-NN block the reference keeps inside a string literal, :224-237).  Writes matrices and metrics to an .npz.

    python tests/_reference_metrics_worker.py <in.npz> <out.npz> <batch_size>
"""
import importlib.util
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(REPO, "baseline", "_ref")
sys.path.insert(0, REPO)
from oracle import oracle as O  # noqa: E402

ext = O.load_reference_ext()
assert ext is not None, "oracle/_ref not built"
sys.path.insert(0, REF)
sys.modules["utils.pytorch_structural_losses.StructuralLossesBackend"] = ext
from losses.champfer_loss import ChamferLoss  # noqa: E402  (the reference's file)
from utils import metrics as M  # noqa: E402               (the reference's file)

assert os.path.samefile(M.__file__, os.path.join(REF, "utils", "metrics.py"))
inp = np.load(sys.argv[1])
bs = int(sys.argv[3])
dev = torch.device("cuda:0")
smp, ref = torch.from_numpy(inp["smp"]).to(dev), torch.from_numpy(inp["ref"]).to(dev)
cl = ChamferLoss().to(dev)
with torch.no_grad():
    res = M.compute_all_metrics(smp, ref, bs, cl)
    M_rs_cd, M_rs_emd = M._pairwise_EMD_CD_(ref, smp, bs, cl)
    M_rr_cd, M_rr_emd = M._pairwise_EMD_CD_(ref, ref, bs, cl)
    M_ss_cd, M_ss_emd = M._pairwise_EMD_CD_(smp, smp, bs, cl)
    knn_cd = M.knn(M_rr_cd, M_rs_cd, M_ss_cd, 1, sqrt=False)
    knn_emd = M.knn(M_rr_emd, M_rs_emd, M_ss_emd, 1, sqrt=False)
out = {("metric:" + k): float(v) for k, v in res.items()}
out.update({("metric:1-NN-CD-" + k): float(v) for k, v in knn_cd.items() if "acc" in k})
out.update({("metric:1-NN-EMD-" + k): float(v) for k, v in knn_emd.items() if "acc" in k})
for name, t in (("M_rs_cd", M_rs_cd), ("M_rs_emd", M_rs_emd), ("M_rr_cd", M_rr_cd), ("M_rr_emd", M_rr_emd), ("M_ss_cd", M_ss_cd),
                ("M_ss_emd", M_ss_emd)):
    out[name] = t.cpu().numpy()
np.savez(sys.argv[2], **out)
print("reference metrics written:", sorted(k for k in out if k.startswith("metric:")))
