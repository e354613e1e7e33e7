"""Generate tests/golden/cpu_reference.npz by IMPORTING the reference's own Python modules
(pure-torch ChamferLoss, TargetNetwork, mmd_cov, knn) from /root/reference on CPU.

Run in the build container only (the reference tree does not exist on the GPU box):
    python tests/golden/make_golden_cpu.py
The committed .npz is what the tests read; this script is committed so the vectors can be
regenerated and audited.  Nothing here is used by the product path.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("HP_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, REPO)
sys.path.insert(0, REF)

from oracle import oracle as O  # noqa: E402

# utils/metrics.py imports the compiled CUDA backend at import time (metrics.py:14-15);
# the module cross-compiled by oracle/build_ref.sh imports fine without a GPU.
O.build_reference_ext()
ext = O.load_reference_ext()
assert ext is not None, "oracle/_ref not built"
sys.modules["utils.pytorch_structural_losses.StructuralLossesBackend"] = ext

from losses.champfer_loss import ChamferLoss  # noqa: E402
from model.target_network import TargetNetwork  # noqa: E402
from utils import metrics as ref_metrics  # noqa: E402

torch.manual_seed(0)
g = torch.Generator().manual_seed(20261017)
out = {}

cl = ChamferLoss()
cl.use_cuda = False

# 1. lattice clouds: both distance forms are exact, many ties -> index parity is meaningful
a = (torch.randint(-4, 5, (2, 96, 3), generator=g).float() / 8).requires_grad_(True)
b = (torch.randint(-4, 5, (2, 80, 3), generator=g).float() / 8).requires_grad_(True)
P = cl.batch_pairwise_dist(a, b)  # [2, 96, 80]
v2, i2 = P.min(2)  # for each a_i: nearest b
v1, i1 = P.min(1)  # for each b_j: nearest a
loss = cl(b, a)  # forward(preds, gts): P = bpd(gts, preds)
loss.backward()
out.update(lat_a=a.detach().numpy(), lat_b=b.detach().numpy(), lat_P=P.detach().numpy(),
           lat_dist_a=v2.detach().numpy(), lat_idx_a=i2.numpy().astype(np.int32),
           lat_dist_b=v1.detach().numpy(), lat_idx_b=i1.numpy().astype(np.int32),
           lat_loss=loss.detach().numpy(), lat_grad_a=a.grad.numpy().copy(), lat_grad_b=b.grad.numpy().copy())

# 2. uniform clouds in [-0.5, 0.5]^3 (dataset scale), N != M
a = (torch.rand(3, 128, 3, generator=g) - 0.5).requires_grad_(True)
b = (torch.rand(3, 100, 3, generator=g) - 0.5).requires_grad_(True)
P = cl.batch_pairwise_dist(a, b)
v2, i2 = P.min(2)
v1, i1 = P.min(1)
loss = cl(b, a)
loss.backward()
out.update(uni_a=a.detach().numpy(), uni_b=b.detach().numpy(), uni_P=P.detach().numpy(),
           uni_dist_a=v2.detach().numpy(), uni_idx_a=i2.numpy().astype(np.int32),
           uni_dist_b=v1.detach().numpy(), uni_idx_b=i1.numpy().astype(np.int32),
           uni_loss=loss.detach().numpy(), uni_grad_a=a.grad.numpy().copy(), uni_grad_b=b.grad.numpy().copy())

# 3. TargetNetwork 3->32->64->128->64->3 with bias (settings/config_3depn_airplane.json.sample:87-94)
cfg = {"use_bias": True, "layer_out_channels": [32, 64, 128, 64]}
W = O.target_network_num_weights(cfg["layer_out_channels"], True)
assert W == 19011
w = (torch.randn(2, W, generator=g) * 0.15).requires_grad_(True)
x = torch.randn(2, 64, 3, generator=g) * 0.6
G = torch.randn(2, 64, 3, generator=g)
y = torch.stack([TargetNetwork(cfg, w[i])(x[i]) for i in range(2)])
(y * G).sum().backward()
out.update(tn_w=w.detach().numpy(), tn_x=x.numpy(), tn_gout=G.numpy(), tn_y=y.detach().numpy(),
           tn_grad_w=w.grad.numpy().copy())
# no-bias, different widths
cfg2 = {"use_bias": False, "layer_out_channels": [16, 8]}
W2 = O.target_network_num_weights(cfg2["layer_out_channels"], False)
w2 = torch.randn(3, W2, generator=g) * 0.5
x2 = torch.randn(3, 33, 3, generator=g)
y2 = torch.stack([TargetNetwork(cfg2, w2[i])(x2[i]) for i in range(3)])
out.update(tn2_w=w2.numpy(), tn2_x=x2.numpy(), tn2_y=y2.numpy())

# 4. mmd_cov / knn on a random distance matrix with ties
M = torch.randint(0, 40, (12, 10), generator=g).float() / 8
r = ref_metrics.mmd_cov(M)
out.update(mc_M=M.numpy(), mc_mmd=r["mmd(Fidelity)"].numpy(), mc_cov=r["cov(Coverage)"].numpy(),
           mc_mmd_smp=r["mmd_smp"].numpy())
Mxx = torch.rand(7, 7, generator=g)
Mxx = (Mxx + Mxx.t()) / 2
Myy = torch.rand(9, 9, generator=g)
Myy = (Myy + Myy.t()) / 2
Mxy = torch.rand(7, 9, generator=g)
k = ref_metrics.knn(Mxx, Mxy, Myy, 1, sqrt=False)
out.update(knn_Mxx=Mxx.numpy(), knn_Mxy=Mxy.numpy(), knn_Myy=Myy.numpy(),
           **{"knn_" + kk: np.float32(float(vv)) for kk, vv in k.items()})

# 5. target-network input sampling (utils/points.py:16-36) from the global torch CPU RNG, seed 1856
#    (settings/config_3depn_airplane.json.sample:105), three consecutive draws per epoch like the
#    per-sample loop of full_model.py:70-74
from utils.points import generate_points as ref_generate_points  # noqa: E402

pcfg = {"target_network_input": {"normalization": {"enable": True, "type": "progressive", "epoch": 100}}}
for ep in (1, 37, 100, 250):
    torch.manual_seed(1856)
    out[f"gp_ep{ep}"] = torch.stack([ref_generate_points(pcfg, ep, (96, 3)) for _ in range(3)]).numpy()
pcfg_off = {"target_network_input": {"normalization": {"enable": False, "type": "progressive", "epoch": 100}}}
torch.manual_seed(1856)
out["gp_plain"] = torch.stack([ref_generate_points(pcfg_off, 5, (96, 3)) for _ in range(2)]).numpy()

# 6. KD-tree Chamfer of the evaluation scripts (utils/evaluation/chamfer.py:8-32), used by TMD
from utils.evaluation.chamfer import compute_trimesh_chamfer  # noqa: E402

tm = (torch.rand(5, 200, 3, generator=g) - 0.5).numpy()
out["tm_pcs"] = tm
out["tm_cd"] = np.array([[compute_trimesh_chamfer(tm[j], tm[k]) for k in range(5)] for j in range(5)])

# 7. the reference's FullModel.forward end to end on CPU (model/full_model.py:54-80, HyperPocket mode, the 3D-EPN sample
#    config): hypernetwork output and reconstruction of a seeded batch, for the batched replacement of the per-sample loop
import json  # noqa: E402

from model.full_model import FullModel as RefFullModel  # noqa: E402

cfg = json.load(open(os.path.join(REF, "settings", "config_3depn_airplane.json.sample")))
torch.manual_seed(1856)
fm = RefFullModel(cfg["full_model"])
fm.eval()
fm_B, fm_NE, fm_NM, fm_NG, fm_epoch = 3, 80, 48, 128, 37
existing = torch.rand(fm_B, fm_NE, 3, generator=g) - 0.5
missing = torch.rand(fm_B, fm_NM, 3, generator=g) - 0.5
captured = {}


def _after_hypernetwork(mod, inp, o):
    captured["w"] = o.detach().clone()
    torch.manual_seed(99)  # re-seed between the hypernetwork and the per-sample loop (the encoders draw from the same global
    # RNG): the test re-applies this seed before drawing the target-network input clouds


hook = fm.hyper_network.register_forward_hook(_after_hypernetwork)
with torch.no_grad():
    rec = fm(existing.clone(), missing.clone(), [fm_B, fm_NG, 3], fm_epoch, "cpu")
hook.remove()
assert rec.shape == (fm_B, 3, fm_NG) and captured["w"].shape == (fm_B, 19011)
out.update(fm_weights=captured["w"].numpy(), fm_rec=rec.numpy(), fm_meta=np.array([fm_B, fm_NG, fm_epoch, 99]))

# 8. JSD between two sets of clouds (utils/metrics.py:243-359; scikit-learn KD-tree on the occupancy grid)
def _ball(n_clouds, n_pts, scale):
    v = torch.randn(n_clouds, n_pts, 3, generator=g)
    v = v / v.norm(dim=2, keepdim=True) * torch.rand(n_clouds, n_pts, 1, generator=g) ** (1 / 3) * 0.5
    return (v * torch.tensor(scale)).numpy().astype(np.float32)

jsd_smp, jsd_ref = _ball(6, 256, [1.0, 0.6, 0.9]), _ball(5, 256, [0.8, 1.0, 0.7])
out.update(jsd_smp=jsd_smp, jsd_ref=jsd_ref)
for res in (8, 28):
    out[f"jsd_r{res}"] = np.float64(ref_metrics.jsd_between_point_cloud_sets(jsd_smp, jsd_ref, res))
    ent, cnt = ref_metrics.entropy_of_occupancy_grid(jsd_smp, res, True)
    out[f"jsd_ent_r{res}"], out[f"jsd_cnt_r{res}"] = np.float64(ent), cnt.astype(np.float32)
grid_c, spacing = ref_metrics.unit_cube_grid_point_cloud(8, True)
out.update(jsd_grid8=grid_c, jsd_spacing8=np.float64(spacing))

np.savez_compressed(os.path.join(HERE, "cpu_reference.npz"), **out)
print("wrote", os.path.join(HERE, "cpu_reference.npz"), {k: v.shape for k, v in out.items()})
