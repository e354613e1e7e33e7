"""Drop-in for the hot-path part of the reference's utils/metrics.py (MMD / COV / 1-NNA).
JSD and the mAP helpers of that file are CPU code outside the hot path and are not provided."""
from _pkg import pkg as _hp

_m = _hp.metrics
match_cost = _hp.match_cost
nn_distance = _hp.nn_distance
earth_mover_distance = _m.earth_mover_distance
emd_approx = _m.emd_approx
dist_chamfer = _m.dist_chamfer
_pairwise_EMD_CD_ = _m._pairwise_EMD_CD_
knn = _m.knn
mmd_cov = _m.mmd_cov
compute_all_metrics = _m.compute_all_metrics
