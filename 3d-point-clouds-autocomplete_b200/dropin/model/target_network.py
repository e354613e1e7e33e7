"""Drop-in for the reference's model/target_network.py (+ the batched op that replaces the per-sample loop
of model/full_model.py:67-74)."""
from _pkg import pkg as _hp

TargetNetwork = _hp.TargetNetwork
target_network_forward = _hp.target_network_forward
generate_points_batched = _hp.generate_points_batched
