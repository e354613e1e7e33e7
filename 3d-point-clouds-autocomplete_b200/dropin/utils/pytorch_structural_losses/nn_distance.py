"""Drop-in for the reference's utils/pytorch_structural_losses/nn_distance.py."""
from _pkg import pkg as _hp

NNDistanceFunction = _hp.NNDistanceFunction
nn_distance = _hp.nn_distance
