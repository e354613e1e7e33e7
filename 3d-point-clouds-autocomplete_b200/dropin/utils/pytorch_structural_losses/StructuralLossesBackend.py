"""Drop-in for the reference's compiled pybind module StructuralLossesBackend
(structural_loss.cpp:130-136): same function names, argument order and return lists."""
from _pkg import pkg as _hp

NNDistance = _hp.NNDistance
NNDistanceGrad = _hp.NNDistanceGrad
ApproxMatch = _hp.ApproxMatch
MatchCost = _hp.MatchCost
MatchCostGrad = _hp.MatchCostGrad
