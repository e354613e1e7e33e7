"""Drop-in for the reference's utils/pytorch_structural_losses/match_cost.py."""
from _pkg import pkg as _hp

MatchCostFunction = _hp.MatchCostFunction
match_cost = _hp.match_cost
approx_match = _hp.approx_match
