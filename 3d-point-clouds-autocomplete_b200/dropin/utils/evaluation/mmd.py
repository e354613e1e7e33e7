"""Drop-in for the compute part of the reference's utils/evaluation/mmd.py (minimum_mathing_distance); its `process`
(file loading) stays in the reference."""
from _pkg import pkg as _hp

minimum_mathing_distance = _hp.evaluation.minimum_mathing_distance
