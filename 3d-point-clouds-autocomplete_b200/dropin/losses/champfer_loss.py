"""Drop-in for the reference's losses/champfer_loss.py."""
from _pkg import pkg as _hp

ChamferLoss = _hp.ChamferLoss
