"""Fused inverse of the live flow (wf_live_inverse) -- used by flows.Serial.inverse_fun."""
from . import _live


def flow_inverse(spec, weights, u, exact: bool = False):
    return _live.inverse(spec, weights, u, exact=exact)
