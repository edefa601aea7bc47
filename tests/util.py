"""Shared helpers for the parity tests (oracle side lives in oracle/)."""
import numpy as np
import torch


def to_torch_tree(tree, device):
    if isinstance(tree, np.ndarray):
        return torch.from_numpy(np.ascontiguousarray(tree)).to(device)
    if isinstance(tree, (list, tuple)):
        return type(tree)(to_torch_tree(t, device) for t in tree)
    return tree


def relerr(a, b, scale=None):
    """max |a-b| / (|b| + scale) with scale defaulting to max|b| (SURVEY.md section 7 'hard parts')."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    s = np.abs(b).max() if scale is None else scale
    return float(np.max(np.abs(a - b) / (np.abs(b) + s)))


def spec_from_live(m):
    """oracle.live.LiveModel -> waveflow_b200._live.LiveSpec (same static description, product tables)."""
    from waveflow_b200._live import LiveSpec
    from waveflow_b200.splines.tables import SplineTables
    T = m.tab_I.shape[-1]
    n_i = m.tab_I.shape[1] - m.k_i
    tI = SplineTables.get("I", m.k_i, n_i, T)
    tP = None
    prior = None
    if m.prior == "B":
        prior = "B"; tP = SplineTables.get("B", m.k_p, m.tab_OB.shape[1] - m.k_p + 1, T)
    elif m.prior == "M":
        prior = "M"; tP = SplineTables.get("M", m.k_p, m.tab_P.shape[1] - m.k_p + 2, T)
    return LiveSpec(D=m.D, n_layers=m.n_layers, tab_I=tI, k_I=m.k_i, reg=m.reg, tol=m.tol, bc_I_left=m.bc_i_left,
                    bc_I_right=m.bc_i_right, prior=prior, tab_P=tP, k_P=m.k_p, bc_P_left=m.bc_p_left,
                    bc_P_right=m.bc_p_right, box=m.box, coord=m.coord)
