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


PARITY_LOG: list = []      # one record per assert_fp32_grade call; tests/conftest.py writes it out at the end of the session


def assert_fp32_grade(got, ref64, ref32, rtol, scale=None, name="", max_slack=4.0):
    """The GPU result must agree with the float64 oracle to `rtol` (relative to |ref| + scale) -- or, where float32
    arithmetic itself cannot (ill-conditioned points: log(p + 1e-7) next to a node of psi, log-dets of tiny bins), be as
    accurate as the reference's own float32 arithmetic, i.e. the numpy restatement run in float32 (`ref32`):
    median, 99th percentile and maximum error at most 2x / 2x / `max_slack`x those of the float32 restatement."""
    got, ref64, ref32 = [np.asarray(a, dtype=np.float64) for a in (got, ref64, ref32)]
    s = np.abs(ref64).max() if scale is None else scale
    eg = np.abs(got - ref64) / (np.abs(ref64) + s)
    eo = np.abs(ref32 - ref64) / (np.abs(ref64) + s)
    assert np.all(np.isfinite(got)), name
    # north_star states FLAT tolerances (1e-5 log-prob / inverse, 1e-4 local energy): record the fraction of elements that
    # meet `rtol` outright, for this implementation and for the reference's own float32 arithmetic, next to the error statistics
    PARITY_LOG.append({"name": name, "n": int(eg.size), "flat_tol": float(rtol), "frac_within_flat_tol": float(np.mean(eg <= rtol)),
                       "frac_within_flat_tol_float32_restatement": float(np.mean(eo <= rtol)),
                       "median": float(np.median(eg)), "p99": float(np.quantile(eg, 0.99)), "max": float(eg.max()),
                       "median_float32_restatement": float(np.median(eo)), "max_float32_restatement": float(eo.max())})
    assert np.median(eg) <= max(rtol / 10, 2 * np.median(eo)), (name, "median", np.median(eg), np.median(eo))
    assert np.quantile(eg, 0.99) <= max(rtol, 2 * np.quantile(eo, 0.99)), (name, "p99", np.quantile(eg, 0.99), np.quantile(eo, 0.99))
    assert eg.max() <= max(rtol, max_slack * eo.max()), (name, "max", eg.max(), eo.max())


def record_flat(name, got, ref64, tol, ref32=None):
    """Record (no assertion) the fraction of elements with |got - ref| <= tol * |ref| -- north_star's flat relative tolerance --
    and, when given, the same fraction for the reference's own float32 arithmetic."""
    got, ref64 = np.asarray(got, dtype=np.float64), np.asarray(ref64, dtype=np.float64)
    eg = np.abs(got - ref64) / (np.abs(ref64) + 1e-300)
    rec = {"name": name, "n": int(eg.size), "flat_tol": float(tol), "frac_within_flat_tol": float(np.mean(eg <= tol)),
           "frac_within_flat_tol_float32_restatement": float("nan"), "median": float(np.median(eg)),
           "p99": float(np.quantile(eg, 0.99)), "max": float(eg.max())}
    if ref32 is not None:
        eo = np.abs(np.asarray(ref32, dtype=np.float64) - ref64) / (np.abs(ref64) + 1e-300)
        rec["frac_within_flat_tol_float32_restatement"] = float(np.mean(eo <= tol))
        rec["median_float32_restatement"] = float(np.median(eo)); rec["max_float32_restatement"] = float(eo.max())
    PARITY_LOG.append(rec)
    return rec
