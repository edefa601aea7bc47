"""The C-ABI library loads (no GPU needed) and exports every symbol include/waveflow_b200.h declares."""
import ctypes
import re
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]


def declared_symbols():
    text = (ROOT / "include" / "waveflow_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(wf_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for must in ["wf_spline_apply_local", "wf_spline_apply_dense", "wf_rqs_apply", "wf_live_forward", "wf_local_energy",
                 "wf_spline_reverse", "wf_enforce_bc", "wf_remove_bias", "wf_bspline_apply", "wf_table_layout_host"]:
        assert must in syms


def test_library_exports_every_declared_symbol():
    from waveflow_b200 import _ffi
    lib = ctypes.CDLL(str(_ffi.LIB_PATH))
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    assert _ffi.lib.wf_abi_version(None) == _ffi.ABI_VERSION == 2
    arch = ctypes.c_int(0)
    _ffi.lib.wf_abi_version(ctypes.byref(arch))
    assert arch.value == 100
    assert _ffi.lib.wf_status_string(-2) == b"unsupported configuration"


def test_no_cpu_fallback():
    import pytest
    import torch
    from waveflow_b200 import _ffi
    with pytest.raises(_ffi.WaveflowB200Error):
        _ffi.ptr(torch.zeros(4))
    with pytest.raises(_ffi.WaveflowB200Error):
        _ffi.f32(torch.zeros(4))


def test_table_layout_host_reconstructs_dense_tables():
    from waveflow_b200 import _ffi
    from waveflow_b200.splines.tables import SplineTables
    for kind, k, n in [("I", 6, 23), ("M", 3, 15), ("B", 6, 23), ("I", 5, 16)]:
        t = SplineTables.get(kind, k, n)
        dense, rec, lo = _ffi.table_layouts(t.tab32, kind)
        assert rec is not None, (kind, k, n)
        P, T = t.P, t.T
        assert np.array_equal(dense[:, :, :P], np.transpose(t.tab32, (2, 0, 1)))
        rebuilt = np.zeros((T, 4, P), dtype=np.float32)
        if kind == "I":
            for m in range(T):
                rebuilt[m, 0, :lo[m]] = 1.0
        for m in range(T):
            hi = min(P, lo[m] + 8)
            rebuilt[m, :, lo[m]:hi] = rec[m, :, :hi - lo[m]]
        assert np.array_equal(rebuilt, np.transpose(t.tab32, (2, 0, 1))), (kind, k, n)
        assert np.all(np.diff(lo) >= 0) and np.all(np.diff(lo) <= 1)
        assert np.all(rec[:, :, 7] == 0)          # one spare slot so that a one-basis shift stays inside the window
    # orthonormalised B tables are dense: no compact form
    t = SplineTables.get("B", 6, 23)
    _, rec, _ = _ffi.table_layouts(t.ob64.astype(np.float32), "B")
    assert rec is None
