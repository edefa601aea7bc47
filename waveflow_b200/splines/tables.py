"""Basis tables of one spline family: generation/caching on the host, kernel layouts on the device.

The reference loads float64 `.npy` tables through jnp (-> float32, x64 disabled) into a `[4, P, T]` array
(isplines_jax.py:112-131, msplines_jax.py:90-108, bsplines_jax.py:74-116).  Here the same float32 values are kept in the
layouts the kernels want (include/waveflow_b200.h, wf_table_layout_host).
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _ffi
from . import tablegen

_CACHE: dict = {}


class SplineTables:
    def __init__(self, kind: str, k: int, n_internal_knots: int, n_mesh_points: int = 2000, cache_root: str | None = None):
        self.kind, self.k, self.n_internal_knots, self.T = kind, k, n_internal_knots, n_mesh_points
        if kind == "I":
            self.tab64, self.knots = tablegen.build_I_tables(k, n_internal_knots, n_mesh_points, cache_root)
        elif kind == "M":
            self.tab64, self.knots = tablegen.build_M_tables(k, n_internal_knots, n_mesh_points, cache_root)
        elif kind == "B":
            b = tablegen.build_B_tables(k, n_internal_knots, n_mesh_points, cache_root)
            self.tab64, self.knots = b["b"], b["knots"]
            self.ob64, self.b_to_ob64, self.ob_to_b64 = b["ob"], b["b_to_ob"], b["ob_to_b"]
        else:
            raise ValueError(kind)
        self.P = self.tab64.shape[1]
        self.tab32 = self.tab64.astype(np.float32)
        self._dev: dict = {}

    @staticmethod
    def get(kind, k, n_internal_knots, n_mesh_points=2000, cache_root=None) -> "SplineTables":
        key = (kind, k, n_internal_knots, n_mesh_points, cache_root)
        if key not in _CACHE:
            _CACHE[key] = SplineTables(kind, k, n_internal_knots, n_mesh_points, cache_root)
        return _CACHE[key]

    # boundary basis values the reference closes over in enforce_boundary_conditions: {I,M,B}_cached(0.0 / 1.0, j, nd)
    def boundary_value(self, nd: int, j: int, right: bool) -> float:
        return float(self.tab32[nd][j][self.T - 1 if right else 0])

    @staticmethod
    def _pad32(tab32: np.ndarray) -> np.ndarray:
        _, P, T = tab32.shape
        out = np.zeros((T, 4, _ffi.WF_MAX_P), dtype=np.float32)
        if P <= _ffi.WF_MAX_P:
            out[:, :, :P] = np.transpose(tab32, (2, 0, 1))
        return out

    def dev(self, device) -> dict:
        """Device-resident layouts (built once per device)."""
        key = str(device)
        if key not in self._dev:
            d = {}
            dense, rec, lo = _ffi.table_layouts(self.tab32, self.kind)
            d["dense"] = torch.from_numpy(dense).to(device)
            d["rec"] = None if rec is None else torch.from_numpy(rec).to(device)
            d["lo"] = None if lo is None else torch.from_numpy(lo).to(device)
            # [T][8 window slots][4 derivative orders]: one 128-bit load per window slot in the tensor-core live kernels
            d["rec_t"] = None if rec is None else torch.from_numpy(np.ascontiguousarray(np.transpose(rec, (0, 2, 1)))).to(device)
            d["dense32"] = torch.from_numpy(self._pad32(self.tab32)).to(device)
            if self.kind == "B":
                ob32 = self.ob64.astype(np.float32)
                d["ob_dense"] = torch.from_numpy(_ffi.table_layouts(ob32, "B")[0]).to(device)
                d["ob_dense32"] = torch.from_numpy(self._pad32(ob32)).to(device)
                d["ob_to_b"] = torch.from_numpy(self.ob_to_b64.astype(np.float32)).to(device).contiguous()
                d["b_to_ob"] = torch.from_numpy(self.b_to_ob64.astype(np.float32)).to(device).contiguous()
            self._dev[key] = d
        return self._dev[key]
