"""Regenerates tests/golden/*.npz|json from the read-only reference checkout.

Run in the BUILD container only (needs /root/reference; the GPU box has no reference):

    python tests/golden/make_golden.py

What it extracts:
  1. waveflow/tests/splines/cached_bases/{I,B}/*.npy  -- the reference's shipped basis tables
     (degree 5, 16 internal knots, 2000 mesh points), stored losslessly (float64, compressed) plus sha256.
  2. the reference's own pure-numpy generator (waveflow/splines/splines_np.py, imported with matplotlib
     stubbed) evaluated for the He run's configuration (degree 6, 23 knots) -> sha256 of every table and a
     strided sample of values, so the product generator can be pinned without the reference present.
  3. data_submission_apl_ml/He_1d_L10box_batch256/checkpoints (a JAX pickle, read without jax) -> plain
     arrays, together with the published psi grids evaluated with exactly those parameters.
Nothing here is product code; no reference source is copied.
"""
import hashlib
import json
import pickle
import sys
import types
from multiprocessing import Pool
from pathlib import Path

import numpy as np

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ----------------------------------------------------------------------------- 1. shipped tables
def shipped_tables():
    root = REF / "waveflow/tests/splines/cached_bases"
    d = {}
    for nd in range(4):
        d[f"I_nd{nd}"] = np.load(root / f"I/degree_5_niknots_21_nmp_2000_nd_{nd}.npy")
        d[f"B_nd{nd}"] = np.load(root / f"B/b_degree_5_niknots_21_nmp_2000_nd_{nd}.npy")
        d[f"OB_nd{nd}"] = np.load(root / f"B/ob_degree_5_niknots_21_nmp_2000_nd_{nd}.npy")
    d["b_to_ob"] = np.load(root / "B/degree_5_niknots_21_nmp_2000_b_to_ob.npy")
    d["ob_to_b"] = np.load(root / "B/degree_5_niknots_21_nmp_2000_ob_to_b.npy")
    np.savez_compressed(OUT / "ref_tables_deg5_k16.npz", **d)
    return {k: sha(v) for k, v in d.items()}


# ----------------------------------------------------------------------------- 2. reference generator
def _import_ref_generator():
    sys.modules.setdefault("matplotlib", types.ModuleType("matplotlib"))
    sys.modules.setdefault("matplotlib.pyplot", types.ModuleType("matplotlib.pyplot"))
    sys.path.insert(0, str(REF))
    from waveflow.splines import splines_np
    return splines_np


def _row(args):
    kind, k, i, t, nd, T = args
    sp = _import_ref_generator()
    mesh = np.linspace(0, 1, T)
    if kind == "I":
        return np.array([sp.I(x, k, i, t, k + 1, n_derivatives=nd) for x in mesh], dtype=np.float64)
    if kind == "B":
        return np.array([sp.B(x, k, i, t, k, n_derivatives=nd) for x in mesh], dtype=np.float64)
    return np.array([sp.M(x, k, i, t, k, n_derivatives=nd) for x in mesh], dtype=np.float64)


def ref_generated(kind, k, n, T=2000):
    sys.path.insert(0, str(OUT.parents[1]))
    from waveflow_b200.splines import tablegen as tg
    t = {"I": tg.knots_I, "B": tg.knots_B, "M": tg.knots_M}[kind](k, n)
    nb = {"I": len(t) - k, "B": len(t) - k - 1, "M": len(t) - k}[kind]
    jobs = [(kind, k, i, t, nd, T) for nd in range(4) for i in range(nb)]
    with Pool(8) as p:
        rows = p.map(_row, jobs, chunksize=1)
    return np.stack(rows).reshape(4, nb, T)


# ----------------------------------------------------------------------------- 3. He checkpoint
class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module == "jax._src.array" and name == "_reconstruct_array":
            def rec(fun, args, arr_state, aval_state):
                a = fun(*args); a.__setstate__(arr_state); return a
            return rec
        if module.startswith("numpy.core"):
            module = module.replace("numpy.core", "numpy._core")
        return super().find_class(module, name)


def he_checkpoint():
    base = REF / "data_submission_apl_ml/He_1d_L10box_batch256"
    with open(base / "checkpoints", "rb") as f:
        (transform_params, sp_params), epoch = _Unpickler(f).load()
    d = {"epoch": np.int64(epoch)}

    def put(prefix, net):
        nn, zero = net
        (W1, b1), _, (W2, b2), _, (W3, b3) = nn
        for nm, a in [("W1", W1), ("b1", b1), ("W2", W2), ("b2", b2), ("W3", W3), ("b3", b3), ("zero", zero)]:
            d[f"{prefix}_{nm}"] = np.asarray(a)
    li = 0
    for p in transform_params:
        if len(p):
            put(f"imade{li}", p); li += 1
    put("prior", sp_params)
    d["n_imade"] = np.int64(li)
    out = base / "outputs"
    d["psi_grid"] = np.load(out / "wavefunctions_2d/values_epoch100000.npy")
    for nm in ["onproton", "random"]:
        d[f"{nm}_coord"] = np.load(out / f"density_1e/{nm}_coord_epoch100000.npy")
        d[f"{nm}_values"] = np.load(out / f"density_1e/{nm}_values_epoch100000.npy")
    d["samples"] = np.load(out / "sample_points/values_epoch100000.npy")
    d["loss_tail"] = np.load(base / "loss.npy")[-2000:]
    np.savez_compressed(OUT / "he_checkpoint_epoch100000.npz", **d)
    return {k: list(np.shape(v)) for k, v in d.items()}


if __name__ == "__main__":
    meta = {"shipped_sha256": shipped_tables()}
    gen = {}
    samples = {}
    for kind, k, n in [("I", 6, 23), ("B", 6, 23), ("M", 3, 15), ("I", 5, 23)]:
        tab = ref_generated(kind, k, n)
        key = f"{kind}_deg{k}_k{n}"
        gen[key] = {"shape": list(tab.shape), "sha256": [sha(tab[nd]) for nd in range(4)]}
        samples[key] = tab[:, :, ::97].copy()
        print(key, tab.shape, flush=True)
    np.savez_compressed(OUT / "ref_generated_samples.npz", **samples)
    meta["ref_generated"] = gen
    meta["he_checkpoint"] = he_checkpoint()
    (OUT / "golden_meta.json").write_text(json.dumps(meta, indent=1))
    print("done")
