"""Golden vectors for BASELINE configs[0] (2-D MFlow density, benchmark_tests.get_model('MFlow')), produced by the REFERENCE'S
OWN SOURCE FILES (build container only):

    python tests/golden/make_mflow_golden.py        # -> tests/golden/ref_mflow_vectors.npz

flows.MFlow(Serial(IMADE, Reverse) x L, masked transform, M-spline prior) is built by the reference's unmodified
flows/distributions.py, flows/bijections/made.py, model_factory.get_masked_transform, splines/{isplines,msplines}_jax.py on the
numpy stand-in for jax (tests/golden/jax_numpy_shim), once in float32 (JAX's default) and once in float64.  The I tables are the
ones the reference ships (degree 5, 16 knots); the M tables (degree 3, 15 knots) are not shipped: the reference's own generator
(splines_np.M) computes them into a scratch directory, and they are stored here as well.  log_pdf (and the flow output u) of
seeded points in [0, 1]^2 are recorded together with the parameters the reference's init functions created.
"""
import json
import os
import subprocess
import sys
import tempfile
import types
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF = Path("/root/reference")


def leaves(tree, out=None):
    out = [] if out is None else out
    if isinstance(tree, (tuple, list)):
        for t in tree:
            leaves(t, out)
    else:
        out.append(np.asarray(tree))
    return out


def structure(tree, counter=None):
    counter = [0] if counter is None else counter
    if isinstance(tree, (tuple, list)):
        return [structure(t, counter) for t in tree]
    counter[0] += 1
    return counter[0] - 1


def run(mode):
    sys.path.insert(0, str(HERE / "jax_numpy_shim"))
    sys.path.insert(1, str(REF))
    for m in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(m, types.ModuleType(m))
    work = os.environ["MFLOW_WORK"]
    os.chdir(work)
    from waveflow import flows
    from waveflow.model_factory import get_masked_transform
    D, L, k_i, n_i, reg, k_p, n_p = 2, 2, 5, 16, 0.02, 3, 15
    init = flows.MFlow(
        flows.Serial(*(flows.IMADE(get_masked_transform(), spline_degree=k_i, n_internal_knots=n_i, spline_regularization=reg,
                                   reverse_fun_tol=1e-6), flows.Reverse()) * L),
        get_masked_transform(), spline_degree=k_p, n_internal_knots=n_p)
    params, log_pdf, sample = init(21, D)
    rng = np.random.default_rng(5)
    x = rng.uniform(0.02, 0.98, (256, D))
    x = x.astype(np.float32) if mode == "f32" else x
    lp, u = log_pdf(params, x, return_sample=True)
    # the inverse direction (sampling path) of the same flow: Serial(IMADE, Reverse).inverse_fun -- IMADE's inverse is the vmapped
    # bisection of helpers.binary_search and conditions on its INPUT -- rebuilt with the rng MFlow handed down
    from jax import random
    _, trng = random.split(21)
    tparams, _direct, inverse = flows.Serial(*(flows.IMADE(get_masked_transform(), spline_degree=k_i, n_internal_knots=n_i,
                                                          spline_regularization=reg, reverse_fun_tol=1e-6), flows.Reverse()) * L)(trng, D)
    assert all(np.array_equal(a, b) for a, b in zip(leaves(tparams), leaves(params[0])))
    import jax.numpy as jnp                                     # (the stand-in) arrays with jax's functional .at[] updates
    x_back = np.asarray(inverse(tparams, jnp.asarray(np.asarray(u)))[0])
    # box transform, both coordinate types, both directions (made.py:108-190; reverse_fun_mean is quirk Q2)
    box = {}
    for coord, Db, Lb in (("mean", 2, 4.0), ("first", 3, 5.0)):
        _, bdirect, breverse = flows.BoxTransformLayer(Lb, xu_coord_type=coord)(0, Db)
        xb = np.sort(rng.uniform(-0.9 * Lb, 0.9 * Lb, (64, Db)), axis=-1)
        xb = xb.astype(np.float32) if mode == "f32" else xb
        ub, ldb = bdirect((), jnp.asarray(xb))
        box[f"box_{coord}_x"], box[f"box_{coord}_u"], box[f"box_{coord}_ld"] = np.asarray(xb), np.asarray(ub), np.asarray(ldb)
        box[f"box_{coord}_back"] = np.asarray(breverse((), jnp.asarray(np.asarray(ub)))[0])
        box[f"box_{coord}_L"] = np.float64(Lb)
    out = {"x": np.asarray(x), "logpdf": np.asarray(lp), "u": np.asarray(u), "x_back": x_back, **box,
           "treedef": np.array(json.dumps(structure(params))),
           "cfg": np.array(json.dumps(dict(D=D, L=L, k_i=k_i, n_i=n_i, reg=reg, k_p=k_p, n_p=n_p)))}
    for i, leaf in enumerate(leaves(params)):
        out[f"param{i:03d}"] = leaf
    for f in sorted(Path(work, "cached_splines_bases", "M").glob("*.npy")):
        out["Mtab_" + f.stem] = np.load(f)
    np.savez_compressed(Path(work) / f"out_{mode}.npz", **out)
    print(mode, "logpdf", np.asarray(lp)[:3], np.asarray(lp).dtype, flush=True)


def main():
    if len(sys.argv) > 1:
        return run(sys.argv[1])
    work = tempfile.mkdtemp()
    os.makedirs(work + "/cached_splines_bases/M")
    os.symlink(REF / "waveflow/tests/splines/cached_bases/I", work + "/cached_splines_bases/I")
    merged = {}
    for mode in ("f64", "f32"):                       # separate processes (the dtype mode is fixed at import); float64 first, so that
                                                      # the M tables the reference generates into the scratch cache are float64
        env = dict(os.environ, MFLOW_WORK=work, JAX_SHIM_X64="1" if mode == "f64" else "0")
        subprocess.run([sys.executable, __file__, mode], check=True, env=env)
        g = np.load(Path(work) / f"out_{mode}.npz")
        for k in g.files:
            if k.startswith("Mtab_") or k in ("treedef", "cfg"):
                merged[k] = g[k]
            elif k.startswith("param"):
                if mode == "f64":
                    merged[k] = g[k]                   # float64 copy of the same draws (numpy PRNG, cast per mode)
            else:
                merged[f"{mode}_{k}"] = g[k]
    np.savez_compressed(HERE / "ref_mflow_vectors.npz", **merged)
    print("wrote", HERE / "ref_mflow_vectors.npz", sorted(k for k in merged if not k.startswith("param"))[:12])


if __name__ == "__main__":
    main()
