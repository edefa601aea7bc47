"""Golden vectors for the rational-quadratic-spline family, produced by the REFERENCE'S OWN SOURCE FILE.

Run in the BUILD container only (needs /root/reference):

    python tests/golden/make_rqs_golden.py        # -> tests/golden/ref_rqs_vectors.npz

waveflow/flows/bijections/neural_splines.py cannot be imported with a current JAX (it uses the removed jax.ops API, and JAX
is not installable here anyway).  It only needs array primitives, though: with tests/golden/jax_numpy_shim first on sys.path
its unmodified source runs on numpy with JAX's float32 ("x64 disabled") dtype rules.  This script loads that file by path,
calls searchsorted / RQS / unconstrained_RQS / NeuralSplineCoupling on seeded inputs and stores inputs, parameters and
outputs.  searchsorted is wrapped (not modified) to record the knot vectors and bin indices RQS computes internally.
No reference source is copied into the repository; the fixture holds arrays only.
"""
import importlib.util
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF_FILE = Path("/root/reference/waveflow/flows/bijections/neural_splines.py")


def load_reference():
    sys.path.insert(0, str(HERE / "jax_numpy_shim"))
    for name in [m for m in sys.modules if m == "jax" or m.startswith("jax.")]:
        del sys.modules[name]
    spec = importlib.util.spec_from_file_location("ref_neural_splines", REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def operator_case(mod, rng, n, K, B, tag, out, scale=1.5):
    """unconstrained_RQS forward and inverse on raw (un-normalised) parameters; bins and knots recorded from searchsorted."""
    x = rng.uniform(-1.3 * B, 1.3 * B, n).astype(np.float32)                 # ~23 % outside the interval (identity tails)
    uw = (rng.standard_normal((n, K)) * scale).astype(np.float32)      # scale 1.5: bins down to the 1e-3 floor (ill-conditioned
    uh = (rng.standard_normal((n, K)) * scale).astype(np.float32)      # in float32); scale 0.3: well-conditioned
    ud = rng.standard_normal((n, K - 1)).astype(np.float32)
    x[:8] = np.float32([-B, B, 0.0, -B * 0.999999, B * 0.999999, 1e-7, -1e-7, B * 0.5])   # interval ends and near-zero inputs
    rec = []
    orig = mod.searchsorted

    def recording(bin_locations, inputs, eps=1e-6):
        r = orig(bin_locations, inputs, eps)
        rec.append((np.array(bin_locations, copy=True), np.array(r, copy=True)))
        return r

    mod.searchsorted = recording
    try:
        y, ld = mod.unconstrained_RQS(x, uw, uh, ud, inverse=False, tail_bound=B)
        knots_f, bins_f = rec[-1]
        xi, ldi = mod.unconstrained_RQS(y, uw, uh, ud, inverse=True, tail_bound=B)
        knots_i, bins_i = rec[-1]
    finally:
        mod.searchsorted = orig
    inside = (x >= -B) & (x <= B)
    inside_i = (y >= -B) & (y <= B)
    for k, v in dict(x=x, uw=uw, uh=uh, ud=ud, y=y, ld=ld, xi=xi, ldi=ldi, inside=inside, inside_inv=inside_i, knots_fwd=knots_f,
                     bins_fwd=bins_f, knots_inv=knots_i, bins_inv=bins_i, K=np.int32(K), B=np.float32(B)).items():
        out[f"{tag}_{k}"] = np.asarray(v)
    assert y.dtype == np.float32 and ld.dtype == np.float32 and knots_f.dtype == np.float32, (y.dtype, ld.dtype, knots_f.dtype)


def coupling_case(mod, seed, dim, K, B, hidden, n, tag, out):
    """NeuralSplineCoupling(K, B, hidden): the layer's own init (weights stored), direct_fun and inverse_fun."""
    params, direct, inverse = mod.NeuralSplineCoupling(K=K, B=B, hidden_dim=hidden)(seed, dim)
    rng = np.random.default_rng(seed + 1)
    x = rng.uniform(-1.2 * B, 1.2 * B, (n, dim)).astype(np.float32)
    y, ld = direct(params, x)
    xi, ldi = inverse(params, y)
    for which, net in zip(("f1", "f2"), params):
        layers = [p for p in net if len(p)]
        assert len(layers) == 3
        for li, (W, b) in enumerate(layers):
            out[f"{tag}_{which}_W{li}"] = W
            out[f"{tag}_{which}_b{li}"] = b
    for k, v in dict(x=x, y=y, ld=ld, xi=xi, ldi=ldi, K=np.int32(K), B=np.float32(B), hidden=np.int32(hidden)).items():
        out[f"{tag}_{k}"] = np.asarray(v)
    assert y.dtype == np.float32 and ld.dtype == np.float32, (y.dtype, ld.dtype)


def main():
    mod = load_reference()
    out = {}
    rng = np.random.default_rng(2024)
    operator_case(mod, rng, 2048, 8, 3.0, "op_k8", out)
    operator_case(mod, rng, 2048, 32, 3.0, "op_k32", out)
    operator_case(mod, rng, 2048, 32, 3.0, "op_k32_mild", out, scale=0.3)
    operator_case(mod, rng, 512, 5, 1.0, "op_k5", out)
    coupling_case(mod, 11, 2, 8, 3.0, 64, 1024, "cpl_d2", out)
    coupling_case(mod, 12, 8, 32, 3.0, 64, 512, "cpl_d8", out)
    np.savez_compressed(HERE / "ref_rqs_vectors.npz", **out)
    print("wrote", HERE / "ref_rqs_vectors.npz", {k: v.shape for k, v in out.items() if k.endswith(("_y", "_bins_fwd"))})


if __name__ == "__main__":
    main()
