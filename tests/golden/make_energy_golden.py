"""Golden vectors for psi and the LOCAL ENERGY, produced by the REFERENCE'S OWN SOURCE FILES (build container only):

    python tests/golden/make_energy_golden.py        # -> tests/golden/ref_energy_vectors.npz   (takes a few minutes)

model_factory.get_waveflow_model -> wavefunctions.Waveflow -> flows (BoxTransformLayer, IMADE, Reverse) -> isplines_jax /
bsplines_jax, and utils/physics.construct_hamiltonian_function (H psi = -1/2 trace(jax.hessian(psi)) + V psi) are imported
unmodified from /root/reference and executed on tests/golden/jax_numpy_shim in float64 ("x64 enabled") mode:
jax.hessian = forward over forward mode with nested dual numbers carried through the reference's own arithmetic as numpy
object arrays; every I_cached / B_cached call on a dual number applies THE RULE THE REFERENCE REGISTERED with custom_jvp
(derivative = the next table), at both levels -- which is what defines the reference's Laplacian.
The model uses the basis tables the reference ships (degree 5, 16 internal knots, 2000 mesh points); the parameters the
reference's init functions create (numpy PRNG behind jax.random) are stored in the fixture, flattened in pytree order.
"""
import json
import os
import sys
import tempfile
import types
from pathlib import Path

os.environ["JAX_SHIM_X64"] = "1"
import numpy as np  # noqa: E402

HERE = Path(__file__).resolve().parent
REF = Path("/root/reference")


def leaves(tree, out=None):
    out = [] if out is None else out
    if isinstance(tree, (tuple, list)):
        for t in tree:
            leaves(t, out)
    else:
        out.append(np.asarray(tree, dtype=np.float64))
    return out


def structure(tree, counter=None):
    """nested lists mirroring the pytree, leaves replaced by their index in leaves(tree) (tuples and lists both become lists)"""
    counter = [0] if counter is None else counter
    if isinstance(tree, (tuple, list)):
        return [structure(t, counter) for t in tree]
    counter[0] += 1
    return counter[0] - 1


def main():
    sys.path.insert(0, str(HERE / "jax_numpy_shim"))
    sys.path.insert(1, str(REF))
    for m in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(m, types.ModuleType(m))
    work = tempfile.mkdtemp()
    os.makedirs(work + "/cached_splines_bases")
    os.symlink(REF / "waveflow/tests/splines/cached_bases/I", work + "/cached_splines_bases/I")
    os.symlink(REF / "waveflow/tests/splines/cached_bases/B", work + "/cached_splines_bases/B")
    os.chdir(work)                                    # the reference looks for ./cached_splines_bases/{I,B}/

    from waveflow import model_factory
    from waveflow.utils import physics

    out = {}
    for tag, D, layers, coord, box, n in (("d2_mean", 2, 1, "mean", 4.0, 10), ("d3_first", 3, 1, "first", 5.0, 4),
                                          ("d4_mean_l3", 4, 3, "mean", 10.0, 2)):       # the bench's D, depth and box
        init = model_factory.get_waveflow_model(D, base_spline_degree=5, i_spline_degree=5, n_prior_internal_knots=16,
                                                n_i_internal_knots=16, n_flow_layers=layers, box_size=box, xu_coord_type=coord)
        params, psi, log_pdf, sample = init(3 + D, D)
        protons = np.zeros((D, 1))
        h_fn = physics.construct_hamiltonian_function(psi, protons=protons, n_space_dimensions=1, eps=0.0)
        rng = np.random.default_rng(40 + D)
        x = np.sort(rng.uniform(-0.8 * box, 0.8 * box, (n, D)), axis=-1)
        p = np.asarray(psi(params, x), dtype=np.float64)
        hp = np.asarray(h_fn(params, x), dtype=np.float64)[:, 0]
        lp = np.asarray(log_pdf(params, x), dtype=np.float64)
        out.update({f"{tag}_x": x, f"{tag}_psi": p, f"{tag}_hpsi": hp, f"{tag}_logpdf": lp, f"{tag}_box": np.float64(box),
                    f"{tag}_D": np.int32(D), f"{tag}_layers": np.int32(layers), f"{tag}_coord_mean": np.int32(coord == "mean")})
        for i, leaf in enumerate(leaves(params)):
            out[f"{tag}_param{i:03d}"] = leaf
        out[f"{tag}_treedef"] = np.array(json.dumps(structure(params)))
        print(tag, "psi", p[:3], "hpsi", hp[:3], flush=True)
        if tag in ("d2_mean", "d4_mean_l3"):
            # ---- the training loss of vqmc.py:192-212 and its derivative along random PARAMETER directions: the parameters are
            # seeded with dual numbers (the lowest differentiation level) and the reference's own loss_fn_efficient -- with the
            # gradient estimator it registers through custom_jvp -- is evaluated on them; <grad loss, v> is the tangent of the result
            from waveflow import vqmc as ref_vqmc
            from jax._core import Dual, JArr, new_tag
            xb, ra = x[:4], -0.3
            n_dirs = 2 if tag == "d2_mean" else 1
            out[f"{tag}_loss_x"], out[f"{tag}_loss_running_average"] = xb, np.float64(ra)
            out[f"{tag}_loss"] = np.float64(np.asarray(ref_vqmc.loss_fn_efficient(params, psi, h_fn, xb, ra)))
            drng = np.random.default_rng(77)
            for k in range(n_dirs):
                tagp = new_tag()
                dirs = [drng.standard_normal(leaf.shape) for leaf in leaves(params)]
                it = iter(dirs)

                def seed(leaf):
                    v = next(it)
                    a = np.asarray(leaf, dtype=np.float64)
                    o = np.empty(a.shape, dtype=object)
                    for idx in np.ndindex(a.shape):
                        o[idx] = Dual(float(a[idx]), float(v[idx]), tagp)
                    return o.view(JArr)

                def map_tree(t):
                    return type(t)(map_tree(u) for u in t) if isinstance(t, (tuple, list)) else seed(t)

                res = ref_vqmc.loss_fn_efficient(map_tree(params), psi, h_fn, xb, ra)
                res = res.item() if isinstance(res, np.ndarray) else res
                assert isinstance(res, Dual) and res.tag == tagp
                out[f"{tag}_dloss{k}"] = np.float64(res.t)
                for i, v in enumerate(dirs):
                    out[f"{tag}_dir{k}_{i:03d}"] = v
                print(tag, "loss", float(res.v), "directional derivative", k, float(res.t), flush=True)
    np.savez_compressed(HERE / "ref_energy_vectors.npz", **out)
    print("wrote", HERE / "ref_energy_vectors.npz")


if __name__ == "__main__":
    main()
