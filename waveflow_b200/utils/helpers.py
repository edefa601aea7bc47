"""Training artefact writers -- reference: utils/helpers.py:13-89 (wavefunction checkpoints), :170-214 (benchmark checkpoints).

Same directory layout, file names, array shapes and dtypes as the reference writes, so its plot scripts and the published
data directories (data_submission_apl_ml/*) interoperate.  Arrays are evaluated on the GPU through the functions handed in
(psi / log_pdf / sample are this package's closures) and stored as numpy; the parameter pickle holds numpy arrays in the
reference's pytree structure.  `load_checkpoint` is the resume path (the reference's own restart branch is overwritten a few
lines later, vqmc.py:69-93).
"""
from __future__ import annotations

import pickle
from pathlib import Path

import numpy as np
import torch

from . import physics
from .coordinates import get_num_inversion_count


def make_result_dirs(save_dir):
    """utils/helpers.py:13-31."""
    for sub in ("", "figures/eigenfunctions", "figures/densities_random", "figures/densities_on_proton", "outputs/wavefunctions_2d",
                "outputs/sample_points", "outputs/density_1e"):
        Path(f"{save_dir}/{sub}").mkdir(parents=True, exist_ok=True)


def _to_numpy_tree(tree):
    if isinstance(tree, torch.Tensor):
        return tree.detach().cpu().numpy()
    if isinstance(tree, (tuple, list)):
        return type(tree)(_to_numpy_tree(t) for t in tree)
    return tree


def _psi_signed(psi, params, coords: np.ndarray, device) -> np.ndarray:
    """psi on unsorted coordinates: sort each row, multiply by (-1)**inversions (utils/helpers.py:55-59)."""
    inv = get_num_inversion_count(coords)
    x = torch.from_numpy(np.sort(coords, axis=-1).astype(np.float32)).to(device)
    z = psi(params, x).detach().cpu().numpy()
    return z * ((-1.0) ** inv).astype(np.float32)


def create_checkpoint_wavefunc(rng, save_dir, psi, sample, params, epoch, loss, energies, system_dict, ngrid=100, nsample=250,
                               device="cuda"):
    """utils/helpers.py:33-89: pickle (params, epoch), loss / energies, psi on the ngrid x ngrid plane of the first two
    coordinates' box, the two one-electron cuts and nsample sample points."""
    make_result_dirs(save_dir)
    with open(f"{save_dir}/checkpoints", "wb") as f:
        pickle.dump((_to_numpy_tree(params), epoch), f)
    np.save(f"{save_dir}/loss.npy", np.asarray(loss))
    np.save(f"{save_dir}/energies.npy", np.asarray(energies))
    box_length, n_particle = system_dict["box_length"], system_dict["n_particle"]
    n_space_dimension, system_name = system_dict["n_space_dimension"], system_dict["system_name"]
    protons, _ = physics.system_catalogue[n_space_dimension][system_name]
    grid = np.linspace(-box_length, box_length, ngrid)
    if n_particle * n_space_dimension == 2:            # the reference's 2-D plane only exists for two coordinates
        y, x = np.meshgrid(grid, grid)
        coordinates = np.stack([x, y], axis=-1).reshape(-1, 2)
        np.save(f"{save_dir}/outputs/wavefunctions_2d/values_epoch{epoch}.npy", _psi_signed(psi, params, coordinates, device))
    one = f"{save_dir}/outputs/density_1e"
    x = sample(rng, params, 1).detach().cpu().numpy().astype(np.float32)
    x = np.repeat(x, ngrid, axis=0)
    x[:, 0] = grid
    np.save(f"{one}/random_values_epoch{epoch}.npy", _psi_signed(psi, params, x, device))
    np.save(f"{one}/random_coord_epoch{epoch}.npy", x)
    x = np.ones((1, n_particle * n_space_dimension), dtype=np.float32) * np.asarray(protons, dtype=np.float32).reshape(-1)[0]
    x = np.repeat(x, ngrid, axis=0)
    x[:, 0] = grid
    np.save(f"{one}/onproton_values_epoch{epoch}.npy", _psi_signed(psi, params, x, device))
    np.save(f"{one}/onproton_coord_epoch{epoch}.npy", x)
    pts = sample(rng, params, nsample).detach().cpu().numpy()
    np.save(f"{save_dir}/outputs/sample_points/values_epoch{epoch}.npy", pts)


def load_checkpoint(save_dir):
    """-> (params pytree of numpy arrays, epoch, loss list, energies list): the resume path of vqmc.py:69-73.  Also reads
    the reference's own pickles (jax arrays are rebuilt as numpy, numpy.core is mapped to numpy._core)."""
    class _Unpickler(pickle.Unpickler):
        def find_class(self, module, name):
            if module.startswith("jax") and name == "_reconstruct_array":
                def rebuild(fun, args, arr_state, aval_state=None):
                    a = fun(*args)
                    a.__setstate__(arr_state)
                    return a
                return rebuild
            if module.startswith("numpy.core"):
                module = module.replace("numpy.core", "numpy._core", 1)
            return super().find_class(module, name)

    with open(f"{save_dir}/checkpoints", "rb") as f:
        params, epoch = _Unpickler(f).load()
    loss = np.load(f"{save_dir}/loss.npy").tolist() if Path(f"{save_dir}/loss.npy").exists() else [0.0]
    energies = np.load(f"{save_dir}/energies.npy").tolist() if Path(f"{save_dir}/energies.npy").exists() else []
    return params, epoch, loss, energies


def moving_average(running_average, new_data, beta):
    """utils/helpers.py:122-123."""
    return running_average - beta * (running_average - new_data)


def uniform_sliding_average(data, window):
    """utils/helpers.py:127-135 (edge-padded running mean along the last axis)."""
    data = np.asarray(data, dtype=float)
    pad = [(0, 0)] * (data.ndim - 1) + [(window - 1, 0)]
    data = np.pad(data, pad, mode="edge")
    ret = np.cumsum(data, axis=-1)
    ret[..., window:] = ret[..., window:] - ret[..., :-window]
    return ret[..., window - 1:] / window


def make_checkpoint_benchmark(split_rng, params, log_pdf, sample, losses, kde_kl_divergences, kde_hellinger_distances,
                              reconstruction_distances, n_model_sample=5000, save_dir="./results/benchmarks/", epoch=0, ngrid=300,
                              device="cuda"):
    """utils/helpers.py:170-214: density on the ngrid^2 unit-square mesh, model samples, KDE-based KL / Hellinger against the
    model density and the prior-space reconstruction distance; appends to the metric lists and rewrites the *.txt files."""
    from sklearn.neighbors import KernelDensity
    out = f"{save_dir}/outputs/"
    Path(out).mkdir(parents=True, exist_ok=True)
    g = np.linspace(0.0, 1.0, ngrid)
    xv, yv = np.meshgrid(g, g)
    grid = np.concatenate([xv.reshape(-1, 1), yv.reshape(-1, 1)], axis=-1)
    tg = torch.from_numpy(grid.astype(np.float32)).to(device)
    log_pdf_grid = log_pdf(params, tg).detach().cpu().numpy().astype(np.float64).reshape(ngrid, ngrid)
    pdf_grid = np.exp(log_pdf_grid)
    np.save(f"{out}/pdf_grid_epoch{epoch}.npy", pdf_grid)
    model_samples, original_samples = sample(split_rng, params, num_samples=n_model_sample, return_original_samples=True)
    ms = model_samples.detach().cpu().numpy()
    np.save(f"{out}/samples_epoch{epoch}.npy", ms)
    kde = KernelDensity(kernel="gaussian", bandwidth=0.01, rtol=0.1).fit(ms)
    log_pdf_grid_kde = kde.score_samples(grid).reshape(ngrid, ngrid)
    np.save(f"{out}/kde_pdf_grid_epoch{epoch}.npy", np.exp(log_pdf_grid_kde))
    kde_kl_divergences.append(float((pdf_grid * (log_pdf_grid - log_pdf_grid_kde)).mean()))
    kde_hellinger_distances.append(float(((np.sqrt(pdf_grid) - np.sqrt(np.exp(log_pdf_grid_kde))) ** 2).mean()))
    _, rec = log_pdf(params, model_samples, return_sample=True)
    reconstruction_distances.append(float(torch.linalg.norm(original_samples - rec, dim=-1).mean()))
    np.savetxt(f"{save_dir}/losses.txt", np.asarray(losses, dtype=float))
    np.savetxt(f"{save_dir}/kl_divergences.txt", kde_kl_divergences)
    np.savetxt(f"{save_dir}/hellinger_divergences.txt", kde_hellinger_distances)
    np.savetxt(f"{save_dir}/reconstruction_distances.txt", reconstruction_distances)
