"""Bitwise comparison of the local energies of the bench workload evaluated in one call and in 8 contiguous shards."""
import sys
import numpy as np, torch
sys.path.insert(0, '.')
from bench import workload
from waveflow_b200 import _live, model_factory
dev = torch.device('cuda:0')
wl = workload("vqmc_c4")
D = wl["D"]
init = model_factory.get_waveflow_model(D, base_spline_degree=6, i_spline_degree=6, n_prior_internal_knots=23, n_i_internal_knots=23,
                                        i_spline_reg=0.05, i_spline_reverse_fun_tol=1e-6, n_flow_layers=3, box_size=10.0,
                                        xu_coord_type="mean", cached_bases_root=None)
_, psi, _, _ = init(0, D)
spec = psi.wf_spec
w = _live.pack_params(spec, wl["params"][0], wl["params"][1], dev)
x = torch.from_numpy(wl["walkers"]).to(dev)
for mode in ("tc", "simt"):
    fs = torch.zeros(4, dtype=torch.float64, device=dev)
    full = _live.local_energy(spec, w, x, wl["protons"], want=("psi", "hpsi", "eloc"), sums=fs, mode=mode)
    ps = torch.zeros(4, dtype=torch.float64, device=dev)
    parts = [_live.local_energy(spec, w, x[r * 8192:(r + 1) * 8192].contiguous(), wl["protons"], want=("psi", "hpsi", "eloc"), sums=ps, mode=mode)
             for r in range(8)]
    torch.cuda.synchronize()
    for k in ("psi", "hpsi", "eloc"):
        a, b = full[k], torch.cat([p[k] for p in parts])
        bad = (a != b).nonzero().flatten()
        print(mode, k, "differing walkers:", bad.numel(), bad[:8].tolist(), [(float(a[i]), float(b[i])) for i in bad[:3]])
    print(mode, "sums full", fs.tolist(), "parts", ps.tolist(), "mean", float(fs[0] / fs[2]), float(ps[0] / ps[2]))
    again = _live.local_energy(spec, w, x, wl["protons"], want=("eloc",), mode=mode)
    print(mode, "run-to-run differing:", int((again["eloc"] != full["eloc"]).sum()))
