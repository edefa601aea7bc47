"""Profiling driver: a few launches of the local-energy kernel on the BASELINE configs[3] workload (65 536 walkers, D = 4).
Usage: python tools/tc_prof.py [tc|simt] [N]"""
import sys
import numpy as np, torch
sys.path.insert(0, '.')
from bench import workload
from waveflow_b200 import _live, model_factory

mode = sys.argv[1] if len(sys.argv) > 1 else "tc"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
dev = torch.device('cuda:0')
wl = workload("vqmc_c4")
D = wl["D"]
init = model_factory.get_waveflow_model(D, base_spline_degree=6, i_spline_degree=6, n_prior_internal_knots=23, n_i_internal_knots=23,
                                        i_spline_reg=0.05, i_spline_reverse_fun_tol=1e-6, n_flow_layers=3, box_size=10.0,
                                        xu_coord_type="mean", cached_bases_root=None)
_, psi, _, _ = init(0, D)
spec = psi.wf_spec
w = _live.pack_params(spec, wl["params"][0], wl["params"][1], dev)
x = torch.from_numpy(wl["walkers"][:N].copy()).to(dev)
sums = torch.zeros(4, dtype=torch.float64, device=dev)
for _ in range(4):
    _live.local_energy(spec, w, x, wl["protons"], want=("psi", "eloc"), sums=sums, mode=mode)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    _live.local_energy(spec, w, x, wl["protons"], want=("psi", "eloc"), sums=sums, mode=mode)
b.record(); torch.cuda.synchronize()
print(f"{mode} N={N}: {a.elapsed_time(b) / 10:.4f} ms/launch")
