"""Runs the tensor-core local-energy kernel on one batch size and compares with the CUDA-core kernel (debug aid).
python tools/tc_probe.py N [D]"""
import sys
import numpy as np, torch
sys.path.insert(0, '.')
from bench import workload
from waveflow_b200 import _live, model_factory
N = int(sys.argv[1]); 
dev = torch.device('cuda:0')
wl = workload("vqmc_c4" if len(sys.argv) < 3 or sys.argv[2] == "4" else "vqmc_c2")
D = wl["D"]
init = model_factory.get_waveflow_model(D, base_spline_degree=6, i_spline_degree=6, n_prior_internal_knots=23, n_i_internal_knots=23,
                                        i_spline_reg=0.05, i_spline_reverse_fun_tol=1e-6, n_flow_layers=3, box_size=10.0,
                                        xu_coord_type="mean", cached_bases_root=None)
_, psi, _, _ = init(0, D)
spec = psi.wf_spec
w = _live.pack_params(spec, wl["params"][0], wl["params"][1], dev)
rng = np.random.Generator(np.random.PCG64(1))
x = torch.from_numpy(np.sort(rng.uniform(-10, 10, (N, D)), -1).astype(np.float32)).to(dev)
a = _live.local_energy(spec, w, x, wl["protons"], want=("psi", "eloc"), mode="tc"); torch.cuda.synchronize()
b = _live.local_energy(spec, w, x, wl["protons"], want=("psi", "eloc"), mode="simt"); torch.cuda.synchronize()
print(f"N={N} D={D}: ok, max|dpsi|/max|psi| = {float((a['psi'] - b['psi']).abs().max() / b['psi'].abs().max()):.2e}", flush=True)
