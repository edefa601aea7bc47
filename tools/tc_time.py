"""Times wf_local_energy / wf_live_forward on the CUDA-core and the tensor-core path (same inputs), prints ms and the max
difference.  Usage: python tools/tc_time.py [N ...]"""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from bench import workload
from waveflow_b200 import _live, model_factory
from waveflow_b200.utils import physics

dev = torch.device('cuda:0')
for wname, Ns in (("vqmc_c4", [65536, 8192, 2048]), ("vqmc_c2", [65536, 256])):
    wl = workload(wname)
    D = wl["D"]
    init = model_factory.get_waveflow_model(D, base_spline_degree=6, i_spline_degree=6, n_prior_internal_knots=23, n_i_internal_knots=23,
                                            i_spline_reg=0.05, i_spline_reverse_fun_tol=1e-6, n_flow_layers=3, box_size=10.0,
                                            xu_coord_type="mean", cached_bases_root=None)
    _, psi, _, _ = init(0, D)
    spec = psi.wf_spec
    w = _live.pack_params(spec, wl["params"][0], wl["params"][1], dev)
    rng = np.random.Generator(np.random.PCG64(1))
    for N in Ns:
        x = torch.from_numpy(np.sort(rng.uniform(-10, 10, (N, D)), -1).astype(np.float32)).to(dev)
        res = {}
        for mode in ("simt", "tc"):
            for lap in (True, False):
                f = (lambda: _live.local_energy(spec, w, x, wl["protons"], want=("psi", "eloc"), mode=mode)) if lap else \
                    (lambda: _live.forward(spec, w, x, want=("psi", "logdet"), mode=mode))
                for _ in range(3):
                    out = f()
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(10):
                    out = f()
                b.record(); torch.cuda.synchronize()
                res[(mode, lap)] = (a.elapsed_time(b) / 10, out)
        for lap in (True, False):
            ts, os_ = res[("simt", lap)]; tt, ot = res[("tc", lap)]
            dpsi = float((os_["psi"] - ot["psi"]).abs().max() / os_["psi"].abs().max())
            print(f"{wname} D={D} N={N:6d} lap={int(lap)}: simt {ts:8.4f} ms   tc {tt:8.4f} ms   speed-up {ts / tt:5.2f}   max|dpsi|/max|psi| {dpsi:.2e}", flush=True)
