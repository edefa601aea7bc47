"""Device time of the tensor-core local-energy kernel over a scan of batch sizes (D = 4 workload): separates the fixed cost of a
launch from the per-round cost.  python tools/tc_scan.py [N ...]"""
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
Ns = [int(a) for a in sys.argv[1:]] or [20, 40, 2960, 5920, 8192, 8880, 11840, 16384, 32768, 65536]
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
sums = torch.zeros(4, dtype=torch.float64, device=dev)
for N in Ns:
    x = torch.from_numpy(wl["walkers"][:N].copy()).to(dev)
    for mode in ("tc", "simt"):
        for _ in range(3):
            _live.local_energy(spec, w, x, wl["protons"], want=(), sums=sums, mode=mode)
        ts, tf = [], []
        for rep in range(10):
            for fl, acc in ((False, ts), (True, tf)):
                if fl:
                    flush.fill_(1.0)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); _live.local_energy(spec, w, x, wl["protons"], want=(), sums=sums, mode=mode); b.record()
                torch.cuda.synchronize(); acc.append(a.elapsed_time(b))
        print(f"N={N:6d} {mode:4s}: warm L2 {np.median(ts)*1e3:8.1f} us   after L2 flush {np.median(tf)*1e3:8.1f} us", flush=True)
