"""Rough per-kernel timings (CUDA events) -- development aid, not the bench contract."""
import sys, time, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from waveflow_b200 import _live, _ffi
from waveflow_b200.splines.tables import SplineTables
from waveflow_b200.splines.factories import spline_apply
from waveflow_b200.flows.neural_splines import unconstrained_RQS

dev = torch.device("cuda:0")
res = {}
ONLY = sys.argv[1] if len(sys.argv) > 1 else "all"

def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))

# 1. table spline operator
tabs = SplineTables.get("I", 6, 23)
for logM in ((22, 24) if ONLY in ("all", "spline") else ()):
    M = 1 << logM
    c = torch.rand(M, 29, device=dev); c = c / c.sum(-1, keepdim=True)
    x = torch.rand(M, device=dev)
    med, mn = timeit(lambda: spline_apply(tabs, c, x, 0, 2, logd=True))
    res[f"spline_local_M2^{logM}"] = dict(ms=med, min_ms=mn, GBs=M * 128 / med / 1e6, elems_per_s=M / med * 1e3)
    if logM == 22:
        med, mn = timeit(lambda: spline_apply(tabs, c, x, 0, 2, logd=True, force_dense=True))
        res[f"spline_dense_M2^{logM}"] = dict(ms=med, GBs=M * 128 / med / 1e6)
    del c, x
# 2. rqs operator
for K in ((32, 64) if ONLY in ("all", "rqs") else ()):
    M = 1 << 22
    uw, uh = torch.randn(M, K, device=dev), torch.randn(M, K, device=dev); ud = torch.randn(M, K - 1, device=dev)
    x = (torch.rand(M, device=dev) * 6 - 3)
    for inv in (False, True):
        med, mn = timeit(lambda: unconstrained_RQS(x, uw, uh, ud, inverse=inv, tail_bound=3.0))
        res[f"rqs_K{K}_inv{int(inv)}"] = dict(ms=med, min_ms=mn, GBs=M * 4 * (3 * K + 2) / med / 1e6, elems_per_s=M / med * 1e3)
    del uw, uh, ud, x
# 2b. fused coupling flow (BASELINE config 3 ii)
if ONLY in ("all", "coupling"):
    from waveflow_b200.flows.neural_splines import coupling_flow
    from oracle import rqs as orqs
    for D, hidden in ((2, 8), (8, 8), (2, 64), (8, 64)):
        rng = np.random.default_rng(0); K = 32
        out = (3 * K - 1) * D // 2
        t = lambda a: torch.from_numpy(a).to(dev)
        layers = []
        for _ in range(8):
            pair = []
            for _f in range(2):
                (W1, b1), (W2, b2), (W3, b3) = orqs.random_fcnn(rng, D // 2, hidden, out)
                pair.append([(t(W1), t(b1)), (), (t(W2), t(b2)), (), (t(W3), t(b3))])
            layers.append(tuple(pair))
        N = 1 << 22
        x = torch.rand(N, D, device=dev) * 6 - 3
        for inv in (False, True):
            med, mn = timeit(lambda: coupling_flow(layers, x, K, 3.0, hidden, inverse=inv), n=5)
            res[f"coupling_D{D}_h{hidden}_inv{int(inv)}_N2^22"] = dict(ms=med, samples_per_s=N / med * 1e3)
# 3. fused live
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle import fixtures as fx
from tests.util import spec_from_live
for D in ((2, 4) if ONLY in ("all", "live") else ()):
    m = fx.waveflow_model(D)
    params = fx.random_params(np.random.default_rng(0), m)
    spec = spec_from_live(m)
    w = _live.pack_params(spec, params[0], params[1], dev)
    for N in (256, 8192, 65536, 1 << 20):
        x = torch.sort(torch.rand(N, D, device=dev) * 20 - 10, dim=-1).values.contiguous()
        med, mn = timeit(lambda: _live.forward(spec, w, x, want=("psi", "logpdf")))
        res[f"live_fwd_D{D}_N{N}"] = dict(ms=med, min_ms=mn, samples_per_s=N / med * 1e3)
        if N <= 65536 * 4:
            sums = torch.zeros(4, dtype=torch.float64, device=dev)
            med, mn = timeit(lambda: _live.local_energy(spec, w, x, np.zeros((D, 1)), want=("eloc",), sums=sums))
            res[f"local_energy_D{D}_N{N}"] = dict(ms=med, min_ms=mn, walkers_per_s=N / med * 1e3)
for k, v in res.items():
    print(k, json.dumps(v))
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/quick_time.json").write_text(json.dumps(res, indent=1))
