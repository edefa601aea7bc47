"""Profiling driver for the operator kernels: python tools/op_prof.py rqs32|rqs64|coupling2|coupling8|spline [log2 elements]"""
import sys
import numpy as np, torch
sys.path.insert(0, '.')
what = sys.argv[1]
lg = int(sys.argv[2]) if len(sys.argv) > 2 else 24
dev = torch.device('cuda:0')
g = torch.Generator(device=dev); g.manual_seed(0)
M = 1 << lg
if what.startswith("rqs"):
    from waveflow_b200.flows.neural_splines import unconstrained_RQS
    K = int(what[3:])
    uw = torch.randn(M, K, device=dev, generator=g); uh = torch.randn(M, K, device=dev, generator=g)
    ud = torch.randn(M, K - 1, device=dev, generator=g); xs = torch.rand(M, device=dev, generator=g) * 6 - 3
    f = lambda: unconstrained_RQS(xs, uw, uh, ud, inverse=False, tail_bound=3.0)
    nbytes = M * 4 * (3 * K + 2)
elif what.startswith("coupling"):
    from waveflow_b200.flows.neural_splines import coupling_flow
    Dc, K, hidden, L = int(what[8:]), 32, 8, 8
    rng = np.random.Generator(np.random.PCG64(0))
    od = (3 * K - 1) * Dc // 2
    gW = lambda a, b: torch.from_numpy((rng.standard_normal((a, b)) / np.sqrt(a)).astype(np.float32)).to(dev)
    z = lambda n: torch.zeros(n, device=dev)
    layers = [tuple([(gW(Dc // 2, hidden), z(hidden)), (), (gW(hidden, hidden), z(hidden)), (), (gW(hidden, od), z(od))] for _ in range(2))
              for _ in range(L)]
    x = torch.rand(M, Dc, device=dev, generator=g) * 6 - 3
    f = lambda: coupling_flow(layers, x, K, 3.0, hidden)
    nbytes = M * (8 * Dc + 4)
else:
    from waveflow_b200.splines.factories import spline_apply
    from waveflow_b200.splines.tables import SplineTables
    tabs = SplineTables.get("I", 6, 23)
    c = torch.rand(M, tabs.P, device=dev, generator=g); c /= c.sum(-1, keepdim=True)
    xs = torch.rand(M, device=dev, generator=g)
    f = lambda: spline_apply(tabs, c, xs, 0, 2, logd=True)
    nbytes = M * (4 * tabs.P + 12)
for _ in range(3):
    f()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    f()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
print(f"{what} 2^{lg}: {ms:.4f} ms  {nbytes / ms / 1e6:.1f} GB/s algorithmic")
