"""Development aid: GPU float32 error vs the float64 oracle, next to the numpy-float32 oracle's own error."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from oracle import fixtures as fx, live, rqs as orqs, laplacian as olap
from tests.util import spec_from_live
from waveflow_b200 import _live
from waveflow_b200.flows.neural_splines import unconstrained_RQS
dev = torch.device("cuda:0")

def stats(name, got, r64, r32, scale):
    eg = np.abs(got - r64) / (np.abs(r64) + scale); eo = np.abs(r32 - r64) / (np.abs(r64) + scale)
    q = lambda e: (np.median(e), np.quantile(e, 0.99), e.max())
    print(f"{name:28s} gpu med/p99/max = %.2e %.2e %.2e | np32 = %.2e %.2e %.2e" % (*q(eg), *q(eo)))

for D, coord in [(2, "mean"), (4, "mean"), (3, "first")]:
    m = fx.waveflow_model(D, coord=coord)
    params = fx.random_params(np.random.default_rng(D), m, scale=3.0)
    spec = spec_from_live(m); w = _live.pack_params(spec, params[0], params[1], dev)
    x = np.sort(np.random.default_rng(1).uniform(-10, 10, (4001, D)), -1).astype(np.float32)
    out = _live.forward(spec, w, torch.from_numpy(x).to(dev), want=("u", "logdet", "logpdf", "psi"))
    m32, p32 = m.cast(np.float32), fx.cast_params(params, np.float32)
    x64 = x.astype(np.float64)
    p64 = live.psi(m, params, x64); p32v = live.psi(m32, p32, x)
    stats(f"D{D}{coord} psi", out["psi"].cpu().numpy(), p64, p32v, np.abs(p64).max())
    l64 = live.log_pdf(m, params, x64); l32 = live.log_pdf(m32, p32, x)
    stats(f"D{D}{coord} logpdf", out["logpdf"].cpu().numpy(), l64, l32, 1.0)
    u64, d64 = live.flow_direct(m, params[0], x64); u32, d32 = live.flow_direct(m32, p32[0], x)
    stats(f"D{D}{coord} u", out["u"].cpu().numpy(), u64, u32, 1.0)
    stats(f"D{D}{coord} logdet", out["logdet"].cpu().numpy(), d64, d32, 1.0)
    xs = x[:400]
    le = _live.local_energy(spec, w, torch.from_numpy(xs).to(dev), np.zeros((D, 1)), want=("psi", "hpsi", "eloc", "grad", "lap"))
    ref = olap.local_energy_bundle(m, params, xs.astype(np.float64), np.zeros((D, 1)))
    for k in ("psi", "grad", "lap", "hpsi"):
        g = le[k].cpu().numpy(); r = ref[k]
        e = np.abs(g - r) / (np.abs(r) + np.abs(r).max())
        print(f"   LE {k:5s} med/p99/max = %.2e %.2e %.2e" % (np.median(e), np.quantile(e, .99), e.max()))

for K in (5, 8, 32, 64):
    rng = np.random.default_rng(K); N, B = 50000, 3.0
    x = rng.uniform(-3, 3, N).astype(np.float32)
    uw, uh = rng.standard_normal((2, N, K)).astype(np.float32); ud = rng.standard_normal((N, K - 1)).astype(np.float32)
    t = lambda a: torch.from_numpy(a).to(dev)
    d = lambda a: a.astype(np.float64)
    for inv in (False, True):
        o, l = unconstrained_RQS(t(x), t(uw), t(uh), t(ud), inverse=inv, tail_bound=B)
        o64, l64, b64 = orqs.unconstrained_rqs(d(x), d(uw), d(uh), d(ud), inv, B, True)
        o32, l32, b32 = orqs.unconstrained_rqs(x, uw, uh, ud, inv, B, True)
        ok = b64 == b32
        stats(f"rqs K{K} inv{int(inv)} out", o.cpu().numpy()[ok], o64[ok], o32[ok], B)
        stats(f"rqs K{K} inv{int(inv)} lad", l.cpu().numpy()[ok], l64[ok], l32[ok], 1.0)
