import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from oracle import rqs as orqs
from waveflow_b200.flows.neural_splines import coupling_flow_tc, pack_fcnn_tc
dev = torch.device("cuda:0")
D, K, H, B = 64, 64, 512, 3.0
L = int(sys.argv[1]) if len(sys.argv) > 1 else 2
rng = np.random.default_rng(0)
out = (3 * K - 1) * D // 2
layers = [(orqs.random_fcnn(rng, D // 2, H, out), orqs.random_fcnn(rng, D // 2, H, out)) for _ in range(L)]
t = lambda a: torch.from_numpy(a).to(dev)
stax = lambda f: [(t(f[0][0]), t(f[0][1])), (), (t(f[1][0]), t(f[1][1])), (), (t(f[2][0]), t(f[2][1]))]
w = torch.cat([pack_fcnn_tc(stax(f), dev) for pair in layers for f in pair]).contiguous()
N = 1500
x = rng.uniform(-3.3, 3.3, (N, D)).astype(np.float32)
y, ld = coupling_flow_tc(w, L, t(x), B)
torch.cuda.synchronize()
l64 = [tuple([(W.astype(np.float64), b.astype(np.float64)) for W, b in f] for f in pair) for pair in layers]
ry, rld = orqs.coupling_flow_direct(l64, x.astype(np.float64), K, B)
r32y, r32ld = orqs.coupling_flow_direct(layers, x, K, B)
ok = np.abs(r32y - ry).max(-1) < 1e-3
def stats(name, g, r64, r32, scale):
    eg = np.abs(g - r64) / (np.abs(r64) + scale); eo = np.abs(r32 - r64) / (np.abs(r64) + scale)
    print(f"{name}: gpu med/p99/max {np.median(eg):.2e} {np.quantile(eg,.99):.2e} {eg.max():.2e} | np32 {np.median(eo):.2e} {np.quantile(eo,.99):.2e} {eo.max():.2e}")
print("path-consistent fraction", ok.mean())
stats("y", y.cpu().numpy()[ok], ry[ok], r32y[ok], B)
stats("logdet", ld.cpu().numpy()[ok], rld[ok], r32ld[ok], 1.0)
xr, ldi = coupling_flow_tc(w, L, y, B, inverse=True)
print("round trip max", (xr.cpu().numpy() - x).__abs__().max(), "logdet cancel med", np.median(np.abs((ld + ldi).cpu().numpy())))
# timing
N = 1 << 17
xb = torch.rand(N, D, device=dev) * 6 - 3
for _ in range(2): coupling_flow_tc(w, L, xb, B)
torch.cuda.synchronize(); t0 = time.perf_counter()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); coupling_flow_tc(w, L, xb, B); b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b)
fl = L * 2 * 2 * (32 * 512 + 512 * 512 + 512 * 6112) * N
print(f"N={N} L={L}: {ms:.2f} ms, {N/ms*1e3/1e6:.2f} M samples/s, {fl/ms/1e9:.1f} TFLOP/s useful (algorithmic), per-layer-pair {ms/L:.2f} ms")
