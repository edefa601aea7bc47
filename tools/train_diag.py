"""Dev helper: per-layer error growth of the training path (value lanes) vs the float64 / float32 numpy oracle."""
import sys
import numpy as np, torch
sys.path.insert(0, '.')
from oracle import fixtures as fx, live
from tests.util import spec_from_live
from waveflow_b200 import _train, _ffi
import ctypes as C
cuda = torch.device('cuda')
D = 2
m = fx.waveflow_model(D)
params, _ = fx.load_he_checkpoint()
spec = spec_from_live(m)
rng = np.random.default_rng(12)
N = 2000
x = np.sort(rng.uniform(-4, 4, (N, D)), -1).astype(np.float32)
protons = np.zeros((D, 1))
flat = _train.ravel(fx.cast_params(params, np.float32), cuda)
g, out = _train.loss_grad(spec, flat, torch.from_numpy(x).to(cuda), protons, -1.8, want=('psi',), max_chunk=N)
torch.cuda.synchronize()
ws = list(_train._WS.values())[0]
G = D + 2
R = N * G
nn = 4
fixed = int(_ffi.lib.wf_vqmc_grad_workspace_floats(C.byref(spec.struct()), 0)) - 512      # masked weights | folded prior | partials
U = [ws[fixed + i * R * D: fixed + (i + 1) * R * D].view(N, G, D)[:, 0, :].cpu().numpy() for i in range(nn)]
p64 = fx.cast_params(params, np.float64); p32 = fx.cast_params(params, np.float32)
m32 = m.cast(np.float32)
def layers(mm, pp, xx):
    u, _ = live.box_direct(xx, mm.box, mm.coord)
    outs = [u]
    for net in [p for p in pp[0] if len(p)]:
        u, _ = live.imade_direct(mm, net, u)
        u = np.ascontiguousarray(u[:, ::-1])
        outs.append(u)
    return outs
o64 = layers(m, p64, x.astype(np.float64)); o32 = layers(m32, p32, x)
for i in range(nn):
    eg = np.abs(U[i] - o64[i]); eo = np.abs(o32[i] - o64[i])
    print('layer', i, 'gpu abs err med %.2e max %.2e | np32 med %.2e max %.2e' % (np.median(eg), eg.max(), np.median(eo), eo.max()))
# single layer from exact fp64 input (rounded to f32): isolates one layer's own error
for i in range(1, nn):
    pass
DPm = D * 29
off = fixed + (nn + 1) * R * D + nn * (4 * 64 * R + R * DPm)
LDbox = ws[off: off + R].view(N, G)[:, 0].cpu().numpy().astype(np.float64); off += R
LDC = ws[off: off + 3 * R * D].view(3, N, G, D)[:, :, 0, :].cpu().numpy().astype(np.float64); off += 3 * R * D
PHI = ws[off: off + R * D].view(N, G, D)[:, 0, :].cpu().numpy()
ld_gpu = LDbox + LDC.sum((0, 2))
u64, ld64 = live.flow_direct(m, p64[0], x.astype(np.float64)); u32, ld32 = live.flow_direct(m32, p32[0], x)
print('ld: gpu(f64 sum of f32 terms) abs err med %.2e max %.2e | np32 med %.2e max %.2e | |ld| med %.2f' % (
    np.median(np.abs(ld_gpu - ld64)), np.abs(ld_gpu - ld64).max(), np.median(np.abs(ld32 - ld64)), np.abs(ld32 - ld64).max(), np.median(np.abs(ld64))))
f64 = live.prior_factors(m, p64[1], u64); f32 = live.prior_factors(m32, p32[1], u32)
cons = live._constrained(m)
f64[:, cons] /= np.sqrt(2); f32[:, cons] /= np.sqrt(np.float32(2))
eg = np.abs(PHI - f64) / np.abs(f64).max(); eo = np.abs(f32 - f64) / np.abs(f64).max()
print('phi: gpu err med %.2e max %.2e | np32 med %.2e max %.2e' % (np.median(eg), eg.max(), np.median(eo), eo.max()))
# phi from the oracle's own float32 u (isolates the prior head)
psi64 = live.psi(m, p64, x.astype(np.float64)); psi32 = live.psi(m32, p32, x)
pg = out['psi'].cpu().numpy()
print('psi rel: gpu med %.2e | np32 med %.2e' % (np.median(np.abs(pg - psi64) / np.abs(psi64)), np.median(np.abs(psi32 - psi64) / np.abs(psi64))))
pg2 = PHI.astype(np.float64).prod(-1) * np.exp(0.5 * ld_gpu)
print('psi from gpu phi and f64-summed ld: med %.2e' % np.median(np.abs(pg2 - psi64) / np.abs(psi64)))
