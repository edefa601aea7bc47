"""Dev helper: GPU gradient vs oracle, per-leaf errors."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from oracle import fixtures as fx, grad as ograd
from tests.util import spec_from_live
from waveflow_b200 import _train
from tests.util import to_torch_tree
to_t = lambda p: to_torch_tree(fx.cast_params(p, np.float32), torch.device('cuda'))
cuda = torch.device('cuda')
D = int(sys.argv[1]) if len(sys.argv) > 1 else 2
coord = sys.argv[2] if len(sys.argv) > 2 else 'mean'
m = fx.waveflow_model(D, coord=coord)
rng = np.random.default_rng(10 + D)
if D == 2 and coord == 'mean':
    params, _ = fx.load_he_checkpoint()
else:
    params = fx.random_params(rng, m)
spec = spec_from_live(m)
NW = int(sys.argv[3]) if len(sys.argv) > 3 else 24
x = np.sort(rng.uniform(-4, 4, (NW, D)), -1).astype(np.float32)
protons = np.zeros((D, 1))
t = time.time()
loss_ref, g_ref = ograd.loss_and_grad(m, fx.cast_params(params, np.float64), x.astype(np.float64), protons, -1.8)
print('oracle', time.time() - t, 's loss', loss_ref)
flat = _train.ravel(fx.cast_params(params, np.float32), cuda)
sums = torch.zeros(4, dtype=torch.float64, device=cuda)
g, out = _train.loss_grad(spec, flat, torch.from_numpy(x).to(cuda), protons, -1.8, want=('psi', 'hpsi', 'eloc'), sums=sums)
torch.cuda.synchronize()
print('gpu loss', sums.cpu().numpy()[0] / NW)
from oracle import laplacian as olap
ref = olap.local_energy_bundle(m, fx.cast_params(params, np.float64), x.astype(np.float64), protons)
l32, g32 = ograd.loss_and_grad(m, fx.cast_params(params, np.float32), x, protons, -1.8, dtype=np.float32)
print('fp32 oracle loss', l32)
for k in ('psi', 'hpsi', 'eloc'):
    e = np.abs(out[k].cpu().numpy() - ref[k]) / (np.abs(ref[k]) + 1e-30)
    print(k, 'rel err max %.2e med %.2e' % (e.max(), np.median(e)), 'argmax', e.argmax(), 'ref', ref[k][e.argmax()], 'psi there', ref['psi'][e.argmax()])
g32l = _train.tree_leaves(g32)
from waveflow_b200 import _live
w = _live.pack_params(spec, to_t(params)[0], to_t(params)[1], cuda)
lo = _live.local_energy(spec, w, torch.from_numpy(x).to(cuda), protons, want=('psi', 'hpsi', 'eloc'))
for k in ('psi', 'hpsi', 'eloc'):
    e = np.abs(lo[k].cpu().numpy() - ref[k]) / (np.abs(ref[k]) + 1e-30)
    print('live kernel', k, 'rel err max %.2e med %.2e' % (e.max(), np.median(e)))
r32 = olap.local_energy_bundle(m.cast(np.float32), fx.cast_params(params, np.float32), x, protons)
for k in ('psi', 'hpsi', 'eloc'):
    e = np.abs(r32[k] - ref[k]) / (np.abs(ref[k]) + 1e-30)
    print('np32 oracle', k, 'rel err max %.2e med %.2e' % (e.max(), np.median(e)))
gt = _train.unravel(params, g)
for a, b in zip(_train.tree_leaves(gt), _train.tree_leaves(g_ref)):
    a = a.cpu().numpy().astype(np.float64); b = np.asarray(b)
    nb = np.linalg.norm(b)
    c = np.asarray(g32l.pop(0), dtype=np.float64)
    print(a.shape, 'ref norm %.3e' % nb, 'rel err %.3e' % (np.linalg.norm(a - b) / (nb + 1e-300)), 'fp32 oracle %.3e' % (np.linalg.norm(c - b) / (nb + 1e-300)))
