"""Dev helper: folded vs unfolded prior layer on the bench workload -- per-walker agreement and the outliers of the mean."""
import sys
import numpy as np, torch
sys.path.insert(0, '.')
import bench
from oracle import fixtures as fx
from tests.util import spec_from_live
from waveflow_b200 import _live
dev = torch.device('cuda')
wl = bench.workload('vqmc_c4')
m = fx.waveflow_model(4)
spec = spec_from_live(m)
x = torch.from_numpy(wl['walkers']).to(dev)
res = {}
for fold in (False, True):
    w = _live.pack_params(spec, wl['params'][0], wl['params'][1], dev, fold_prior=fold)
    res[fold] = {k: v.double().cpu().numpy() for k, v in _live.local_energy(spec, w, x, wl['protons'], want=('psi', 'hpsi', 'eloc')).items()}
a, b = res[False], res[True]
for k in ('psi', 'hpsi', 'eloc'):
    rel = np.abs(a[k] - b[k]) / (np.abs(a[k]) + 1e-30)
    print(k, 'median rel diff %.2e  p99 %.2e  max %.2e' % (np.median(rel), np.quantile(rel, 0.99), rel.max()))
print('mean eloc', a['eloc'].mean(), b['eloc'].mean())
idx = np.argsort(-np.abs(a['eloc']))[:5]
print('largest |eloc| walkers:', [(int(i), float(a['eloc'][i]), float(b['eloc'][i]), float(a['psi'][i])) for i in idx])
keep = np.abs(a['eloc']) < 1e4
print('mean over |eloc| < 1e4:', a['eloc'][keep].mean(), b['eloc'][keep].mean(), 'dropped', int((~keep).sum()))
