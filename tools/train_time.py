"""Dev helper: time wf_vqmc_loss_grad (forward + backward) and per-kernel shares."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from oracle import fixtures as fx
from tests.util import spec_from_live
from waveflow_b200 import _train
cuda = torch.device('cuda')
D = int(sys.argv[1]) if len(sys.argv) > 1 else 4
N = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
m = fx.waveflow_model(D)
rng = np.random.default_rng(0)
params = fx.random_params(rng, m)
spec = spec_from_live(m)
x = torch.from_numpy(np.sort(rng.uniform(-10, 10, (N, D)), -1).astype(np.float32)).to(cuda)
flat = _train.ravel(fx.cast_params(params, np.float32), cuda)
prot = np.zeros((D, 1))
g = torch.zeros_like(flat)
for with_grad in (True, False):
    for _ in range(3):
        _train.loss_grad(spec, flat, x, prot, 0.0, grad=g, with_grad=with_grad, max_chunk=chunk)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    K = 5
    for _ in range(K):
        _train.loss_grad(spec, flat, x, prot, 0.0, grad=g, with_grad=with_grad, max_chunk=chunk)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print(f'D={D} N={N} chunk={chunk} with_grad={with_grad}: {ms:.3f} ms  {N / ms * 1e3 / 1e6:.2f} M walkers/s')
if N <= 4096:
    from waveflow_b200 import vqmc
    opt_init, opt_update, get_params = _train.adam(1e-4)
    st = opt_init(fx.cast_params(params, np.float32))
    class H: pass
    h = H(); h.wf_spec = spec; h.protons = np.zeros(D, dtype=np.float32)
    for ug in (False, True):
        for i in range(5):
            vqmc.train_step_efficient(i, None, h, opt_update, st, get_params(st), x, 0.0, use_graph=ug)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        for i in range(50):
            vqmc.train_step_efficient(5 + i, None, h, opt_update, st, get_params(st), x, 0.0, use_graph=ug)
        e1.record(); torch.cuda.synchronize()
        print(f'train_step_efficient N={N} use_graph={ug}: {e0.elapsed_time(e1) / 50:.3f} ms/step (wall {(time.perf_counter() - t0) / 50 * 1e3:.3f} ms)')
