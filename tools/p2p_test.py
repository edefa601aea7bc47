"""torchrun helper: the peer-memory one-shot all-reduce against NCCL on random block sums, many steps (skewed ranks)."""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, '.')
rank, world, lr = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lr)
dev = torch.device('cuda', lr)
dist.init_process_group('nccl', device_id=dev)
from waveflow_b200 import vqmc
px = vqmc.PeerExchange(dev)
g = torch.Generator(device=dev); g.manual_seed(100 + rank)
bad = 0
bad_ranks = 0
for step in range(200):
    s = torch.randn(4, dtype=torch.float64, device=dev, generator=g) * 1e6
    if step % 7 == rank % 7:
        torch.cuda._sleep(2_000_000)                       # skew the ranks
    ref = s.clone(); dist.all_reduce(ref)
    out = px.all_reduce(s).clone()
    gathered = [torch.zeros_like(out) for _ in range(world)]
    dist.all_gather(gathered, out)
    same = all(torch.equal(gathered[0], t) for t in gathered)
    mags = s.abs().clone(); dist.all_reduce(mags)            # sum of |terms|: the scale of the rounding error of any order
    if not same:
        bad_ranks += 1
    if not bool(((out - ref).abs() <= 4e-16 * mags).all()):
        bad += 1
torch.cuda.synchronize()
# latency
for name, fn in (('p2p', lambda s: px.all_reduce(s)), ('nccl', lambda s: dist.all_reduce(s))):
    s = torch.ones(4, dtype=torch.float64, device=dev)
    for _ in range(20): fn(s)
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(200): fn(s)
    b.record(); torch.cuda.synchronize()
    if rank == 0: print(f'{name}: {a.elapsed_time(b) / 200 * 1e3:.1f} us per exchange', flush=True)
if rank == 0: print('steps where ranks disagree bitwise:', bad_ranks, ' steps off NCCL by more than 2 ulp of sum|terms|:', bad, flush=True)
dist.barrier(); dist.destroy_process_group()
