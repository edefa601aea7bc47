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
# ---- flat float32 vector (the training gradient): wf_p2p_allreduce_vec against NCCL, eager and replayed from a CUDA graph
n = 48280
gx = vqmc.GradExchange(n, dev)
vbad = vdis = 0
for step in range(100):
    v = torch.randn(n, device=dev, generator=g)
    s4 = torch.randn(4, dtype=torch.float64, device=dev, generator=g)
    if step % 5 == rank % 5:
        torch.cuda._sleep(1_000_000)
    ref = v.clone(); dist.all_reduce(ref)
    rs = s4.clone(); dist.all_reduce(rs)
    mag = v.abs().clone(); dist.all_reduce(mag)
    out = v.clone()
    so = gx.all_reduce(out, s4).clone()
    gathered = [torch.zeros_like(out) for _ in range(world)]
    dist.all_gather(gathered, out)
    vdis += 0 if all(torch.equal(gathered[0], t) for t in gathered) else 1
    vbad += 0 if bool(((out - ref).abs() <= 1e-6 * mag + 1e-30).all()) and bool(((so - rs).abs() <= 1e-12 * rs.abs() + 1e-12).all()) else 1
buf = torch.randn(n, device=dev, generator=g); s4 = torch.ones(4, dtype=torch.float64, device=dev)
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    gx.all_reduce(buf, s4)
torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize(); dist.barrier()
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    gx.all_reduce(buf, s4)
gbad = 0
for step in range(50):
    src = torch.randn(n, device=dev, generator=g)
    buf.copy_(src); ref = src.clone(); dist.all_reduce(ref)
    mag = src.abs().clone(); dist.all_reduce(mag)
    graph.replay()
    gbad += 0 if bool(((buf - ref).abs() <= 1e-6 * mag + 1e-30).all()) else 1
torch.cuda.synchronize()
for name, fn in (('p2p vec', lambda: gx.all_reduce(buf, s4)), ('nccl vec + sums', lambda: (dist.all_reduce(buf), dist.all_reduce(s4)))):
    for _ in range(10): fn()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(100): fn()
    b.record(); torch.cuda.synchronize()
    if rank == 0: print(f'{name}: {a.elapsed_time(b) / 100 * 1e3:.1f} us per exchange of {n} floats', flush=True)
if rank == 0: print('vector exchange: steps where ranks disagree bitwise:', vdis, ' off NCCL:', vbad, ' graph replays off NCCL:', gbad,
                    ' sticky error word:', gx.failed_step(), flush=True)
dist.barrier(); dist.destroy_process_group()
