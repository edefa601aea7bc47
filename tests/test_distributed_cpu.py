"""N > 1 host logic on CPU (gloo, world_size 2): walker sharding + the 32-byte estimator all-reduce.

The per-rank local sums come from the CPU oracle here (no GPU in this test); on the GPUs the same
EnergyEstimator.shard / EnergyEstimator.reduce wrap wf_local_energy (tests/test_gpu_api.py, bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import fast_cpu
    from oracle import fixtures as fx
    from waveflow_b200.vqmc import EnergyEstimator
    torch.set_num_threads(1)
    params, _ = fx.load_he_checkpoint()
    m = fx.waveflow_model(2, dtype=np.float32)
    f = fast_cpu.FastLocalEnergy(m, params, np.array([[0.0], [0.0]]), dtype=torch.float64)
    x = np.sort(np.random.default_rng(0).uniform(-10, 10, (n_total, 2)), -1)
    lo, hi = EnergyEstimator.shard(n_total, rank, world)
    r = f(x[lo:hi])
    e = r["eloc"].astype(np.float64)
    sums = torch.tensor([e.sum(), (e * e).sum(), float(hi - lo), (r["psi"].astype(np.float64) ** 2).sum()], dtype=torch.float64)
    est = EnergyEstimator.reduce(sums)
    if rank == 0:
        full = f(x)["eloc"].astype(np.float64)
        np.save(out_path, np.array([est["energy"], est["variance"], est["n"], full.mean(), full.var(), len(full)]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_covers_every_walker_once():
    from waveflow_b200.vqmc import EnergyEstimator
    for n, w in [(65536, 8), (10, 4), (7, 8), (1, 2), (256, 1)]:
        blocks = [EnergyEstimator.shard(n, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(blocks[:-1], blocks[1:]))
        sizes = [hi - lo for lo, hi in blocks]
        assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(300)
def test_two_rank_estimator_equals_single_process(tmp_path):
    out = str(tmp_path / "res.npy")
    mp.spawn(_worker, args=(2, _free_port(), 101, out), nprocs=2, join=True)
    e, v, n, e1, v1, n1 = np.load(out)
    assert n == n1 == 101
    assert abs(e - e1) <= 1e-10 * abs(e1) and abs(v - v1) <= 1e-8 * abs(v1)
