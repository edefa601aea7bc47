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


def _grad_worker(rank, world, port, n_total, out_path):
    """Sharded training step, host side: per-rank shard gradients (float64 oracle here, wf_vqmc_loss_grad on the GPUs) with the
    GLOBAL 1/N, merged by vqmc.total_walkers / vqmc.reduce_loss_and_grad -- must equal the single-process gradient."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import fixtures as fx
    from oracle import grad as ograd
    from waveflow_b200 import _train, vqmc
    torch.set_num_threads(1)
    params, _ = fx.load_he_checkpoint()
    p64 = fx.cast_params(params, np.float64)
    m = fx.waveflow_model(2)
    prot = np.array([[0.0], [0.0]])
    x = np.sort(np.random.default_rng(0).uniform(-3, 3, (n_total, 2)), -1)
    lo, hi = vqmc.EnergyEstimator.shard(n_total, rank, world)
    n_all = vqmc.total_walkers(hi - lo, torch.device("cpu"))
    loss_l, g_l = ograd.loss_and_grad(m, p64, x[lo:hi], prot, -1.8)            # mean over the SHARD ...
    flat = torch.cat([torch.as_tensor(np.asarray(a, dtype=np.float64)).reshape(-1) for a in _train.tree_leaves(g_l)])
    flat = flat * (hi - lo) / n_all                                              # ... rescaled to the global 1/N
    sums = torch.tensor([loss_l * (hi - lo), 0.0, float(hi - lo), 0.0], dtype=torch.float64)
    loss = vqmc.reduce_loss_and_grad(flat, sums, n_all)
    if rank == 0:
        loss_f, g_f = ograd.loss_and_grad(m, p64, x, prot, -1.8)
        full = torch.cat([torch.as_tensor(np.asarray(a, dtype=np.float64)).reshape(-1) for a in _train.tree_leaves(g_f)])
        np.save(out_path, np.array([n_all, float(loss), loss_f, float((flat - full).abs().max()), float(full.abs().max())]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_training_exchange_equals_single_process(tmp_path):
    out = str(tmp_path / "g.npy")
    mp.spawn(_grad_worker, args=(2, _free_port(), 11, out), nprocs=2, join=True)
    n_all, loss, loss_f, err, scale = np.load(out)
    assert n_all == 11
    assert abs(loss - loss_f) <= 1e-6 * abs(loss_f)
    assert err <= 1e-9 * scale
