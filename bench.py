#!/usr/bin/env python
"""bench.py -- VQMC local-energy throughput of the fused live path on B200 (+ the HBM-bound spline operator sweep).

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference ...                      (CPU restatement of the reference on the host cores)

Workload (BASELINE.json configs[3], the VQMC config the north star asks to scale 1 -> 8 GPUs): 1-D 4-electron box,
L = 10, `get_waveflow_model(4, degree 6, 23 knots, 3 flow layers, reg 0.05)`, 65 536 walkers sharded over the ranks
(STRONG scaling); one step = one local-energy pass (psi, H psi, E_loc, block sums) over the rank's walkers followed by the
32-byte estimator all-reduce.  `--workload vqmc_c2` runs BASELINE configs[1] (He, D = 2, published checkpoint, batch 256).
Synthetic data: walkers sorted U(-10, 10)^D (PCG64 seed 1), parameters U(+-1/sqrt(fan_in)) (PCG64 seed 0).

The JSON line also carries: `parity` (the CUDA results of the TIMED workloads checked against the oracle in the same run:
fraction of elements inside north_star's flat tolerances), `configs` (BASELINE configs[0] MFlow log-prob + sampling at
batch 256 and configs[1] He batch-256 local energy, GPU next to the CPU port), the HBM-bound operator sweeps with their
rooflines, the coupling-flow sweeps of configs[2] / [4] and the training step.  The reference arm (`--impl reference`)
times the CPU port on ALL walkers of the same workload (same `config`).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "vqmc_local_energy_walkers_per_s"
UNIT = "walkers/s"

# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures (profiles/)
TRAFFIC = {
    ("live_simt", 4, 65536): 2376960,        # profiles/r01_live_kernel_d4_lap_ncu_summary.txt (algorithmic: 24 B/walker = 1.57 MB)
    ("live_tc", 4, 65536): 2564096,          # profiles/r02_live_tc_kernel_d4_lap_ncu_summary.txt (reads only; 0 B written back through DRAM)
    ("rqs_staged", 32, 1 << 24, False): 6147839000 + 137697792,   # profiles/r02_rqs32_ncu_summary.txt (algorithmic: 392 B x 2^24 = 6.58 GB)
    ("spline_local", 29, 1 << 24): 2013555000 + 196376576,        # profiles/r02_spline_ncu_summary.txt (algorithmic: 128 B x 2^24 = 2.15 GB)
}


# ------------------------------------------------------------------------------------------------- workload definition
def net_params(rng, D, P, hidden=64):
    """One masked conditioner in the reference's pytree layout; W, b ~ U(+-1/sqrt(fan_in)) (model_factory.py:25-28)."""
    def u(shape, fan_in):
        b = 1.0 / np.sqrt(fan_in)
        return rng.uniform(-b, b, size=shape).astype(np.float32)
    nn = [(u((D, hidden), D), u((hidden,), D)), (), (u((hidden, hidden), hidden), u((hidden,), hidden)), (),
          (u((hidden, D * P), hidden), u((D * P,), hidden))]
    return (nn, rng.uniform(-0.5, 0.5, size=(D, P)).astype(np.float32))


def workload(name: str):
    if name == "vqmc_c4":
        D, n_walkers, protons = 4, 65536, np.zeros((4, 1))
        rng = np.random.Generator(np.random.PCG64(0))
        tp = [()]
        for _ in range(3):
            tp += [net_params(rng, D, 29), ()]
        params = (tp, net_params(rng, D, 28))
        desc = "BASELINE configs[3]: 1-D 4-electron box L=10, degree 6 / 23 knots / 3 IMADE layers, 65536 walkers (strong scaling)"
    elif name == "vqmc_c2":
        D, n_walkers, protons = 2, 256, np.array([[0.0], [0.0]])
        z = np.load(ROOT / "tests" / "golden" / "he_checkpoint_epoch100000.npz")
        names = ["W1", "b1", "W2", "b2", "W3", "b3", "zero"]

        def net(prefix):
            W1, b1, W2, b2, W3, b3, zero = [z[f"{prefix}_{n}"] for n in names]
            return ([(W1, b1), (), (W2, b2), (), (W3, b3)], zero)
        tp = [()]
        for li in range(3):
            tp += [net(f"imade{li}"), ()]
        params = (tp, net("prior"))
        desc = "BASELINE configs[1]: He 1-D, L=10, published checkpoint (epoch 100000), batch 256"
    else:
        raise SystemExit(f"unknown workload {name}")
    wrng = np.random.Generator(np.random.PCG64(1))
    walkers = np.sort(wrng.uniform(-10.0, 10.0, size=(n_walkers, D)), axis=-1).astype(np.float32)
    return dict(name=name, D=D, n_walkers=n_walkers, protons=protons, params=params, walkers=walkers, desc=desc,
                degree=6, knots=23, layers=3, box=10.0, reg=0.05)


def common_config(wl):
    """The part of `config` both arms (B200 and the CPU reference arm) share verbatim: what is computed, on what."""
    return {"workload": wl["name"], "description": wl["desc"], "walkers_total": wl["n_walkers"],
            "walkers_per_step": wl["n_walkers"], "D": wl["D"], "precision": "float32"}


def rel_stats(got, ref, tol, ref32=None, scale=None):
    """Parity record of one output: pointwise relative error |got - ref| / (|ref| + scale) (scale defaults to 0: north_star's
    flat relative tolerance; the ill-conditioned outputs also get the figure relative to the batch maximum), and the same
    fraction for the reference's own float32 arithmetic when `ref32` is given."""
    got, ref = np.asarray(got, dtype=np.float64).reshape(-1), np.asarray(ref, dtype=np.float64).reshape(-1)
    e = np.abs(got - ref) / (np.abs(ref) + (1e-300 if scale is None else scale))
    out = {"n": int(e.size), "tol": tol, "frac_within_tol": float(np.mean(e <= tol)), "median_rel": float(np.median(e)),
           "p99_rel": float(np.quantile(e, 0.99)), "max_rel": float(e.max()),
           "max_abs_over_batch_max": float(np.abs(got - ref).max() / (np.abs(ref).max() + 1e-300))}
    if ref32 is not None:
        r32 = np.asarray(ref32, dtype=np.float64).reshape(-1)
        e32 = np.abs(r32 - ref) / (np.abs(ref) + (1e-300 if scale is None else scale))
        out["frac_within_tol_float32_restatement"] = float(np.mean(e32 <= tol))
        out["max_rel_float32_restatement"] = float(e32.max())
        out["max_abs_over_batch_max_float32_restatement"] = float(np.abs(r32 - ref).max() / (np.abs(ref).max() + 1e-300))
    return out


def flops_per_walker(D, P_I=29, P_P=28, H=64, L=3):
    """dense forward-Laplacian count (D+2) * sum_nets 2 (D H + H^2 + H D P)."""
    net = lambda P: 2 * (D * H + H * H + H * D * P)
    return (D + 2) * (L * net(P_I) + net(P_P))


# ------------------------------------------------------------------------------------------------- CPU reference arm
def cpu_reference(wl, n_sample, reps, threads=None):
    """Times the vectorised CPU restatement of the reference (oracle/fast_cpu.py) on all host cores."""
    import torch
    from oracle import fast_cpu
    from oracle import fixtures as fx
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    m = fx.waveflow_model(wl["D"], degree=wl["degree"], n_knots=wl["knots"], n_layers=wl["layers"], box=wl["box"], reg=wl["reg"],
                          dtype=np.float32)
    f = fast_cpu.FastLocalEnergy(m, wl["params"], wl["protons"], dtype=torch.float32)
    x = wl["walkers"][:n_sample]
    f(x[: min(512, len(x))])                       # warm-up (allocator, MKL threads)
    times = []
    for _ in range(reps):
        t0 = time.perf_counter(); f(x); times.append(time.perf_counter() - t0)
    return len(x) / float(np.median(times)), float(np.median(times)), threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = workload(args.workload)
    n_sample = wl["n_walkers"]                     # the whole workload every step: same config as the B200 arm
    import torch
    from oracle import fast_cpu
    from oracle import fixtures as fx
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    m = fx.waveflow_model(wl["D"], dtype=np.float32)
    f = fast_cpu.FastLocalEnergy(m, wl["params"], wl["protons"], dtype=torch.float32)
    x = wl["walkers"][:n_sample]
    chunk = 8192                                   # cache-sized blocks (the port's intermediates are [N, D + 2, D, P] tensors)

    def step():
        for i in range(0, n_sample, chunk):
            f(x[i:i + chunk])
    # bounded run: a step costs ~1 s on 16 cores; cap the untimed warm-up so that K + W steps stay within a few minutes
    for _ in range(min(args.warmup, 2)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    val = n_sample / dt
    sample = f"all {n_sample} walkers of the workload every step, in blocks of {chunk}"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": common_config(wl),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                             "note": "vectorised torch-CPU restatement of the reference (oracle/fast_cpu.py); the reference's "
                                     "own JAX path cannot be installed in this image"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25",
                                       "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in Path(self.f.name).read_text().strip().splitlines() if r.count(",") >= 8]
        os.unlink(self.f.name)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [r for r in rows if float(r[3]) > 200.0] or rows     # samples taken under load (warm-up + timed steps)
        sm = [float(r[1]) for r in busy]
        reasons = set()
        for r in rows:
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if "Active" in r[col] and "Not" not in r[col]:
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(rows[0][2]), "samples": len(rows),
                "power_w_max": max(float(r[3]) for r in rows), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------- small-batch configs
def gpu_us(fn, reps=200, warm=20):
    """Mean device time of `fn` in microseconds (CUDA events around `reps` back-to-back calls, after `warm` calls)."""
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


def wall_us(fn, reps=50, warm=5):
    """Mean wall time of a synchronous call in microseconds (the caller's view: launch + kernel + result on the host)."""
    for _ in range(warm):
        fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e6


def bench_c1(dev, with_cpu):
    """BASELINE configs[0]: benchmark_tests.get_model('MFlow', 0.02, degree 5, 23 knots, 3 layers), 2-D, batch 256:
    log_pdf + sample through the reference's closures, next to the CPU port (numpy float32 restatement)."""
    import torch
    from waveflow_b200 import benchmark_tests as bt
    init = bt.get_model('MFlow', 0.02, spline_degree=5, num_knots=23, num_layers=3, prior_spline_degree=3, prior_num_knots=15)
    _p0, log_pdf, sample = init(0, 2)
    rng = np.random.Generator(np.random.PCG64(0))
    tp = [()] if False else []
    for _ in range(3):
        tp += [net_params(rng, 2, 28), ()]
    params = (tp, net_params(rng, 2, 16))
    # double-circles-like data in [0.025, 0.975]^2 (benchmark_tests.py:40-45): two noisy rings, min-max scaled
    ang = rng.uniform(0, 2 * np.pi, 256); rad = np.where(rng.uniform(size=256) < 0.5, 1.0, 0.5) + 0.05 * rng.standard_normal(256)
    pts = np.stack([rad * np.cos(ang), rad * np.sin(ang)], 1)
    pts = (0.025 + 0.95 * (pts - pts.min(0)) / (pts.max(0) - pts.min(0))).astype(np.float32)
    x = torch.from_numpy(pts).to(dev)
    tparams = to_device_tree(params, dev)
    out = {"description": "BASELINE configs[0]: MFlow(3 x (IMADE, Reverse), I degree 5 / 23 knots, reg 0.02, M prior degree 3 / 15 knots), "
                          "2-D double-circles-like batch of 256; the reference's own point is CPU-only",
           "batch": 256,
           "gpu_log_pdf_us": gpu_us(lambda: log_pdf(tparams, x)),
           "gpu_sample_us": gpu_us(lambda: sample(7, tparams, 256, device=dev), reps=50, warm=5),
           "gpu_log_pdf_wall_us": wall_us(lambda: log_pdf(tparams, x).cpu()),
           "gpu_sample_wall_us": wall_us(lambda: sample(7, tparams, 256, device=dev).cpu())}
    if with_cpu:
        from oracle import fixtures as fx
        from oracle import live
        m32 = fx.mflow_model(dtype=np.float32)
        m64 = fx.mflow_model()
        p32 = fx.cast_params(params, np.float32)
        crng = np.random.default_rng(3)
        out["cpu_log_pdf_us"] = wall_us(lambda: live.log_pdf(m32, p32, pts), reps=20, warm=2)
        out["cpu_sample_us"] = wall_us(lambda: live.mflow_sample(m32, p32, 256, crng), reps=5, warm=1)
        out["cpu"] = {"kind": "port", "cores": 1, "note": "numpy float32 restatement (oracle/live.py: log_pdf, mflow_sample)"}
        lp64 = live.log_pdf(m64, fx.cast_params(params, np.float64), pts.astype(np.float64))
        out["parity_log_pdf"] = rel_stats(log_pdf(tparams, x).cpu().numpy(), lp64, 1e-5, live.log_pdf(m32, p32, pts), scale=1.0)
    return out


def bench_c2(dev, with_cpu):
    """BASELINE configs[1]: He, L = 10, published checkpoint, batch 256: local energy through h_fn(params, x)."""
    import torch
    from waveflow_b200 import model_factory, vqmc
    from waveflow_b200.utils import physics
    w2 = workload("vqmc_c2")
    init2 = model_factory.get_waveflow_model(2, base_spline_degree=6, i_spline_degree=6, n_prior_internal_knots=23,
                                             n_i_internal_knots=23, i_spline_reg=0.05, i_spline_reverse_fun_tol=1e-6,
                                             n_flow_layers=3, box_size=10.0, xu_coord_type="mean", cached_bases_root=None)
    _p2, psi2, _lp2, _s2 = init2(0, 2)
    h2 = physics.construct_hamiltonian_function(psi2, protons=w2["protons"], n_space_dimensions=1, eps=0.0)
    tparams = to_device_tree(w2["params"], dev)
    xb = torch.from_numpy(w2["walkers"]).to(dev)
    out = {"description": w2["desc"], "batch": 256,
           "gpu_h_fn_us": gpu_us(lambda: h2(tparams, xb)),
           "gpu_loss_fn_efficient_wall_us": wall_us(lambda: float(vqmc.loss_fn_efficient(tparams, psi2, h2, xb)))}
    if with_cpu:
        from oracle import fast_cpu
        from oracle import fixtures as fx
        f32 = fast_cpu.FastLocalEnergy(fx.waveflow_model(2, dtype=np.float32), w2["params"], w2["protons"], dtype=torch.float32)
        out["cpu_local_energy_us"] = wall_us(lambda: f32(w2["walkers"]), reps=10, warm=2)
        out["cpu"] = {"kind": "port", "cores": torch.get_num_threads(), "note": "oracle/fast_cpu.py, float32"}
        f64 = fast_cpu.FastLocalEnergy(fx.waveflow_model(2), w2["params"], w2["protons"], dtype=torch.float64)
        r64, r32 = f64(w2["walkers"]), f32(w2["walkers"])
        g = h2(tparams, xb, return_all=True)
        out["parity"] = {"psi": rel_stats(g["psi"].cpu().numpy(), r64["psi"], 1e-5, r32["psi"]),
                         "eloc": rel_stats(g["eloc"].cpu().numpy(), r64["eloc"], 1e-4, r32["eloc"])}
    return out


def to_device_tree(tree, dev):
    import torch
    if isinstance(tree, np.ndarray):
        return torch.from_numpy(np.ascontiguousarray(tree)).to(dev)
    if isinstance(tree, (list, tuple)):
        return type(tree)(to_device_tree(t, dev) for t in tree)
    return tree


# ------------------------------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # One JSON line on stdout and nothing else: libraries (NCCL prints its version banner to stdout) get stderr as
    # their fd 1 for the whole run; the result line is written to the saved descriptor at the end.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from waveflow_b200 import _ffi, _live, model_factory, vqmc
    from waveflow_b200.utils import physics

    wl = workload(args.workload)
    D = wl["D"]
    # the reference's own entry points: model factory -> psi -> Hamiltonian
    init_fun = model_factory.get_waveflow_model(D, base_spline_degree=wl["degree"], i_spline_degree=wl["degree"],
                                                n_prior_internal_knots=wl["knots"], n_i_internal_knots=wl["knots"],
                                                i_spline_reg=wl["reg"], i_spline_reverse_fun_tol=1e-6, n_flow_layers=wl["layers"],
                                                box_size=wl["box"], xu_coord_type="mean", cached_bases_root=None)
    _init_params, psi, log_pdf, sample = init_fun(0, D)
    h_fn = physics.construct_hamiltonian_function(psi, protons=wl["protons"], n_space_dimensions=1, eps=0.0)
    params = wl["params"]
    est = vqmc.EnergyEstimator(h_fn, params, dev)
    lo, hi = est.shard(wl["n_walkers"], rank, world)
    x_host = torch.from_numpy(wl["walkers"][lo:hi].copy()).pin_memory()
    x_dev = x_host.to(dev)
    n_local = hi - lo
    steps, warm = args.steps, args.warmup

    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev) if not args.no_l2_flush else None

    def l2_flush():
        if flush is not None:
            flush.fill_(1.0)

    # estimator exchange: one-shot all-reduce over NVLink peer memory (wf_p2p_allreduce_sums), NCCL if unavailable
    exchange_kind = "none"
    if world > 1:
        exchange_kind = "nccl all_reduce"
        if not args.nccl_exchange:
            try:
                est.peer = vqmc.PeerExchange(dev)
                exchange_kind = "peer-memory one-shot kernel (wf_p2p_allreduce_sums)"
            except Exception as exc:                                      # noqa: BLE001 -- any failure: keep NCCL
                print(f"[bench] peer-memory exchange unavailable ({type(exc).__name__}: {exc}); using NCCL", file=sys.stderr)
                est.peer = None
        flag = torch.tensor([1 if est.peer is not None else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)                       # all ranks or none
        if int(flag.item()) == 0:
            est.peer = None
            exchange_kind = "nccl all_reduce"

    # tensor-core path + peer exchange: the exchange runs in the tail of the local-energy kernel (one launch per step)
    live_mode = os.environ.get("WAVEFLOW_B200_LIVE", "auto")
    _w_sel, layout = _live.select_weights(est.spec, est.packed, n_local, lap=True)
    fused = world > 1 and est.peer is not None and layout == _ffi.WEIGHTS_TC

    def step(sums):
        return est.step_sums(x_dev, sums)

    # ---------------- device-timed region: K steps, inputs resident in HBM
    all_sums = torch.zeros(warm + steps, 4, dtype=torch.float64, device=dev)
    clocks = ClockSampler(local_rank)        # nvidia-smi needs ~0.2 s to start: launched before the warm-up
    time.sleep(0.3)
    for i in range(warm):
        l2_flush(); step(all_sums[i])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(steps)]
    t_wall0 = time.perf_counter()
    for i in range(steps):
        l2_flush()                                   # outside the per-step event brackets
        a, k, b = ev[i]
        a.record()
        if fused:
            est.step_sums(x_dev, all_sums[warm + i])     # ONE launch: local energy + peer exchange in the kernel tail
            k.record()
        else:
            est.local_sums(x_dev, all_sums[warm + i])
            k.record()                                   # end of the dominant kernel
            est.exchange(all_sums[warm + i])
        b.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    clk = clocks.stop()
    step_ms = np.array([a.elapsed_time(b) for a, k, b in ev])
    kern_ms = np.array([a.elapsed_time(k) for a, k, b in ev])
    total_ms = float(step_ms.sum())

    # ---------------- end-to-end: host walkers -> pinned H2D -> h_fn(params, x) -> estimator D2H, every step.
    # The call is the reference's own signature, h_fn(params, walkers): the raw parameter pytree goes in on every call
    # (the packed kernel layout comes from the per-model cache keyed on the leaves' identity / version, _live.packed_for).
    params_dev = to_device_tree(params, dev)
    # A few steps in flight, as a user's loop would be written: step i + 1's walkers are copied (copy stream, pinned -> device)
    # while step i's kernel runs, and a step's 32-byte result is read back (pinned, asynchronous) and consumed on the host
    # DEPTH - 1 steps later, so a host hiccup of a millisecond or two does not drain the GPU queue (with two steps in flight the
    # end-to-end figure moved between 96 % and 99 % of the device figure from run to run).
    # Every step still does its own H2D copy of the inputs and its own D2H read of the result inside the timed region.
    DEPTH = 4 if world == 1 else 2          # the multi-GPU runs of the round were taken (and were stable) with two steps in flight
    copy_stream = torch.cuda.Stream(device=dev)
    xbuf = [torch.empty_like(x_dev) for _ in range(DEPTH)]
    sbuf = [torch.zeros(4, dtype=torch.float64, device=dev) for _ in range(DEPTH)]
    hbuf = [torch.zeros(4, dtype=torch.float64).pin_memory() for _ in range(DEPTH)]
    ev_in = [torch.cuda.Event() for _ in range(DEPTH)]
    ev_out = [torch.cuda.Event() for _ in range(DEPTH)]
    ev_free = [torch.cuda.Event() for _ in range(DEPTH)]
    main_stream = torch.cuda.current_stream(dev)

    def e2e_run(n_steps):
        results = []
        for i in range(n_steps + DEPTH - 1):
            if i < n_steps:
                b = i % DEPTH
                with torch.cuda.stream(copy_stream):
                    if i >= DEPTH:
                        copy_stream.wait_event(ev_free[b])             # the kernel of step i - DEPTH has consumed this buffer
                    xbuf[b].copy_(x_host, non_blocking=True)
                    ev_in[b].record(copy_stream)
                main_stream.wait_event(ev_in[b])
                sbuf[b].zero_()
                h_fn(params_dev, xbuf[b], return_all=True, sums=sbuf[b], exchange=est.peer if world > 1 else None, want=())
                ev_free[b].record(main_stream)
                res = est.peer.out if (world > 1 and est.peer is not None) else (est.exchange(sbuf[b]) if world > 1 else sbuf[b])
                hbuf[b].copy_(res, non_blocking=True)
                ev_out[b].record(main_stream)
            j = i - (DEPTH - 1)
            if 0 <= j < n_steps:
                pb = j % DEPTH
                ev_out[pb].synchronize()                               # the result of step j is on the host
                results.append(hbuf[pb].clone())
        return results
    e2e_run(warm)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    s_last = e2e_run(steps)[-1]
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0

    # max over ranks
    t = torch.tensor([total_ms, e2e_s, float(kern_ms.mean())], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_s, kern_ms_mean = [float(v) for v in t.cpu()]
    ms_per_step = total_ms / steps
    value = wl["n_walkers"] / (ms_per_step * 1e-3)
    e2e_value = wl["n_walkers"] / (e2e_s / steps)

    # ---------------- the training step (value_and_grad(loss_fn_efficient) + Adam) on the same workload
    train_ms = None
    if not args.no_sweep:
        from waveflow_b200 import _train
        opt_init, opt_update, get_params = _train.adam(1e-4, device=dev)
        opt_state = opt_init(params)
        tparams = get_params(opt_state)
        # training walkers: the workload's set without the neighbourhood of the nodes of psi (|psi| > 1e-4 max|psi|; next to a
        # node one walker's float32 rounding noise in H psi / psi swamps the batch loss and gradient).  Every
        # rank evaluates the whole set, so the filter -- and the contiguous shards cut from it -- are identical for every N.
        x_all = torch.from_numpy(wl["walkers"]).to(dev)
        psi_all = h_fn(params_dev, x_all, return_all=True, want=("psi",))["psi"]
        x_keep = x_all[psi_all.abs() > 1e-4 * psi_all.abs().max()]
        n_train = (x_keep.shape[0] // 8) * 8
        tlo, thi = est.shard(n_train, rank, world)
        x_train = x_keep[tlo:thi].contiguous()
        del x_all, psi_all, x_keep
        # step-0 loss and gradient norm BEFORE any update: a cross-N checksum (must agree between 1, 2, 4, 8 GPUs)
        loss0, g0 = vqmc.value_and_grad_efficient(tparams, psi, h_fn, x_train, 0.0, opt_state=opt_state, flat_grad=True, n_total=n_train)
        train_check = {"walkers": n_train, "step0_loss": float(loss0), "step0_grad_l2": float(g0.double().norm()),
                       "step0_grad_abs_sum": float(g0.double().abs().sum())}

        gx = None
        if world > 1 and est.peer is not None:
            try:
                gx = vqmc.GradExchange(opt_state.flat.numel(), dev)
            except Exception as exc:                                      # noqa: BLE001 -- any failure: NCCL all-reduce
                print(f"[bench] peer-memory gradient exchange unavailable ({type(exc).__name__}: {exc}); using NCCL", file=sys.stderr)
                gx = None
            ok = torch.tensor([1 if gx is not None else 0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0:
                gx = None

        def tstep(i):
            return vqmc.train_step_efficient(i, psi, h_fn, opt_update, opt_state, tparams, x_train, 0.0, n_total=n_train, exchange=gx)
        for i in range(3):
            tstep(i)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        kt = 10
        a.record()
        for i in range(kt):
            _, tl = tstep(3 + i)
        b.record(); torch.cuda.synchronize()
        train_ms = a.elapsed_time(b) / kt
        if world > 1:
            tms = torch.tensor([train_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            train_ms = float(tms.item())
        train_loss = float(tl)

    # ---------------- BASELINE configs 3 and 5 over N GPUs: the fixed sample sets sharded by rows, no collective on the data
    # path; whole-job samples/s = total samples / max over ranks of the device time
    flow_scaling = None
    if world > 1 and not args.no_sweep:
        from waveflow_b200.flows.neural_splines import coupling_flow, coupling_flow_tc, pack_fcnn_tc
        flow_scaling = {}

        def rank_time(fn, reps):
            fn(); torch.cuda.synchronize(); dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record(); torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        gW = lambda r_, a_, b_: torch.from_numpy((r_.standard_normal((a_, b_)) / np.sqrt(a_)).astype(np.float32)).to(dev)
        zz = lambda n_: torch.zeros(n_, device=dev)
        gsh = torch.Generator(device=dev); gsh.manual_seed(1000 + rank)
        # config 3: D = 8, K = 32, hidden 8, 8 layers, 2^24 samples in total
        crng = np.random.Generator(np.random.PCG64(0))
        Dc, K3, h3, L3, M3 = 8, 32, 8, 8, (1 << 24) // world
        od = (3 * K3 - 1) * Dc // 2
        layers3 = [tuple([(gW(crng, Dc // 2, h3), zz(h3)), (), (gW(crng, h3, h3), zz(h3)), (), (gW(crng, h3, od), zz(od))] for _f in range(2))
                   for _ in range(L3)]
        x3 = torch.rand(M3, Dc, device=dev, generator=gsh) * 6 - 3
        for inv in (False, True):
            ms = rank_time(lambda: coupling_flow(layers3, x3, K3, 3.0, h3, inverse=inv), 3)
            flow_scaling[f"c3_D8_{'inverse' if inv else 'density'}"] = {"samples_total": M3 * world, "ms": ms,
                                                                        "samples_per_s": M3 * world / (ms * 1e-3)}
        del x3, layers3
        # config 5: D = 64, K = 64, hidden 512, 8 layers, 2^22 samples in total (tensor-core path)
        crng = np.random.Generator(np.random.PCG64(0))
        D5, K5, H5, L5, M5 = 64, 64, 512, 8, (1 << 22) // world
        od = (3 * K5 - 1) * D5 // 2
        w5 = torch.cat([pack_fcnn_tc([(gW(crng, D5 // 2, H5), zz(H5)), (), (gW(crng, H5, H5), zz(H5)), (), (gW(crng, H5, od), zz(od))], dev)
                        for _ in range(2 * L5)]).contiguous()
        x5 = torch.rand(M5, D5, device=dev, generator=gsh) * 6 - 3
        coupling_flow_tc(w5, L5, x5[: 1 << 16], 3.0)
        ms = rank_time(lambda: coupling_flow_tc(w5, L5, x5, 3.0), 1)
        flow_scaling["c5_D64_logprob"] = {"samples_total": M5 * world, "ms": ms, "samples_per_s": M5 * world / (ms * 1e-3),
                                          "algorithmic_tflops_total": 109051904.0 * M5 * world / (ms * 1e-3) / 1e12}
        del x5, w5
        torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    hbm_peak, hbm_src = (peaks["hbm_gbs"], "MEASURED_PEAKS.json (measured copy)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")

    # ---------------- FP32 ceiling (measured here): the local-energy kernel is FMA-issue bound, not HBM / tensor bound
    out = torch.zeros(1, device=dev)
    iters, blocks = 1 << 16, 148 * 8
    for _ in range(2):
        _ffi.check(_ffi.lib.wf_probe_fma(iters, blocks, _ffi.ptr(out), _ffi.stream_ptr()))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); _ffi.check(_ffi.lib.wf_probe_fma(iters, blocks, _ffi.ptr(out), _ffi.stream_ptr())); b.record()
    torch.cuda.synchronize()
    fp32_peak = 2.0 * 8 * iters * blocks * 256 / (a.elapsed_time(b) * 1e-3) / 1e12

    fl = flops_per_walker(D)
    achieved = fl * n_local / (kern_ms_mean * 1e-3) / 1e12
    tc_path = layout == _ffi.WEIGHTS_TC
    sm_max = clk.get("sm_max_mhz") or 1965.0
    fp32_nominal = 148 * 128 * 2 * sm_max * 1e6 / 1e12            # SMs x FP32 lanes x 2 FLOP x max SM clock
    roofline = {"bound": "fp32",
                "kernel": (f"wf::ltc::live_tc_kernel<{D}, true> (wf_local_energy, WF_WEIGHTS_TC: layers 2/3 on tcgen05 3xTF32, jet "
                           "algebra as the TMEM epilogue)") if tc_path else f"wf::live_kernel<{D}, true> (wf_local_energy)",
                "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak,
                "traffic": TRAFFIC.get(("live_tc" if tc_path else "live_simt", D, n_local)),
                "algorithmic_flops_per_walker": fl, "walkers_per_launch": n_local,
                "peak_source": "FFMA probe kernel (wf_probe_fma) timed in this run; MEASURED_PEAKS.json holds no FP32 figure",
                "peak_nominal": fp32_nominal, "frac_of_nominal": achieved / fp32_nominal,
                "note": "issue bound (SURVEY F8 / 8d): 24 B/walker of HBM traffic, so the HBM roofline does not apply; the dense "
                        "products run on the tensor pipe but the step is bound by the FP32 / SFU / shuffle epilogue, hence the "
                        "FP32 ceiling as denominator; the HBM-bound operators are reported under 'spline_sweep' / 'rqs_sweep'"}

    # ---------------- HBM-bound operator sweep: fused table-spline value + derivative + log-derivative, 2^24 elements
    sweep = None
    if world == 1 and not args.no_sweep:
        from waveflow_b200.splines.factories import spline_apply
        from waveflow_b200.splines.tables import SplineTables
        tabs = SplineTables.get("I", 6, 23)
        M = 1 << 24
        g = torch.Generator(device=dev); g.manual_seed(0)
        c = torch.rand(M, tabs.P, device=dev, generator=g); c /= c.sum(-1, keepdim=True)
        xs = torch.rand(M, device=dev, generator=g)
        for _ in range(3):
            spline_apply(tabs, c, xs, 0, 2, logd=True)
        ts = []
        for _ in range(10):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); spline_apply(tabs, c, xs, 0, 2, logd=True); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = float(np.mean(ts))
        bytes_per_el = 4 * tabs.P + 12
        gbs = M * bytes_per_el / (ms * 1e-3) / 1e9
        val, grd, logd = spline_apply(tabs, c, xs, 0, 2, logd=True)
        par_spline = None
        if not args.no_cpu_baseline:
            from oracle import live as olive
            ns = 1 << 16
            c64, x64 = c[:ns].double().cpu().numpy(), xs[:ns].double().cpu().numpy()
            c32, x32 = c[:ns].cpu().numpy(), xs[:ns].cpu().numpy()
            t64, t32 = tabs.tab64, tabs.tab32
            par_spline = {"sample": f"first {ns} elements of the timed sweep vs oracle/live.py (float64)",
                          "value": rel_stats(val[:ns].cpu().numpy(), olive.spline_apply(t64, c64, x64, 0), 1e-5,
                                             olive.spline_apply(t32, c32, x32, 0)),
                          "derivative": rel_stats(grd[:ns].cpu().numpy(), olive.spline_apply(t64, c64, x64, 1), 1e-5,
                                                  olive.spline_apply(t32, c32, x32, 1)),
                          "log_derivative": rel_stats(logd[:ns].cpu().numpy(), np.log(olive.spline_apply(t64, c64, x64, 1) + 1e-7), 1e-5,
                                                      np.log(olive.spline_apply(t32, c32, x32, 1) + np.float32(1e-7)), scale=1.0)}
        del val, grd, logd
        sweep = {"kernel": "spline_local_kernel<true> (wf_spline_apply_local)", "elements": M, "P": tabs.P, "ms": ms, "parity": par_spline,
                 "elements_per_s": M / (ms * 1e-3),
                 "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                              "traffic": TRAFFIC.get(("spline_local", tabs.P, M)), "algorithmic_bytes_per_element": bytes_per_el, "peak_source": hbm_src,
                              "inputs": "2.1 GB per launch (> 126 MB L2)"}}
        del c, xs

    # ---------------- BASELINE config 3: RQS operator (HBM-bound) and the fused 8-layer coupling flow, 2^24 samples
    rqs_sweep = None
    coupling = None
    if world == 1 and not args.no_sweep:
        from waveflow_b200.flows.neural_splines import coupling_flow, unconstrained_RQS
        g = torch.Generator(device=dev); g.manual_seed(0)
        M, K = 1 << 24, 32
        uw = torch.randn(M, K, device=dev, generator=g); uh = torch.randn(M, K, device=dev, generator=g)
        ud = torch.randn(M, K - 1, device=dev, generator=g)
        xs = torch.rand(M, device=dev, generator=g) * 6 - 3
        res = {}
        for inv in (False, True):
            for _ in range(3):
                unconstrained_RQS(xs, uw, uh, ud, inverse=inv, tail_bound=3.0)
            ts = []
            for _ in range(5):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); unconstrained_RQS(xs, uw, uh, ud, inverse=inv, tail_bound=3.0); b.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            ms = float(np.mean(ts))
            gbs = M * 4 * (3 * K + 2) / (ms * 1e-3) / 1e9
            res["inverse" if inv else "forward"] = {"ms": ms, "elements_per_s": M / (ms * 1e-3),
                                                     "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                                                                  "frac": gbs / hbm_peak,
                                                                  "traffic": TRAFFIC.get(("rqs_staged", K, M, inv)),
                                                                  "algorithmic_bytes_per_element": 4 * (3 * K + 2), "peak_source": hbm_src}}
        par_rqs = None
        if not args.no_cpu_baseline:
            from oracle import rqs as orqs
            ns = 1 << 16
            h = lambda t_: t_[:ns].cpu().numpy()
            x_, uw_, uh_, ud_ = h(xs), h(uw), h(uh), h(ud)
            d64 = lambda a_: a_.astype(np.float64)
            par_rqs = {"sample": f"first {ns} elements of the timed sweep vs oracle/rqs.py (float64 values, float32 bins)"}
            for inv in (False, True):
                o_, l_, b_ = unconstrained_RQS(xs[:ns], uw[:ns], uh[:ns], ud[:ns], inverse=inv, tail_bound=3.0, return_bin_idx=True)
                _, _, be_ = unconstrained_RQS(xs[:ns], uw[:ns], uh[:ns], ud[:ns], inverse=inv, tail_bound=3.0, return_bin_idx=True,
                                              exact_bins=True)
                ro, rl, _rb = orqs.unconstrained_rqs(d64(x_), d64(uw_), d64(uh_), d64(ud_), inv, 3.0, return_bin=True)
                o32, l32, b32 = orqs.unconstrained_rqs(x_, uw_, uh_, ud_, inv, 3.0, return_bin=True)
                par_rqs["inverse" if inv else "forward"] = {
                    "outputs": rel_stats(o_.cpu().numpy(), ro, 1e-5, o32, scale=3.0),
                    "logabsdet": rel_stats(l_.cpu().numpy(), rl, 1e-5, l32, scale=1.0),
                    "bins_equal_float32_reference_exact_mode": float(np.mean(be_.cpu().numpy() == b32)),
                    "bins_equal_float32_reference_fast_mode": float(np.mean(b_.cpu().numpy() == b32))}
        rqs_sweep = {"kernel": "rqs_staged_kernel<32> (wf_rqs_apply)", "elements": M, "K": K, "inputs": "6.6 GB per launch (> L2)",
                     "parity": par_rqs, **res}
        del uw, uh, ud, xs
        coupling = {}
        crng = np.random.Generator(np.random.PCG64(0))
        for Dc in (2, 8):
            hidden, L = 8, 8
            out_dim = (3 * K - 1) * Dc // 2
            layers = []
            for _ in range(L):
                pair = []
                for _f in range(2):
                    gW = lambda a, b: torch.from_numpy((crng.standard_normal((a, b)) / np.sqrt(a)).astype(np.float32)).to(dev)
                    z = lambda n: torch.zeros(n, device=dev)
                    pair.append([(gW(Dc // 2, hidden), z(hidden)), (), (gW(hidden, hidden), z(hidden)), (), (gW(hidden, out_dim), z(out_dim))])
                layers.append(tuple(pair))
            x = torch.rand(M, Dc, device=dev, generator=g) * 6 - 3
            r = {}
            for inv in (False, True):
                for _ in range(2):
                    coupling_flow(layers, x, K, 3.0, hidden, inverse=inv)
                ts = []
                for _ in range(3):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(); coupling_flow(layers, x, K, 3.0, hidden, inverse=inv); b.record(); torch.cuda.synchronize()
                    ts.append(a.elapsed_time(b))
                ms = float(np.mean(ts))
                r["inverse" if inv else "density"] = {"ms": ms, "samples_per_s": M / (ms * 1e-3),
                                                      "hbm_frac_of_measured": M * (8 * Dc + 4) / (ms * 1e-3) / 1e9 / hbm_peak}
            par_c = None
            if not args.no_cpu_baseline:
                from oracle import rqs as orqs
                ns = 1 << 14
                npl = [tuple([(a_.double().cpu().numpy(), b_.double().cpu().numpy()) for (a_, b_) in (f[0], f[2], f[4])] for f in pair)
                       for pair in layers]
                np32 = [tuple([(a_.astype(np.float32), b_.astype(np.float32)) for (a_, b_) in f] for f in pair) for pair in npl]
                xh = x[:ns].cpu().numpy()
                y_, ld_ = coupling_flow(layers, x[:ns], K, 3.0, hidden)
                ry, rld = orqs.coupling_flow_direct(npl, xh.astype(np.float64), K, 3.0)
                y32, ld32 = orqs.coupling_flow_direct(np32, xh, K, 3.0)
                par_c = {"sample": f"first {ns} samples of the timed sweep vs oracle/rqs.py (float64)",
                         "outputs": rel_stats(y_.cpu().numpy(), ry, 1e-5, y32, scale=3.0),
                         "log_det": rel_stats(ld_.cpu().numpy(), rld, 1e-5, ld32, scale=1.0)}
            coupling[f"D{Dc}"] = {"samples": M, "K": K, "layers": L, "hidden": hidden, "parity": par_c, **r,
                                 "note": "MUFU/issue bound by construction (SURVEY 8d regime ii): 8D+4 bytes per sample"}
            del x

    # ---------------- BASELINE config 5: 64-D coupling flow, K = 64, width-512 conditioners on the tensor cores (2^22 samples)
    tc_sweep = None
    if world == 1 and not args.no_sweep:
        from waveflow_b200.flows.neural_splines import coupling_flow_tc, pack_fcnn_tc
        D5, K5, H5, L5, M5 = 64, 64, 512, 8, 1 << 22
        crng = np.random.Generator(np.random.PCG64(0))
        out_dim = (3 * K5 - 1) * D5 // 2
        parts, nets5 = [], []
        for _ in range(2 * L5):
            gN = lambda a, b: (crng.standard_normal((a, b)) / np.sqrt(a)).astype(np.float32)
            z = lambda n: torch.zeros(n, device=dev)
            Ws = [gN(D5 // 2, H5), gN(H5, H5), gN(H5, out_dim)]
            nets5.append([(W_, np.zeros(W_.shape[1], dtype=np.float32)) for W_ in Ws])
            tW = [torch.from_numpy(W_).to(dev) for W_ in Ws]
            parts.append(pack_fcnn_tc([(tW[0], z(H5)), (), (tW[1], z(H5)), (), (tW[2], z(out_dim))], dev))
        w5 = torch.cat(parts).contiguous()
        del parts
        g5 = torch.Generator(device=dev); g5.manual_seed(0)
        x5 = torch.rand(M5, D5, device=dev, generator=g5) * 6 - 3
        coupling_flow_tc(w5, L5, x5[: 1 << 17], 3.0)                       # warm-up
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); y5, ld5 = coupling_flow_tc(w5, L5, x5, 3.0); b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        par5 = None
        if not args.no_cpu_baseline:
            from oracle import rqs as orqs
            ns = 512
            lay32 = [(nets5[2 * i], nets5[2 * i + 1]) for i in range(L5)]
            lay64 = [tuple([(a_.astype(np.float64), b_.astype(np.float64)) for (a_, b_) in f] for f in pair) for pair in lay32]
            xh = x5[:ns].cpu().numpy()
            ry, rld = orqs.coupling_flow_direct(lay64, xh.astype(np.float64), K5, 3.0)
            y32, ld32 = orqs.coupling_flow_direct(lay32, xh, K5, 3.0)
            par5 = {"sample": f"first {ns} samples of the timed sweep vs oracle/rqs.py (float64)",
                    "outputs": rel_stats(y5[:ns].cpu().numpy(), ry, 1e-5, y32, scale=3.0),
                    "log_det": rel_stats(ld5[:ns].cpu().numpy(), rld, 1e-5, ld32, scale=1.0)}
        flops = 2.0 * L5 * 2 * (32 * H5 + H5 * H5 + H5 * out_dim) * M5            # 109.1 MFLOP / sample
        bf16 = peaks.get("bf16_tflops_sustained", 1400.0)
        useful = flops / (ms * 1e-3) / 1e12
        tc_sweep = {"kernel": "tc_rqs_kernel / tc_hidden_kernel (wf_rqs_coupling_flow_tc): tcgen05.mma kind::tf32, 3-pass hi/lo split",
                    "samples": M5, "D": D5, "K": K5, "hidden": H5, "layers": L5, "ms": ms, "samples_per_s": M5 / (ms * 1e-3), "parity": par5,
                    "roofline": {"bound": "tensor", "achieved": useful, "peak": bf16, "unit": "TFLOP/s", "frac": useful / bf16,
                                 "traffic": None, "algorithmic_flops_per_sample": flops / M5,
                                 "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if "bf16_tflops_sustained" in peaks else "fallback",
                                 "tf32_tflops_issued": 3 * useful, "tf32_peak_estimate": bf16 / 2,
                                 "frac_of_tf32_peak_issued": 3 * useful / (bf16 / 2),
                                 "note": "float32-grade log-probs need a 3-pass TF32 split: the issued tensor work is 3x the algorithmic "
                                         "FLOPs and TF32 runs at half the bf16 rate, so frac <= 1/6 by construction"}}
        del x5, y5, ld5, w5

    train = None
    if train_ms is not None:
        ms = train_ms
        tfl = 3.0 * flops_per_walker(D) * int(x_train.shape[0]) / (ms * 1e-3) / 1e12          # forward + ~2x for the reverse pass
        train = {"api": "vqmc.train_step_efficient(epoch, psi, h_fn, opt_update, opt_state, params, batch, running_average)",
                 "kernels": "wf_vqmc_loss_grad (layer-wise jets; the 64-wide conditioner layers, their input adjoints and all weight gradients on "
                            "tcgen05 3xTF32 -- csrc/train_tc.cuh: lin_tc_kernel, wgrad_tc_kernel; tanh / spline-head kernels on CUDA cores; "
                            "forward + reverse) + wf_adam_step",
                 "walkers_total": n_train, "walkers_per_gpu": int(x_train.shape[0]), "ms_per_step": ms,
                 "walkers_per_s": n_train / (ms * 1e-3), "loss_after_13_steps": train_loss, "cross_n_checksum": train_check,
                 "walker_set": "the workload's 65536 walkers filtered to |psi| > 1e-4 max|psi| (identical on every rank and for every N)",
                 "exchange": "none" if world == 1 else (
                     f"flat gradient ({opt_state.flat.numel()} floats) + 32-byte loss sums per step: " +
                     ("one peer-memory kernel (wf_p2p_allreduce_vec) inside the CUDA graph of the step" if gx is not None
                      else "two NCCL all-reduces between eagerly launched kernels")),
                 "algorithmic_tflops_per_gpu": tfl, "frac_of_measured_fp32_fma": tfl / fp32_peak,
                 "gpu_launches_per_step": 70 * ((int(x_train.shape[0]) + 65535) // 65536) + 1}
        if world == 1:
            # the reference's own training configuration (BASELINE configs[1]): He, batch 256 -- launch-bound, CUDA-graph replay
            from waveflow_b200 import _train
            w2 = workload("vqmc_c2")
            init2 = model_factory.get_waveflow_model(2, base_spline_degree=6, i_spline_degree=6, n_prior_internal_knots=23,
                                                     n_i_internal_knots=23, i_spline_reg=0.05, i_spline_reverse_fun_tol=1e-6,
                                                     n_flow_layers=3, box_size=10.0, xu_coord_type="mean", cached_bases_root=None)
            _p2, psi2, _lp2, _s2 = init2(0, 2)
            h2 = physics.construct_hamiltonian_function(psi2, protons=w2["protons"], n_space_dimensions=1, eps=0.0)
            oi2, ou2, gp2 = _train.adam(1e-4, device=dev)
            st2 = oi2(w2["params"])
            xb = torch.from_numpy(w2["walkers"]).to(dev)
            he = {}
            for ug in (False, True):
                for i in range(5):
                    vqmc.train_step_efficient(i, psi2, h2, ou2, st2, gp2(st2), xb, -1.8, use_graph=ug)
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for i in range(50):
                    vqmc.train_step_efficient(5 + i, psi2, h2, ou2, st2, gp2(st2), xb, -1.8, use_graph=ug)
                b.record(); torch.cuda.synchronize()
                he["cuda_graph_ms_per_step" if ug else "eager_ms_per_step"] = a.elapsed_time(b) / 50
            train["he_batch256"] = {"description": w2["desc"], **he}
        if world == 1 and not args.no_cpu_baseline:
            from oracle import fixtures as fx
            from oracle import grad as ograd
            ns = 64
            m32 = fx.waveflow_model(D, degree=wl["degree"], n_knots=wl["knots"], n_layers=wl["layers"], box=wl["box"], reg=wl["reg"])
            p32 = fx.cast_params(params, np.float32)
            t0 = time.perf_counter()
            ograd.loss_and_grad(m32, p32, wl["walkers"][:ns], wl["protons"], 0.0, dtype=np.float32)
            sec = time.perf_counter() - t0
            train["cpu_baseline"] = {"value": ns / sec, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                     "sample": f"first {ns} walkers, one pass", "seconds_per_pass": sec,
                                     "note": "torch-CPU float32 autograd restatement (three reverse passes, oracle/grad.py); Adam not included"}


    # ---------------- CPU baseline (rank 0, N = 1): bounded sample of the same workload -- and the parity of the timed workload:
    # the CUDA psi / H psi / E_loc of those walkers against the float64 oracle, with the float32 CPU values beside them
    cpu = None
    parity = None
    if not args.no_cpu_baseline:
        n_sample = min(n_local, 8192)
        if world == 1:
            v, sec, thr = cpu_reference(wl, n_sample, reps=5)
            cpu = {"value": v, "unit": UNIT, "cores": thr, "kind": "port", "sample": f"first {n_sample} walkers of the workload, 5 passes, median",
                   "seconds_per_pass": sec,
                   "note": "vectorised torch-CPU float32 restatement of the reference's jax.hessian path (oracle/fast_cpu.py); JAX is "
                           "not installable in this image"}
        from oracle import fast_cpu
        from oracle import fixtures as fx
        mk = dict(degree=wl["degree"], n_knots=wl["knots"], n_layers=wl["layers"], box=wl["box"], reg=wl["reg"])
        f64 = fast_cpu.FastLocalEnergy(fx.waveflow_model(D, **mk), wl["params"], wl["protons"], dtype=torch.float64)
        f32 = fast_cpu.FastLocalEnergy(fx.waveflow_model(D, dtype=np.float32, **mk), wl["params"], wl["protons"], dtype=torch.float32)
        xs_ = wl["walkers"][lo:lo + n_sample]
        r64, r32 = f64(xs_), f32(xs_)
        gout = h_fn(params_dev, x_dev[:n_sample], return_all=True)
        gp, gh, ge = [gout[k].cpu().numpy() for k in ("psi", "hpsi", "eloc")]
        well = np.abs(r64["psi"]) > 1e-3 * np.abs(r64["psi"]).max()          # away from the nodes of psi (E_loc = H psi / psi)
        parity = {"sample": f"first {n_sample} walkers of the timed workload vs oracle/fast_cpu.py in float64; *_float32_restatement = the "
                            "reference's own float32 arithmetic on the CPU",
                  "tolerances": "north_star: psi / log-prob 1e-5 relative, local energy 1e-4 relative",
                  "psi": rel_stats(gp, r64["psi"], 1e-5, r32["psi"]),
                  "hpsi": rel_stats(gh, r64["hpsi"], 1e-4, r32["hpsi"]),
                  "eloc": rel_stats(ge, r64["eloc"], 1e-4, r32["eloc"]),
                  "eloc_away_from_nodes": {"criterion": "|psi| > 1e-3 max|psi|", **rel_stats(ge[well], r64["eloc"][well], 1e-4, r32["eloc"][well])},
                  "path": "tensor-core" if tc_path else "cuda-core"}

    configs = None
    if world == 1 and not args.no_sweep:
        configs = {"c1_mflow_batch256": bench_c1(dev, not args.no_cpu_baseline), "c2_he_batch256": bench_c2(dev, not args.no_cpu_baseline)}

    s = s_last.numpy()
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": common_config(wl),
            "details": {"walkers_per_gpu": n_local, "parallelism": f"walker-sharded x{world}, 32-byte estimator all-reduce per step",
                        "exchange": exchange_kind + (", fused into the tail of the local-energy kernel (one launch per step)" if fused else ""),
                        "live_path": "tensor-core (wf::ltc::live_tc_kernel)" if tc_path else "cuda-core (wf::live_kernel)",
                        "l2": "flushed between timed steps (256 MiB fill outside the event brackets)" if flush is not None else "not flushed",
                        "timing": "CUDA events per step on the launch stream, summed over K steps, max over ranks"},
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(wl["n_walkers"] * D * 4),
                    "d2h_bytes_per_step": 32 * world,
                    "api": "h_fn = utils.physics.construct_hamiltonian_function(psi, protons); h_fn(params, walkers, sums=...) -- the reference's "
                           "signature with the raw parameter pytree on every call (packed layout from the per-model cache)",
                    "ms_per_step": e2e_s / steps * 1e3,
                    "pipelining": ("four" if world == 1 else "two") + " steps in flight: the H2D copy of step i + 1 and the D2H read of step i overlap the kernel of the "
                                  "neighbouring step (copy stream + events); every step performs its own copies inside the timed region"},
            # launches of this repo's kernels inside the timed region: the local-energy kernel, + the 32-thread exchange kernel
            # when it is not fused into the kernel tail
            "gpu_launches": steps * (2 if (world > 1 and est.peer is not None and not fused) else 1),
            "kernel_ms_per_step": kern_ms_mean, "wall_s_timed_region": t_wall,
            "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "configs": configs, "spline_sweep": sweep, "rqs_sweep": rqs_sweep, "coupling_flow_sweep": coupling, "tc_coupling_flow_sweep": tc_sweep, "flow_scaling": flow_scaling, "train_step": train,
            "energy_estimate": {"mean": float(s[0] / s[2]), "n": int(s[2])}}
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    os.close(json_fd)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="vqmc_c4", choices=["vqmc_c4", "vqmc_c2"])
    ap.add_argument("--no-l2-flush", action="store_true")
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--nccl-exchange", action="store_true", help="use the NCCL all-reduce for the estimator sums instead of the peer-memory kernel")
    args = ap.parse_args()
    if os.environ.get("WF_BENCH_WATCHDOG"):          # debugging aid: dump every thread's stack and exit after N seconds
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["WF_BENCH_WATCHDOG"]), exit=True)
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
