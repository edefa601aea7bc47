#!/usr/bin/env python
"""bench.py -- VQMC local-energy throughput of the fused live path on B200 (+ the HBM-bound spline operator sweep).

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference ...                      (CPU restatement of the reference on the host cores)

Workload (BASELINE.json configs[3], the VQMC config the north star asks to scale 1 -> 8 GPUs): 1-D 4-electron box,
L = 10, `get_waveflow_model(4, degree 6, 23 knots, 3 flow layers, reg 0.05)`, 65 536 walkers sharded over the ranks
(STRONG scaling); one step = one local-energy pass (psi, H psi, E_loc, block sums) over the rank's walkers followed by the
32-byte estimator all-reduce.  `--workload vqmc_c2` runs BASELINE configs[1] (He, D = 2, published checkpoint, batch 256).
Synthetic data: walkers sorted U(-10, 10)^D (PCG64 seed 1), parameters U(+-1/sqrt(fan_in)) (PCG64 seed 0).
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
    n_sample = min(wl["n_walkers"], 8192)
    import torch
    from oracle import fast_cpu
    from oracle import fixtures as fx
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    m = fx.waveflow_model(wl["D"], dtype=np.float32)
    f = fast_cpu.FastLocalEnergy(m, wl["params"], wl["protons"], dtype=torch.float32)
    x = wl["walkers"][:n_sample]
    for _ in range(args.warmup):
        f(x)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        f(x)
    dt = (time.perf_counter() - t0) / args.steps
    val = n_sample / dt
    sample = f"{n_sample} of the {wl['n_walkers']} walkers per step"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["name"], "description": wl["desc"], "sample": sample},
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

    def step(sums):
        est.local_sums(x_dev, sums)
        return est.exchange(sums)

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

    # ---------------- end-to-end: host walkers -> pinned H2D -> h_fn public API -> estimator D2H, every step
    def e2e_step():
        xd = x_host.to(dev, non_blocking=True)
        s = torch.zeros(4, dtype=torch.float64, device=dev)
        h_fn(params, xd, return_all=True, sums=s, packed=est.packed)
        return est.exchange(s).cpu()
    for _ in range(warm):
        e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        s_last = e2e_step()
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

        def tstep(i):
            return vqmc.train_step_efficient(i, psi, h_fn, opt_update, opt_state, tparams, x_dev, 0.0, n_total=wl["n_walkers"])
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
    roofline = {"bound": "fp32", "kernel": f"wf::live_kernel<{D}, true> (wf_local_energy)", "achieved": achieved, "peak": fp32_peak,
                "unit": "TFLOP/s", "frac": achieved / fp32_peak,
                # dram__bytes_read.sum + dram__bytes_write.sum of one 65536-walker D=4 launch (ncu --set full,
                # profiles/r01_live_kernel_d4_lap_ncu_summary.txt); algorithmic bytes are 24 B/walker = 1.57 MB
                "traffic": 2376960 if (D == 4 and n_local == 65536) else None,
                "algorithmic_flops_per_walker": fl, "walkers_per_launch": n_local,
                "peak_source": "FFMA probe kernel (wf_probe_fma) timed in this run; MEASURED_PEAKS.json holds no FP32 figure",
                "note": "FMA-issue bound (SURVEY F8 / 8d): 24 B/walker of HBM traffic, so an HBM or tensor roofline does not "
                        "apply to this kernel; the HBM-bound operator is reported under 'spline_sweep'"}

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
        sweep = {"kernel": "spline_local_kernel<true> (wf_spline_apply_local)", "elements": M, "P": tabs.P, "ms": ms,
                 "elements_per_s": M / (ms * 1e-3),
                 "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                              # ncu --set full of this launch (profiles/r01_spline_local_ncu_summary.txt): 2.014 GB read + 0.197 GB written
                              "traffic": 2210474864, "algorithmic_bytes_per_element": bytes_per_el, "peak_source": hbm_src,
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
                                                                  # ncu --set full at 2^22 elements (profiles/r01_rqs_k32_ncu_summary.txt:
                                                                  # 1.537 GB read + 0.035 GB written), scaled x4 to this launch
                                                                  "traffic": 4 * 1571992784 if not inv else None,
                                                                  "algorithmic_bytes_per_element": 4 * (3 * K + 2), "peak_source": hbm_src}}
        rqs_sweep = {"kernel": "rqs_kernel<32, true> (wf_rqs_apply)", "elements": M, "K": K, "inputs": "6.6 GB per launch (> L2)", **res}
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
            coupling[f"D{Dc}"] = {"samples": M, "K": K, "layers": L, "hidden": hidden, **r,
                                 "note": "MUFU/issue bound by construction (SURVEY 8d regime ii): 8D+4 bytes per sample"}
            del x

    # ---------------- BASELINE config 5: 64-D coupling flow, K = 64, width-512 conditioners on the tensor cores (2^22 samples)
    tc_sweep = None
    if world == 1 and not args.no_sweep:
        from waveflow_b200.flows.neural_splines import coupling_flow_tc, pack_fcnn_tc
        D5, K5, H5, L5, M5 = 64, 64, 512, 8, 1 << 22
        crng = np.random.Generator(np.random.PCG64(0))
        out_dim = (3 * K5 - 1) * D5 // 2
        parts = []
        for _ in range(2 * L5):
            gW = lambda a, b: torch.from_numpy((crng.standard_normal((a, b)) / np.sqrt(a)).astype(np.float32)).to(dev)
            z = lambda n: torch.zeros(n, device=dev)
            parts.append(pack_fcnn_tc([(gW(D5 // 2, H5), z(H5)), (), (gW(H5, H5), z(H5)), (), (gW(H5, out_dim), z(out_dim))], dev))
        w5 = torch.cat(parts).contiguous()
        del parts
        g5 = torch.Generator(device=dev); g5.manual_seed(0)
        x5 = torch.rand(M5, D5, device=dev, generator=g5) * 6 - 3
        coupling_flow_tc(w5, L5, x5[: 1 << 17], 3.0)                       # warm-up
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); y5, ld5 = coupling_flow_tc(w5, L5, x5, 3.0); b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        flops = 2.0 * L5 * 2 * (32 * H5 + H5 * H5 + H5 * out_dim) * M5            # 109.1 MFLOP / sample
        bf16 = peaks.get("bf16_tflops_sustained", 1400.0)
        useful = flops / (ms * 1e-3) / 1e12
        tc_sweep = {"kernel": "tc_rqs_kernel / tc_hidden_kernel (wf_rqs_coupling_flow_tc): tcgen05.mma kind::tf32, 3-pass hi/lo split",
                    "samples": M5, "D": D5, "K": K5, "hidden": H5, "layers": L5, "ms": ms, "samples_per_s": M5 / (ms * 1e-3),
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
        tfl = 3.0 * flops_per_walker(D) * n_local / (ms * 1e-3) / 1e12          # forward + ~2x for the reverse pass
        train = {"api": "vqmc.train_step_efficient(epoch, psi, h_fn, opt_update, opt_state, params, batch, running_average)",
                 "kernels": "wf_vqmc_loss_grad (layer-wise jets: linear / tanh / spline-head kernels, forward + reverse) + wf_adam_step",
                 "walkers_total": wl["n_walkers"], "walkers_per_gpu": n_local, "ms_per_step": ms,
                 "walkers_per_s": wl["n_walkers"] / (ms * 1e-3), "loss": train_loss,
                 "exchange": "none" if world == 1 else f"all-reduce of the flat gradient ({opt_state.flat.numel()} floats) + 32-byte loss sums per step",
                 "algorithmic_tflops_per_gpu": tfl, "frac_of_measured_fp32_fma": tfl / fp32_peak,
                 "gpu_launches_per_step": 76 * ((n_local + 65535) // 65536) + 1}
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


    # ---------------- CPU baseline (rank 0, N = 1): bounded sample of the same workload
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        n_sample = min(wl["n_walkers"], 8192)
        v, sec, thr = cpu_reference(wl, n_sample, reps=5)
        cpu = {"value": v, "unit": UNIT, "cores": thr, "kind": "port", "sample": f"first {n_sample} walkers of the workload, 5 passes, median",
               "seconds_per_pass": sec,
               "note": "vectorised torch-CPU float32 restatement of the reference's jax.hessian path (oracle/fast_cpu.py); JAX is "
                       "not installable in this image"}

    s = s_last.numpy()
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": wl["name"], "description": wl["desc"], "walkers_total": wl["n_walkers"],
                       "walkers_per_gpu": n_local, "parallelism": f"walker-sharded x{world}, 32-byte estimator all-reduce per step", "exchange": exchange_kind,
                       "l2": "flushed between timed steps (256 MiB fill outside the event brackets)" if flush is not None else "not flushed",
                       "timing": "CUDA events per step on the launch stream, summed over K steps, max over ranks"},
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(x_host.numel() * 4) * world,
                    "d2h_bytes_per_step": 32, "api": "utils.physics.construct_hamiltonian_function(psi, protons)(params, walkers)",
                    "ms_per_step": e2e_s / steps * 1e3},
            "gpu_launches": steps * (2 if (world > 1 and est.peer is not None) else 1),     # live_kernel (+ p2p_allreduce_kernel)
            "kernel_ms_per_step": kern_ms_mean, "wall_s_timed_region": t_wall,
            "roofline": roofline, "cpu_baseline": cpu, "spline_sweep": sweep, "rqs_sweep": rqs_sweep, "coupling_flow_sweep": coupling, "tc_coupling_flow_sweep": tc_sweep, "flow_scaling": flow_scaling, "train_step": train,
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
