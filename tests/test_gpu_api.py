"""The reference-shaped Python API (model_factory / flows / wavefunctions / physics / vqmc) on top of the C ABI."""
import numpy as np
import pytest
import torch

from oracle import fixtures as fx
from oracle import laplacian as olap
from oracle import live
from tests.util import assert_fp32_grade, relerr, to_torch_tree

pytestmark = pytest.mark.gpu


def test_published_checkpoint_through_model_factory(cuda):
    """get_waveflow_model + the reference's own parameter pytree (published pickle) -> psi on the published grid."""
    from waveflow_b200 import model_factory
    from waveflow_b200.utils import physics
    params, gold = fx.load_he_checkpoint()
    init = model_factory.get_waveflow_model(2, base_spline_degree=6, i_spline_degree=6, n_prior_internal_knots=23,
                                            n_i_internal_knots=23, i_spline_reg=0.05, i_spline_reverse_fun_tol=1e-6,
                                            n_flow_layers=3, box_size=10, cached_bases_root=None)
    p0, psi, log_pdf, sample = init(0, 2)
    # same pytree structure as the reference: [(), (nn, zero), (), ...], (nn, zero)
    assert [len(p) for p in p0[0]] == [len(p) for p in params[0]]
    assert tuple(p0[0][1][0][0][0].shape) == params[0][1][0][0][0].shape
    tp = to_torch_tree(params, cuda)
    y, x = np.meshgrid(np.linspace(-10, 10, 100), np.linspace(-10, 10, 100))
    c = np.stack([x, y], -1).reshape(-1, 2)
    inv = (c[:, 0] > c[:, 1]).astype(int)
    sc = torch.from_numpy(np.sort(c, -1).astype(np.float32)).to(cuda)
    z = psi(tp, sc).cpu().numpy() * (-1.0) ** inv
    assert np.abs(z - gold["psi_grid"]).max() < 3e-5
    lp = log_pdf(tp, sc).cpu().numpy()
    assert abs(np.exp(lp.astype(np.float64)).sum() * (20 / 99) ** 2 - 1.0) < 0.12            # test_waveflow.py:52
    # Hamiltonian (physics.construct_hamiltonian_function) -> [N, 1]
    h_fn = physics.construct_hamiltonian_function(psi, protons=physics.system_catalogue[1]["He"][0], n_space_dimensions=1, eps=0.0)
    xs = np.sort(gold["samples"], -1).astype(np.float32)
    hp = h_fn(tp, torch.from_numpy(xs).to(cuda))
    assert tuple(hp.shape) == (len(xs), 1)
    ref = olap.local_energy_bundle(fx.waveflow_model(2), params, xs.astype(np.float64), np.array([[0.0], [0.0]]))
    assert relerr(hp[:, 0].cpu().numpy(), ref["hpsi"]) < 1e-4
    # samples land in the box, sorted coordinates not required by the reference-mode inverse
    s = sample(7, tp, 512, device=cuda)
    assert tuple(s.shape) == (512, 2) and bool(torch.isfinite(s).all())


def test_layerwise_operator_path_equals_fused_path(cuda):
    """flows.Serial evaluates [Box, (IMADE, Reverse) x L] fused; the same layers called one by one go through the
    operator-boundary kernels (conditioner -> remove_bias -> enforce_bc -> spline apply)."""
    from waveflow_b200 import flows
    from waveflow_b200.model_factory import get_masked_transform
    layers = [flows.BoxTransformLayer(10.0, xu_coord_type="mean")]
    for _ in range(2):
        layers += [flows.IMADE(get_masked_transform(), spline_degree=6, n_internal_knots=23, spline_regularization=0.05,
                               reverse_fun_tol=1e-6, constraints_dict_left={0: 0}, constraints_dict_right={0: 1},
                               cached_bases_path_root=None), flows.Reverse()]
    params, direct, inverse = flows.Serial(*layers)(3, 3)
    assert direct.wf_spec is not None
    params = to_torch_tree([p if not len(p) else ([tuple(t.numpy() for t in l) if len(l) else () for l in p[0]], p[1].numpy())
                            for p in params], cuda)
    x = torch.from_numpy(np.sort(np.random.default_rng(0).uniform(-9, 9, (2000, 3)), -1).astype(np.float32)).to(cuda)
    uf, ldf = direct(params, x)
    ul, ldl = direct.wf_layerwise(params, x)
    assert relerr(uf.cpu().numpy(), ul.cpu().numpy(), 1.0) < 5e-6
    assert relerr(ldf.cpu().numpy(), ldl.cpu().numpy(), 1.0) < 2e-5
    # Serial.inverse_fun (reference mode) = layer-wise inverse functions in reverse order
    xf = inverse(params, uf)[0]
    xl = inverse.wf_layerwise(params, uf)[0]
    assert np.median(np.abs((xf - xl).cpu().numpy())) < 1e-4


def test_general_boundary_constraints_fall_back_to_operator_path(cuda):
    from waveflow_b200 import flows
    from waveflow_b200.model_factory import get_masked_transform
    m = fx.mflow_model(n_layers=1, bc_i_left={0: 0, 2: 0, 3: 0})          # tests/test_boundary_constraints.py:18-21
    layer = flows.IMADE(get_masked_transform(), spline_degree=5, n_internal_knots=23, spline_regularization=0.02,
                        constraints_dict_left={0: 0, 2: 0, 3: 0}, constraints_dict_right={0: 1.0}, cached_bases_path_root=None)
    params, direct, inverse = flows.Serial(layer, flows.Reverse())(0, 2)
    assert direct.wf_spec is None                                              # not fusible -> layer by layer
    net = fx.random_net(np.random.default_rng(0), 2, 28, scale=2.0)
    x = np.random.default_rng(1).uniform(0.05, 0.95, (1500, 2)).astype(np.float32)
    y, ld = direct([to_torch_tree(net, cuda), ()], torch.from_numpy(x).to(cuda))
    ry, rld = live.imade_direct(m, fx.cast_params(net, np.float64), x.astype(np.float64))
    assert relerr(y.cpu().numpy(), ry[:, ::-1], 1.0) < 1e-5
    assert relerr(ld.cpu().numpy(), rld, 1.0) < 2e-5


def test_mflow_get_model_and_energy_estimator(cuda):
    from waveflow_b200 import model_factory, vqmc
    from waveflow_b200.utils import physics
    init = model_factory.get_model(3, 5, 15, 23, 0.02, 1e-6, 3, {0: 0}, {0: 0}, {0: 0.0}, {0: 1.0}, cached_bases_root=None)
    params, log_pdf, sample = init(0, 2)
    m = fx.mflow_model()
    npar = fx.random_params(np.random.default_rng(3), m, scale=2.0)
    tp = to_torch_tree(npar, cuda)
    x = np.random.default_rng(4).uniform(0.03, 0.97, (3000, 2)).astype(np.float32)
    lp, u = log_pdf(tp, torch.from_numpy(x).to(cuda), return_sample=True)
    r64, u64 = live.log_pdf(m, npar, x.astype(np.float64), return_sample=True)
    r32 = live.log_pdf(m.cast(np.float32), fx.cast_params(npar, np.float32), x)
    assert_fp32_grade(lp.cpu().numpy(), r64, r32, 1e-5, 1.0, "MFlow.log_pdf")
    s = sample(11, tp, 256, device=cuda)
    assert tuple(s.shape) == (256, 2)
    # VQMC estimator over emulated shards == single call
    psi, log_pdf_w, sample_w, opt_state, opt_update, get_params = vqmc.create_train_state(10, 1e-4, n_particle=2, rng=0,
                                                                                         cached_bases_root=None)
    h_fn = physics.construct_hamiltonian_function(psi, protons=np.array([[0.0], [0.0]]), n_space_dimensions=1)
    hparams, gold = fx.load_he_checkpoint()
    est = vqmc.EnergyEstimator(h_fn, hparams, cuda)
    walkers = torch.from_numpy(np.sort(np.random.default_rng(5).uniform(-10, 10, (1000, 2)), -1).astype(np.float32)).to(cuda)
    full = est.estimate(walkers)
    sums = torch.zeros(4, dtype=torch.float64, device=cuda)
    for r in range(4):
        lo, hi = est.shard(1000, r, 4)
        est.local_sums(walkers[lo:hi].contiguous(), sums)
    s = sums.cpu().numpy()
    assert s[2] == 1000 and abs(s[0] / s[2] - full["energy"]) <= 1e-9 * abs(full["energy"])
    assert [est.shard(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    loss = vqmc.loss_fn_efficient(hparams, psi, h_fn, walkers)
    assert abs(float(loss) - full["energy"]) <= 1e-4 * abs(full["energy"])


def test_peer_memory_exchange_protocol_emulated_ranks(cuda):
    """The slot / flag / parity protocol of wf_p2p_allreduce_sums with ALL ranks emulated by the warps of one CTA in ONE launch
    (wf_p2p_allreduce_emulated): kernels that wait on one another must never be separate launches on one GPU
    (B200_PROFILING.md).  The real multi-GPU run is tools/p2p_test.py and bench.py --gpus N."""
    import ctypes as C
    from waveflow_b200._ffi import check, lib, ptr
    assert lib.wf_p2p_allreduce_buffer_bytes(0) == -1 and lib.wf_p2p_allreduce_buffer_bytes(17) == -1
    rng = np.random.default_rng(0)
    for world in (1, 2, 4, 8):
        nbytes = int(lib.wf_p2p_allreduce_buffer_bytes(world))
        assert nbytes == (2 * world * 8 + 8) * 8
        bufs = [torch.zeros(nbytes // 8, dtype=torch.float64, device=cuda) for _ in range(world)]
        ptrs = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device=cuda)
        outs = torch.zeros(world, 4, dtype=torch.float64, device=cuda)
        for step in range(1, 8):
            vals = torch.from_numpy(rng.standard_normal((world, 4)) * 1e3).to(cuda)
            check(lib.wf_p2p_allreduce_emulated(ptr(ptrs), world, C.c_uint64(step), ptr(vals), ptr(outs), -1, 0, C.c_void_p(0)),
                  "wf_p2p_allreduce_emulated")
            torch.cuda.synchronize()
            want = torch.zeros(4, dtype=torch.float64, device=cuda)
            for r in range(world):
                want = want + vals[r]                       # rank order, as the kernel adds
            for r in range(world):
                assert torch.equal(outs[r], want), (world, step, r)
        assert all(int(b.view(torch.int64)[2 * world * 8].item()) == 0 for b in bufs)
    # a peer that never arrives: the waiting rank times out (short timeout here), returns NaN, sets its sticky error word,
    # and poisons the NEXT exchange so the failure reaches the ranks that did not time out themselves
    world = 2
    nbytes = int(lib.wf_p2p_allreduce_buffer_bytes(world))
    bufs = [torch.zeros(nbytes // 8, dtype=torch.float64, device=cuda) for _ in range(world)]
    ptrs = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device=cuda)
    outs = torch.zeros(world, 4, dtype=torch.float64, device=cuda)
    vals = torch.ones(world, 4, dtype=torch.float64, device=cuda)
    check(lib.wf_p2p_allreduce_emulated(ptr(ptrs), world, C.c_uint64(1), ptr(vals), ptr(outs), 1, 2_000_000, C.c_void_p(0)))
    torch.cuda.synchronize()
    assert bool(torch.isnan(outs[0]).all()) and int(bufs[0].view(torch.int64)[2 * world * 8].item()) == 1
    check(lib.wf_p2p_allreduce_emulated(ptr(ptrs), world, C.c_uint64(2), ptr(vals), ptr(outs), -1, 2_000_000, C.c_void_p(0)))
    torch.cuda.synchronize()
    assert bool(torch.isnan(outs).all())                   # rank 0 is poisoned -> NaN everywhere
    # the production entry point, single rank: the sum is the local block; step 0 is reserved
    one = torch.zeros(int(lib.wf_p2p_allreduce_buffer_bytes(1)) // 8, dtype=torch.float64, device=cuda)
    p1 = torch.tensor([one.data_ptr()], dtype=torch.int64, device=cuda)
    v = torch.tensor([1.5, -2.0, 3.0, 4.25], dtype=torch.float64, device=cuda)
    o = torch.zeros(4, dtype=torch.float64, device=cuda)
    check(lib.wf_p2p_allreduce_sums(ptr(p1), 0, 1, C.c_uint64(1), ptr(v), ptr(o), C.c_void_p(0)), "wf_p2p_allreduce_sums")
    torch.cuda.synchronize()
    assert torch.equal(o, v)
    assert lib.wf_p2p_allreduce_sums(ptr(p1), 0, 1, C.c_uint64(0), ptr(v), ptr(o), C.c_void_p(0)) == -1


def test_vector_allreduce_single_rank_and_step_counter(cuda):
    """wf_p2p_allreduce_vec with world = 1 (the multi-rank run is tools/p2p_test.py under torchrun): the vector and the loss sums
    come back unchanged, the device-side step counter advances once per call, a captured graph replays."""
    import ctypes as C
    from waveflow_b200._ffi import check, lib, ptr
    n = 5000
    nbytes = int(lib.wf_p2p_allreduce_vec_buffer_bytes(1, n))
    assert nbytes > 2 * 6144 * 4 and lib.wf_p2p_allreduce_vec_buffer_bytes(0, n) == -1
    buf = torch.zeros((nbytes + 7) // 8, dtype=torch.float64, device=cuda)
    ptrs = torch.tensor([buf.data_ptr()], dtype=torch.int64, device=cuda)
    step = torch.ones(1, dtype=torch.int64, device=cuda)
    cnt = torch.zeros(1, dtype=torch.int32, device=cuda)
    v = torch.randn(n, device=cuda)
    keep = v.clone()
    s4 = torch.tensor([1.0, 2.0, 3.0, 4.5], dtype=torch.float64, device=cuda)
    so = torch.zeros(4, dtype=torch.float64, device=cuda)
    call = lambda: check(lib.wf_p2p_allreduce_vec(ptr(ptrs), 0, 1, C.c_uint64(0), ptr(step), ptr(v), n, ptr(s4), ptr(so), ptr(cnt),
                                                  C.c_void_p(torch.cuda.current_stream().cuda_stream)), "wf_p2p_allreduce_vec")
    for k in range(3):
        call()
        torch.cuda.synchronize()
        assert torch.equal(v, keep) and torch.equal(so, s4) and int(step.item()) == 2 + k and int(cnt.item()) == 0
    side = torch.cuda.Stream(device=cuda)
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        call()
    before = int(step.item())
    graph.replay(); graph.replay()
    torch.cuda.synchronize()
    assert int(step.item()) == before + 2 and torch.equal(v, keep)
    assert lib.wf_p2p_allreduce_vec(ptr(ptrs), 0, 1, C.c_uint64(0), None, ptr(v), n, None, None, None, C.c_void_p(0)) == -1   # no step
