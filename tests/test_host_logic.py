"""Host-side logic that needs no GPU: parameter pytree ravel / unravel, weight packing (MADE masks, degree sort, prior fold)."""
import numpy as np
import torch

from oracle import fixtures as fx
from oracle import live
from tests.util import spec_from_live


def test_ravel_unravel_roundtrip_and_leaf_order():
    from waveflow_b200 import _train
    params, _ = fx.load_he_checkpoint()
    leaves = _train.tree_leaves(params)
    # jax.tree_util order: per conditioner W1, b1, W2, b2, W3, b3, zero_params; IMADE nets first, prior last
    assert [tuple(a.shape) for a in leaves[:7]] == [(2, 64), (64,), (64, 64), (64,), (64, 58), (58,), (2, 29)]
    assert len(leaves) == 28 and tuple(leaves[-3].shape) == (64, 56)
    flat = _train.ravel(params, torch.device("cpu"))
    assert flat.numel() == sum(int(np.prod(a.shape)) for a in leaves)
    tree = _train.unravel(params, flat)
    for a, b in zip(_train.tree_leaves(tree), leaves):
        assert np.array_equal(a.numpy(), np.asarray(b, dtype=np.float32))
    # views, not copies: an in-place update of the flat buffer is seen through the tree (what wf_adam_step relies on)
    flat += 1.0
    assert float(_train.tree_leaves(tree)[0][0, 0]) == float(np.float32(leaves[0][0, 0]) + 1.0)
    assert type(tree) is type(params) and type(tree[0]) is type(params[0])


def test_pack_net_masks_sort_and_prior_fold():
    """The packed conditioner reproduces the oracle's conditioner output (masks, degree-sorted hidden units, per-dimension
    regrouping), and the folded prior layer reproduces mask @ ob_to_b applied to that output (+ the sign-sum column)."""
    from waveflow_b200 import _live
    m = fx.waveflow_model(3)
    rng = np.random.default_rng(3)
    params = fx.random_params(rng, m)
    spec = spec_from_live(m)
    D, P = 3, m.P_P
    x = rng.uniform(0.05, 0.95, (7, D))
    nn, _ = params[1]
    (W1, b1), _, (W2, b2), _, (W3, b3) = nn
    m1, m2, m3 = live.made_masks(D)
    h = np.tanh(np.tanh(x @ (W1 * m1) + b1) @ (W2 * m2) + b2)
    o = (h @ (W3 * np.tile(m3, P)) + b3).reshape(-1, P, D).transpose(0, 2, 1)           # [n, d, p] raw conditioner output

    def run_packed(w):
        w = w.double().numpy()
        H = 64
        o1 = D * H; o2 = o1 + H; o3 = o2 + H * H; o4 = o3 + H; o5 = o4 + H * D * 32
        hh = np.tanh(np.tanh(x @ w[:o1].reshape(D, H) + w[o1:o2]) @ w[o2:o3].reshape(H, H) + w[o3:o4])
        return np.einsum("nh,hdq->ndq", hh, w[o4:o5].reshape(H, D, 32)) + w[o5:o5 + D * 32].reshape(D, 32)

    plain = run_packed(_live.pack_net(params[1], D, P, torch.device("cpu")))
    assert np.abs(plain[:, :, :P] - o).max() < 1e-5 and np.abs(plain[:, :, P:]).max() == 0
    packed = _live.pack_params(spec, params[0], params[1], torch.device("cpu"))
    assert packed.wf_folded
    folded = run_packed(packed[-_live._ffi.lib.wf_live_net_floats(D):])
    mask = np.ones(P); mask[0] = mask[-1] = 0
    want = (o * mask) @ np.asarray(spec.tab_P.ob_to_b64)
    assert np.abs(folded[:, :, :P] - want).max() < 1e-4 * np.abs(want).max()
    assert np.abs(folded[:, :, 31] - o.sum(-1)).max() < 1e-4 * np.abs(o.sum(-1)).max()
    assert not _live.pack_params(spec, params[0], params[1], torch.device("cpu"), fold_prior=False).wf_folded


def test_reference_arm_prints_one_json_line():
    """bench.py --impl reference (the CPU arm the driver runs next to ours): one JSON line with the contract's keys."""
    import json
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    r = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "vqmc_local_energy_walkers_per_s" and d["unit"] == "walkers/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["config"]["workload"] == "vqmc_c4"


def test_helpers_running_statistics():
    """utils/helpers.py:122-135: moving_average and the edge-padded sliding mean."""
    from waveflow_b200.utils import helpers
    assert helpers.moving_average(2.0, 4.0, 0.25) == 2.5
    x = np.arange(10, dtype=float) ** 2
    w = 4
    got = helpers.uniform_sliding_average(x, w)
    padded = np.concatenate([np.full(w - 1, x[0]), x])
    want = np.array([padded[i:i + w].mean() for i in range(len(x))])
    assert got.shape == x.shape and np.allclose(got, want)
    x2 = np.stack([x, 2 * x])
    assert np.allclose(helpers.uniform_sliding_average(x2, w)[1], 2 * want)


def test_pack_cache_keys_on_leaf_identity_and_version():
    """The reference signature carries the raw pytree on every call; the packed layout is rebuilt only when a leaf changed
    (new tensor, in-place torch update, or a raw-pointer update announced through bump_version)."""
    import torch
    from waveflow_b200 import _ffi
    a, b = torch.zeros(4), torch.ones(3)
    tree = ([(a, b), ()], (np.zeros(2),))
    k0 = _ffi.params_key(tree)
    assert _ffi.params_key(([(a, b), ()], tree[1])) == k0                  # same leaves, new containers
    cache, made = _ffi.PackCache(size=2), []
    make = lambda: made.append(1) or len(made)
    assert cache.get(k0, tree, make) == 1 and cache.get(_ffi.params_key(tree), tree, make) == 1
    a.add_(1.0)                                                             # in-place update through torch
    k1 = _ffi.params_key(tree)
    assert k1 != k0 and cache.get(k1, tree, make) == 2
    view = a[1:3]
    _ffi.bump_version(a)                                                    # raw-pointer update (wf_adam_step) announced by hand
    assert _ffi.params_key(tree) != k1 and view._version == a._version
    assert _ffi.params_key(([(a.clone(), b), ()], tree[1])) != _ffi.params_key(tree)
    for i in range(3):                                                      # LRU keeps `size` entries
        cache.get(("k", i), None, make)
    assert len(cache.items) == 2
