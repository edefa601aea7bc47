"""T0: the product's table generator against the reference's shipped tables and the reference generator's digests."""
import hashlib
import json
from pathlib import Path

import numpy as np
import pytest

from waveflow_b200.splines import tablegen as tg

GOLD = Path(__file__).parent / "golden"


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def shipped():
    return np.load(GOLD / "ref_tables_deg5_k16.npz")


def test_I_tables_bit_exact(shipped):
    tab, _ = tg.build_I_tables(5, 16)
    for nd in range(4):
        assert np.array_equal(tab[nd], shipped[f"I_nd{nd}"]), f"I nd={nd}"


def test_B_tables_bit_exact_and_OB_close(shipped):
    b = tg.build_B_tables(5, 16)
    for nd in range(4):
        assert np.array_equal(b["b"][nd], shipped[f"B_nd{nd}"]), f"B nd={nd}"
        ref = shipped[f"OB_nd{nd}"]
        # the orthonormalised tables go through BLAS (pinv, matmul): reproducible to rounding, not bitwise
        assert np.abs(b["ob"][nd] - ref).max() <= 2e-14 * np.abs(ref).max() + 1e-13
    assert np.abs(b["b_to_ob"] - shipped["b_to_ob"]).max() < 1e-11
    assert np.abs(b["ob_to_b"] - shipped["ob_to_b"]).max() < 1e-13


def test_against_reference_generator_digests():
    meta = json.loads((GOLD / "golden_meta.json").read_text())["ref_generated"]
    samples = np.load(GOLD / "ref_generated_samples.npz")
    build = {"I": lambda k, n: tg.build_I_tables(k, n)[0], "M": lambda k, n: tg.build_M_tables(k, n)[0],
             "B": lambda k, n: tg.build_B_tables(k, n)["b"]}
    for key, info in meta.items():
        kind, deg, kn = key.split("_")
        tab = build[kind](int(deg[3:]), int(kn[1:]))
        assert list(tab.shape) == info["shape"]
        assert np.array_equal(tab[:, :, ::97], samples[key])
        for nd in range(4):
            assert sha(tab[nd]) == info["sha256"][nd], f"{key} nd={nd}"


def test_orthonormality_and_basis_change():
    b = tg.build_B_tables(6, 23)
    ob = b["ob"][0]
    assert np.abs(ob @ ob.T / ob.shape[1] - np.eye(ob.shape[0])).max() < 1e-12
    assert np.abs(b["ob_to_b"] @ b["b_to_ob"] - np.eye(ob.shape[0])).max() < 1e-10


def test_odd_number_of_B_bases_rejected():
    with pytest.raises(ValueError):
        tg.build_B_tables(5, 17)          # 17 + 5 - 1 = 21 bases: ortho_splines.py:59-63 exits


def test_cache_files_are_reference_compatible(tmp_path):
    tg.build_I_tables(3, 6, 50, str(tmp_path / "I"))
    tg.build_M_tables(3, 6, 50, str(tmp_path / "M"))
    tg.build_B_tables(3, 6, 50, str(tmp_path / "B"))
    assert (tmp_path / "I" / "degree_3_niknots_9_nmp_50_nd_0.npy").exists()      # isplines_jax.py:115
    assert (tmp_path / "M" / "degree_3_niknots_7_nmp_50_nd_3.npy").exists()      # msplines_jax.py:93
    assert (tmp_path / "B" / "ob_degree_3_niknots_9_nmp_50_nd_1.npy").exists()   # bsplines_jax.py:79-80
    assert (tmp_path / "B" / "degree_3_niknots_9_nmp_50_b_to_ob.npy").exists()
    again, _ = tg.build_I_tables(3, 6, 50, str(tmp_path / "I"))
    fresh, _ = tg.build_I_tables(3, 6, 50)
    assert np.array_equal(again, fresh)
