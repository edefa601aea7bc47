import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def pytest_sessionfinish(session, exitstatus):
    """Write the per-check parity statistics (fraction of elements inside north_star's flat tolerance, see
    tests/util.py::assert_fp32_grade) to gpurun_out/parity_fractions.json and echo a one-line summary per check."""
    import json
    import os
    from tests import util
    if not util.PARITY_LOG:
        return
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    test = os.environ.get("PYTEST_CURRENT_TEST", "")
    (out / "parity_fractions.json").write_text(json.dumps(util.PARITY_LOG, indent=1))
    tr = session.config.pluginmanager.get_plugin("terminalreporter")
    if tr is not None:
        tr.write_line("")
        for r in util.PARITY_LOG:
            tr.write_line(f"parity {r['name']:<28s} n={r['n']:<9d} within {r['flat_tol']:.0e}: {r['frac_within_flat_tol']:.6f} "
                          f"(float32 restatement {r['frac_within_flat_tol_float32_restatement']:.6f})  max {r['max']:.2e}")
