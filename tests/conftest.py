"""Shared fixtures. GPU tests are marked ``@pytest.mark.gpu`` and run only on a B200 box
(`pytest -m gpu`); everything else runs on CPU (`pytest -m "not gpu"`)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "link-prediction-gnn_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    return load


@pytest.fixture(scope="session")
def fb(golden):
    """fb-pages-food seed-0 split as produced by the reference (oracle/gen_golden.py)."""
    g = golden("fb_pages_food_seed0.npz")
    return {k: g[k] for k in g.files}


def pytest_sessionfinish(session, exitstatus):
    """Write the parity ledger (tests/helpers.parity): how many elements of every end-to-end tensor passed by which clause."""
    from helpers import LEDGER
    if LEDGER.rows:
        try:
            LEDGER.dump(os.environ.get("TWOWL_PARITY_REPORT", os.path.join(ROOT, "gpurun_out", "parity_report.json")))
        except OSError:
            pass


def pytest_terminal_summary(terminalreporter):
    from helpers import LEDGER
    if LEDGER.rows:
        t = LEDGER.totals()
        terminalreporter.write_line(f"parity ledger ({len(LEDGER.rows)} tensors): " + ", ".join(f"{k}={v}" for k, v in t.items()))
        for r in LEDGER.rows:
            if r["ref_error_clause"] or r["scale_floor"]:
                terminalreporter.write_line(f"  relaxed: {r}")
