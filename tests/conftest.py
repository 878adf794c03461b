import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(GOLDEN, name))
    return load


def pytest_terminal_summary(terminalreporter):
    """Table of every log-mel comparison of the run: max-abs on well-conditioned cells, on the near-floor carve-out, and
    the carve-out's population (tests/parity.py)."""
    import parity
    if not parity.REPORT:
        return
    terminalreporter.write_line("log-mel parity (max-abs main | near-floor | carve-out cells / cells):")
    for what, main, floor, n_floor, cells in parity.REPORT:
        terminalreporter.write_line(f"  {what[:70]:70s} {main:.2e} | {floor:.2e} | {n_floor} / {cells}")
