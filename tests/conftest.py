import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long-running CPU oracle check")


def pytest_addoption(parser):
    parser.addoption("--runslow", action="store_true", default=False, help="also run the multi-minute oracle goldens")


def pytest_collection_modifyitems(config, items):
    # the slow oracle golden (SCnsIM body force, 500 steps) takes minutes on CPU; it was
    # run when the oracle was pinned (result quoted in DESIGN.md) and is opt-in: --runslow or IFEM_RUN_SLOW=1
    if config.getoption("--runslow") or os.environ.get("IFEM_RUN_SLOW") == "1":
        return
    skip = pytest.mark.skip(reason="slow oracle golden: use --runslow")
    for item in items:
        if "slow" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
