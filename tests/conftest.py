import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: minutes-long regression (the RT60 testbench sweep); runs with EAR_RUN_SLOW=1")


def pytest_collection_modifyitems(config, items):
    if os.environ.get("EAR_RUN_SLOW"):
        return
    skip = pytest.mark.skip(reason="slow regression: set EAR_RUN_SLOW=1 (its last run is kept under profiles/)")
    for item in items:
        if "slow" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import binding
    binding.build()
    return binding.lib()
