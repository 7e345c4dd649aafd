import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "live: needs the live reference under /root/reference (dev container only)")


def pytest_collection_modifyitems(config, items):
    from oracle import live_reference

    if not live_reference.available():
        skip = pytest.mark.skip(reason="live reference (/root/reference) not present")
        for it in items:
            if "live" in it.keywords:
                it.add_marker(skip)
