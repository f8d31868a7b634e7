import glob
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
    config.addinivalue_line("markers", "needs_reference: needs /root/reference (build container only)")


def golden_files(prefix):
    return sorted(glob.glob(os.path.join(GOLDEN, f"{prefix}_*.npz")))


def pytest_collection_modifyitems(config, items):
    ref_ok = os.path.isdir("/root/reference/dbgsom")
    skip_ref = pytest.mark.skip(reason="reference tree not present on this machine")
    for item in items:
        if "needs_reference" in item.keywords and not ref_ok:
            item.add_marker(skip_ref)
