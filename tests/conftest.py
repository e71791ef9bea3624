import os
import sys

import pytest

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs the live reference (/root/reference or the staged oracle/_ref)")


def pytest_collection_modifyitems(config, items):
    have_ref = (os.path.isdir("/root/reference/src/rbvfit")
                or os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "rbvfit", "vfit_mcmc.py")))
    skip_ref = pytest.mark.skip(reason="neither /root/reference nor oracle/_ref present on this box")
    for item in items:
        if "reference" in item.keywords and not have_ref:
            item.add_marker(skip_ref)
