import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


# GPU tests of these files were written after round 1's GPU budget was spent (not yet run on a B200): they are ordered
# last so that, under -x, a surprise there cannot hide the rest of the suite.  Remove an entry once it has run green.
RUN_LAST = ("test_bce_head.py", "test_shard_io.py", "test_hardneg.py", "test_roc_two_tier.py")


def pytest_collection_modifyitems(config, items):
    items.sort(key=lambda it: os.path.basename(str(it.fspath)) in RUN_LAST)      # stable: everything else keeps its order
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
