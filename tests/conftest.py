import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "needs_reference: needs /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    skip_gpu = pytest.mark.skip(reason="no CUDA device")
    skip_ref = pytest.mark.skip(reason="/root/reference not present")
    has_ref = any(os.path.isdir(os.path.join(c, "beta_rec")) for c in
                  (os.environ.get("BETA_REC_REFERENCE") or "/nonexistent", "/root/reference", os.path.join(ROOT, "baseline", "_ref")))
    for it in items:
        if "gpu" in it.keywords and not has_gpu:
            it.add_marker(skip_gpu)
        if "needs_reference" in it.keywords and not has_ref:
            it.add_marker(skip_ref)
