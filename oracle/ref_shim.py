"""Import shim that lets the UNMODIFIED reference (``/root/reference``) import in
the build container.  TEST INFRASTRUCTURE ONLY; never imported by the product.

The reference needs six third-party modules that are absent here
(tensorboardX, GPUtil, munch, py7zr, mock, ray) and the numpy<1.24 aliases
``np.int / np.long / np.float`` (beta_rec/utils/alias_table.py:50,
beta_rec/utils/common_util.py:115).  None of them touches arithmetic
(SURVEY.md section 8c).  ``/root/reference`` does not exist on the GPU box, so this
module is only used by ``oracle/make_golden.py`` and by tests that skip when
the reference tree is missing.
"""
import os
import sys
import types

def _find_reference():
    """$BETA_REC_REFERENCE, else /root/reference (build container), else baseline/_ref (the git-ignored copy of the
    reference PACKAGE that travels to the GPU box; written by oracle/install_ref.sh)."""
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for cand in (os.environ.get("BETA_REC_REFERENCE"), "/root/reference", os.path.join(here, "baseline", "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "beta_rec")):
            return cand
    return os.environ.get("BETA_REC_REFERENCE", "/root/reference")


REFERENCE_ROOT = _find_reference()


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "beta_rec"))


class _Munch(dict):
    """dict with attribute access (enough of munch.Munch for the reference)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _munchify(x):
    if isinstance(x, dict):
        return _Munch({k: _munchify(v) for k, v in x.items()})
    if isinstance(x, (list, tuple)):
        return type(x)(_munchify(v) for v in x)
    return x


def install():
    """Register stub modules and put the reference on sys.path (idempotent)."""
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    import numpy as np

    for name, val in (("int", int), ("long", np.int64), ("float", float), ("bool", bool)):
        if name not in np.__dict__:
            setattr(np, name, val)

    def stub(name, **attrs):
        if name in sys.modules:
            return sys.modules[name]
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class SummaryWriter(object):
        def __init__(self, *a, **k):
            pass

        def add_scalar(self, *a, **k):
            pass

        def add_scalars(self, *a, **k):
            pass

        def add_text(self, *a, **k):
            pass

        def close(self):
            pass

    stub("tensorboardX", SummaryWriter=SummaryWriter)
    stub("GPUtil", getAvailable=lambda *a, **k: [], getGPUs=lambda: [])
    stub("munch", munchify=_munchify, Munch=_Munch)
    stub("py7zr", unpack_7zarchive=lambda *a, **k: None)
    ray = stub("ray")
    tune = stub("ray.tune", grid_search=lambda v: v, report=lambda **k: None)
    ray.tune = tune
    ray.utils = stub("ray.utils")  # train_engine.py:156 assigns ray.utils.get_user_temp_dir
    import unittest.mock as _um

    sys.modules.setdefault("mock", _um)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
