"""beta_recsys_b200 -- the sm_100a hot path under beta_rec's MF / GMF / NeuMF /
LightGCN training engines: CUDA kernels behind a C ABI (csrc/, include/brs_b200.h)
plus the host-side mirror of the reference's engine interface (engines/)."""
__version__ = "0.1.0"

from . import _lib  # noqa: F401
from ._lib import BrsError  # noqa: F401
from .install import install, uninstall  # noqa: F401,E402
