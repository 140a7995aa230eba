from .mf import MF, MFEngine  # noqa: F401
from .torch_engine import ModelEngine, RowOptimizer  # noqa: F401
