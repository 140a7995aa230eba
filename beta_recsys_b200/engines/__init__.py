from .mf import MF, MFEngine  # noqa: F401
from .ncf import GMF, MLP, GMFEngine, MLPEngine, NeuMF, NeuMFEngine  # noqa: F401
from .torch_engine import ModelEngine, RowOptimizer  # noqa: F401
from .lightgcn import LightGCN, LightGCNEngine  # noqa: F401
