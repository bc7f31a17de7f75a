"""incagg_gnn_b200 — B200-native implementation of IncAgg-GNN's per-batch propagation hot path
behind the reference's ``torch_geometric_autoscale`` Python API (reference ``__init__.py:20-33``).

``import incagg_gnn_b200 as tga`` from the repository root.  Importing the package loads the C-ABI CUDA library
``csrc/libincagg_b200.so`` and fails loudly when it has not been built: there is no CPU fallback.
"""
__version__ = '0.1.0'

from . import _lib  # noqa: F401  (raises ImportError if the CUDA extension is missing)
from . import ops
from .sparse import SparseTensor, spmm, spmm_delta  # noqa
from .data import Data  # noqa
from .synthetic import get_data, synthetic_graph, SHAPES  # noqa
from .history import History  # noqa
from .pool import AsyncIOPool  # noqa
from .metis import metis, permute  # noqa
from .preprocess import set_diag, gcn_norm, to_symmetric  # noqa
from .utils import compute_micro_f1, gen_masks, dropout, index2mask  # noqa
from .loader import SubgraphLoader, EvalSubgraphLoader, SubData  # noqa
from .models import ScalableGNN  # noqa
from . import models  # noqa

# the reference's plugin surface: torch.ops.torch_geometric_autoscale.{relabel_one_hop, ...}
ops.register_reference_ops()

__all__ = [
    'get_data', 'History', 'AsyncIOPool', 'metis', 'permute', 'compute_micro_f1', 'gen_masks',
    'dropout', 'SubgraphLoader', 'EvalSubgraphLoader', 'ScalableGNN', 'SparseTensor', 'Data',
    'set_diag', 'gcn_norm', 'to_symmetric', 'synthetic_graph', 'SHAPES', '__version__',
]
