"""Model classes of the GAS / IncAgg runtime, by the names the reference's ``main.py:190`` looks up
(``getattr(models, conf.model.name)``).  GAT and PNA_JK are not part of the propagation path that was
rebuilt (broken against the fork's runtime)."""
from .base import ScalableGNN
from .gcn import GCN
from .gcn2 import GCN2
from .appnp import APPNP
from .graphsage import GraphSAGE
from .pna import PNA, PNAConv

MODELS = {cls.__name__: cls for cls in (GCN, GCN2, APPNP, GraphSAGE, PNA)}

__all__ = ['ScalableGNN', 'PNAConv', 'MODELS'] + sorted(MODELS)
