"""Drop-in `model_utils.{getModel, getOptimizer}` (reference model_utils.py:700-711 signature).

The driver script imports these names (pyg-mem-tgn.py:24).  In the reference they build the DGL
stack, whose node memory is a frozen constant and never updated; the hot path
BASELINE.json names -- aggregator, GRU memory, temporal attention of `modules/*` -- is the PyG
stack of pyg_model_utils.py, which is what is returned here.  `gnn_param` is
accepted for signature compatibility."""
from pyg_model_utils import getModel, getOptimizer  # noqa: F401
