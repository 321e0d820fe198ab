"""Locates the kernel bindings whether this package is imported as
``gnn_tail_generalization_b200.GNN_model`` or, in drop-in mode, as a top-level ``GNN_model``."""
import os
import sys

if __package__ and '.' in __package__:
    from .. import graph, ops                      # noqa: F401
    from ..errors import DGLError                  # noqa: F401
else:
    _root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    if _root not in sys.path:
        sys.path.append(_root)
    from gnn_tail_generalization_b200 import graph, ops      # noqa: F401
    from gnn_tail_generalization_b200.errors import DGLError  # noqa: F401
