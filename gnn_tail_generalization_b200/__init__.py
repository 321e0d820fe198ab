"""gnn_tail_generalization_b200: B200-native TeacherGNN aggregation path of Cold Brew.

Layout
  csrc/ + libcoldbrew_b200.so   hand-written sm_100a kernels behind the C ABI of include/coldbrew_b200.h
  _cabi.py                      ctypes binding (fails loudly when the library is missing)
  graph.py, ops.py              graph handle and autograd bindings
  GNN_model/                    drop-in mirror of the reference's GNN_model package
  dist.py                       1-D node-sliced multi-GPU path
"""
from .errors import DGLError  # noqa: F401

__all__ = ['DGLError']
