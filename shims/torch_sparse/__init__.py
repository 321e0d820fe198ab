"""Import stand-in for torch_sparse (Label_propagation_model only; off the TeacherGNN path)."""


class SparseTensor:
    def __init__(self, *a, **k):
        raise RuntimeError('torch_sparse is a shim')


def matmul(*a, **k):
    raise RuntimeError('torch_sparse is a shim')
