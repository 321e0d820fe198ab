"""Import stand-in for torch_scatter (GNN_model/drop_tricks.py, diffusion_feature.py; off the TeacherGNN path)."""


def scatter_add(*a, **k):
    raise RuntimeError('torch_scatter is a shim')


def scatter(*a, **k):
    raise RuntimeError('torch_scatter is a shim')
