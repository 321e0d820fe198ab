class NormalizeFeatures:
    """Row-normalises data.x to sum 1 (what torch_geometric.transforms.NormalizeFeatures does)."""

    def __call__(self, data):
        s = data.x.sum(dim=-1, keepdim=True).clamp(min=1.0)
        data.x = data.x / s
        return data


class ToSparseTensor:
    def __init__(self, *a, **k):
        raise RuntimeError('shim: ToSparseTensor is not on the TeacherGNN path')
