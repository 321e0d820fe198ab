import torch


class Data:
    """Attribute bag with the two methods the trainer calls on it (``to``, ``num_nodes``)."""

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    @property
    def num_nodes(self):
        return self.x.shape[0]

    def to(self, device):
        for k, v in list(vars(self).items()):
            if torch.is_tensor(v):
                setattr(self, k, v.to(device))
        return self


class DataLoader:      # imported by utils.py:5, unused on the TeacherGNN path
    def __init__(self, *a, **k):
        raise RuntimeError('torch_geometric.data.DataLoader is a shim (not on the TeacherGNN path)')
