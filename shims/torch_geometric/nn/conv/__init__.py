from . import gcn_conv  # noqa: F401


class MessagePassing:      # base class named by Label_propagation_model/LP_Adj.py:12 (off the TeacherGNN path)
    def __init__(self, *a, **k):
        raise RuntimeError('torch_geometric.nn.conv.MessagePassing is a shim')
