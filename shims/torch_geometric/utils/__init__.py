"""The three edge_index helpers the trainer's load_data uses (trainer_node_classification.py:655-658,
utils.py:667-674), restated with torch; everything else the reference imports from here raises if called."""
import torch

from . import num_nodes  # noqa: F401


def remove_self_loops(edge_index, edge_attr=None):
    keep = edge_index[0] != edge_index[1]
    return edge_index[:, keep], (edge_attr[keep] if edge_attr is not None else None)


def add_self_loops(edge_index, edge_attr=None, fill_value=None, num_nodes=None):
    n = int(edge_index.max()) + 1 if num_nodes is None else num_nodes
    loops = torch.arange(n, dtype=edge_index.dtype, device=edge_index.device)
    return torch.cat([edge_index, torch.stack([loops, loops])], 1), None


def to_undirected(edge_index, edge_attr=None, num_nodes=None, reduce='add'):
    n = int(edge_index.max()) + 1 if num_nodes is None else num_nodes
    both = torch.cat([edge_index, edge_index.flip(0)], 1)
    key = torch.unique(both[0] * n + both[1])          # coalesce: sorted by (row, col), duplicates merged
    return torch.stack([torch.div(key, n, rounding_mode='floor'), key % n])


def _off_path(name):
    def f(*a, **k):
        raise RuntimeError(f'torch_geometric.utils.{name} is a shim (not on the TeacherGNN path)')
    return f


to_networkx = _off_path('to_networkx')
negative_sampling = _off_path('negative_sampling')
dropout_adj = _off_path('dropout_adj')
subgraph = _off_path('subgraph')
