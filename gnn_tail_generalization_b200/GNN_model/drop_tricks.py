"""Graph-dropout hook of the layer stack (mirror of GNN_model/drop_tricks.py:127-171, default path).

On the TeacherGNN path the result of this hook is ignored by the layers: ``TricksComb.forward`` reads
``new_adjs[i]`` but hands the cached full graph to every GCNConv (GCN.py:101,111,115).  Only the
no-dropout behaviour is provided; the sampling tricks (DropEdge/DropNode/FastGCN/LADIES) need
torch-geometric / torch-scatter and are outside the hot path.
"""
from torch import nn

_SAMPLERS = ('DropEdge', 'DropNode', 'FastGCN', 'LADIES')


class DroppedEdges(list):
    """A one-element list answers every index with that element (one edge set shared by all layers)."""

    def __getitem__(self, i):
        return super().__getitem__(0 if len(self) == 1 else i)


class DropoutTrick(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.type_trick = args.type_trick
        self.num_layers = args.num_layers
        self.layerwise_drop = getattr(args, 'layerwise_dropout', False)
        self.graph_dropout = None
        if any(s in self.type_trick for s in _SAMPLERS):
            raise NotImplementedError(
                f'type_trick={self.type_trick!r}: graph-sampling tricks are not part of the B200 TeacherGNN path '
                f'(their output is unused by the reference layers, GCN.py:111-115)')

    def forward(self, edge_index, edge_weight=None, adj_norm=False, num_nodes=-1):
        if adj_norm:
            raise NotImplementedError('adj_norm=True (PyG gcn_norm) is not used on the TeacherGNN path')
        return DroppedEdges([(edge_index, edge_weight)])
