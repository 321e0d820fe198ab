def gcn_norm(*a, **k):
    raise RuntimeError('torch_geometric gcn_norm is a shim (GNN_model/drop_tricks.py computes it and the DGL layer '
                       'ignores the result; not reachable with graph_dropout tricks off)')
