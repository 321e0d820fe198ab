def maybe_num_nodes(edge_index, num_nodes=None):
    return int(edge_index.max()) + 1 if num_nodes is None else num_nodes
