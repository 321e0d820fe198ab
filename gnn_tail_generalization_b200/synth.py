"""Synthetic inputs of the benchmark configurations, generated on the device.

The graph is a Chung-Lu style power-law graph canonicalised the way the reference's trainer hands
graphs to the model (trainer_node_classification.py:655-658): symmetric, no duplicate edges, exactly
one self loop per node.  Same construction as the oracle's ``powerlaw_graph`` (inverse-CDF endpoint
draw with exponent gamma, ids permuted), but with the device RNG so that 10^8 edges take seconds.
"""
import torch


def powerlaw_graph(num_nodes, num_undirected, seed=0, gamma=2.5, device='cuda'):
    """Returns edge_index int64 [2, 2*num_undirected + num_nodes] on ``device`` (unsorted by design:
    [lo->hi | hi->lo | self loops], each block ordered by the packed (lo, hi) key)."""
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed)
    expo = 1.0 / (1.0 - 1.0 / (gamma - 1.0))          # inverse-CDF exponent: i = N * U^expo
    perm = torch.randperm(num_nodes, generator=g, device=dev)
    keys = None
    need = num_undirected
    while True:
        draw = int(need * 1.3) + 1024
        u = torch.rand(2, draw, generator=g, device=dev, dtype=torch.float64)
        ends = (num_nodes * u.pow_(expo)).long().clamp_(max=num_nodes - 1)
        del u
        ends = perm[ends]
        lo, hi = torch.minimum(ends[0], ends[1]), torch.maximum(ends[0], ends[1])
        del ends
        k = (lo * num_nodes + hi)[lo != hi]
        del lo, hi
        keys = torch.unique(k if keys is None else torch.cat([keys, k]))
        del k
        if keys.numel() >= num_undirected:
            break
        need = num_undirected - keys.numel()
    if keys.numel() > num_undirected:
        sel = torch.randperm(keys.numel(), generator=g, device=dev)[:num_undirected].sort().values
        keys = keys[sel]
        del sel
    lo, hi = keys // num_nodes, keys % num_nodes
    del keys
    loops = torch.arange(num_nodes, device=dev)
    return torch.stack([torch.cat([lo, hi, loops]), torch.cat([hi, lo, loops])])


def features(num_nodes, dim, seed=1, device='cuda', out=None):
    """X ~ N(0,1) fp32, filled in row blocks so that no second full-size temporary is needed."""
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed)
    x = out if out is not None else torch.empty(num_nodes, dim, dtype=torch.float32, device=dev)
    step = max(1, (1 << 26) // max(1, dim))
    for r in range(0, num_nodes, step):
        x[r:r + step].normal_(generator=g)
    return x


def labels(num_nodes, num_classes, seed=2, device='cuda'):
    g = torch.Generator(device=torch.device(device)).manual_seed(seed)
    return torch.randint(0, num_classes, (num_nodes,), generator=g, device=device)
