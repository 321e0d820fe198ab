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


# ------------------------------------------------------------------------------------------------
# the same graph law generated shard by shard (BASELINE.json configs[4]: 50 M nodes / 10^9 edges never exist as
# one edge list on one device)
# ------------------------------------------------------------------------------------------------
def _route(keys, owner, world, group):
    """Sends keys[i] to rank owner[i]; returns the keys this rank received (unordered)."""
    import torch.distributed as dist
    if world == 1:
        return keys
    order = torch.argsort(owner)
    keys = keys[order]
    send = torch.bincount(owner, minlength=world)
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    out = keys.new_empty(int(recv.sum()))
    dist.all_to_all_single(out, keys, output_split_sizes=recv.tolist(), input_split_sizes=send.tolist(), group=group)
    return out


def powerlaw_graph_sharded(num_nodes, num_undirected, rank, world, seed=0, gamma=2.5, device='cuda', group=None):
    """This rank's in-edges of a Chung-Lu power-law graph canonicalised like ``powerlaw_graph`` (symmetric, no
    duplicates, one self loop per node): int64 [2, E_r] with every destination in the rank's row range of the 1-D
    node partition (dist.slice_bounds), ordered by (destination, source).  Swapping the two rows gives the rank's
    out-edges (the graph is symmetric).  The ranks draw disjoint shares of the undirected pairs, de-duplicate them
    at the owner of the smaller endpoint and route every directed edge to the owner of its destination; total
    directed edges = 2 * (undirected kept) + num_nodes, within a few edges of 2 * num_undirected + num_nodes.

    Needs an initialised process group when world > 1 (NCCL on the device, or gloo on the CPU for tests)."""
    import torch.distributed as dist
    from .dist import rows_per_rank, slice_bounds
    dev = torch.device(device)
    per = rows_per_rank(num_nodes, world)
    lo_row, hi_row = slice_bounds(num_nodes, world, rank)
    expo = 1.0 / (1.0 - 1.0 / (gamma - 1.0))
    # the id permutation is a property of the graph: same on every rank
    perm = torch.randperm(num_nodes, generator=torch.Generator(device=dev).manual_seed(seed), device=dev)
    g = torch.Generator(device=dev).manual_seed(seed * 1000003 + 7919 * (rank + 1))
    share = (num_undirected + world - 1) // world
    keys = None
    need = share
    for _ in range(8):
        draw = int(need * 1.3) + 1024
        u = torch.rand(2, draw, generator=g, device=dev, dtype=torch.float64)
        ends = (num_nodes * u.pow_(expo)).long().clamp_(max=num_nodes - 1)
        del u
        ends = perm[ends]
        a, b = torch.minimum(ends[0], ends[1]), torch.maximum(ends[0], ends[1])
        del ends
        k = (a * num_nodes + b)[a != b]
        own = torch.div(k, num_nodes, rounding_mode='floor') // per
        del a, b
        got = _route(k, own, world, group)               # undirected pairs, at the owner of their smaller endpoint
        del k, own
        keys = torch.unique(got if keys is None else torch.cat([keys, got]))
        del got
        total = torch.tensor([keys.numel()], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(total, group=group)
        total = int(total)
        if total >= num_undirected:
            break
        need = (num_undirected - total + world - 1) // world
    if total > num_undirected:
        # every rank keeps the same fraction of its pairs: the global count lands within `world` of the target
        keep = min(keys.numel(), int(round(keys.numel() * (num_undirected / total))))
        sel = torch.randperm(keys.numel(), generator=g, device=dev)[:keep].sort().values
        keys = keys[sel]
        del sel
    a, b = torch.div(keys, num_nodes, rounding_mode='floor'), keys % num_nodes     # a < b, a owned by this rank
    del keys
    # (b -> a) stays here; (a -> b) goes to the owner of b.  In-edge key = dst * N + src.
    mine = a * num_nodes + b
    theirs = _route(b * num_nodes + a, b // per, world, group)
    del a, b
    loops = torch.arange(lo_row, hi_row, device=dev)
    key = torch.cat([mine, theirs, loops * num_nodes + loops]).sort().values
    del mine, theirs, loops
    dst, src = torch.div(key, num_nodes, rounding_mode='floor'), key % num_nodes
    return torch.stack([src, dst])
