"""1-D node-sliced multi-GPU path (one process per GPU, torch.distributed).

The reference is single-device; the only behaviour to preserve is "same numbers as one
GPU".  Rank r owns the node range [r*ceil(N/P), (r+1)*ceil(N/P)): its rows of X, of every hidden
matrix, of the SE tables and of the outputs, all in-edges of those nodes (forward CSR) and all their
out-edges (backward CSR).  Dense weights are replicated.

Per aggregation there is exactly one exchange step: every rank needs the source rows its owned rows
gather from.  ``SlicedGraph.exchange`` provides them (NCCL all-gather of the row blocks over NVLink);
forward exchanges H, backward exchanges G (ops._FusedAggregate), so no reduce-scatter is needed.
Dense-parameter gradients are summed with one bucketed all-reduce; row-sharded parameters (``le``,
``embs``) stay local.

The functions that do not touch CUDA (bounds, exchange, gradient reduction, loss normalisation) also
run on the gloo backend, which is how the CPU test-suite covers the world_size=2 path.
"""
import torch
import torch.distributed as dist

from .graph import GraphHandle

ROW_SHARDED_SUFFIXES = ('.le', 'embs')


def rows_per_rank(num_nodes, world):
    return (num_nodes + world - 1) // world


def slice_bounds(num_nodes, world, rank):
    """[begin, end) of the nodes rank owns; trailing ranks may own fewer (or zero) rows."""
    per = rows_per_rank(num_nodes, world)
    lo = min(rank * per, num_nodes)
    return lo, min(lo + per, num_nodes)


def exchange_rows(local, num_nodes, world, group=None):
    """All-gather of equally sized row blocks -> the [num_nodes, d] matrix of every rank's rows.

    The last block is padded up to ceil(N/P) rows so that one all_gather_into_tensor moves everything;
    the returned tensor is the contiguous N-row prefix of the gathered buffer."""
    if world == 1:
        return local
    per = rows_per_rank(num_nodes, world)
    d = local.shape[1]
    send = local
    if local.shape[0] != per:
        send = local.new_zeros((per, d))
        send[:local.shape[0]] = local
    full = local.new_empty((per * world, d))
    dist.all_gather_into_tensor(full, send.contiguous(), group=group)
    return full[:num_nodes]


def allreduce_scalar_sum(t, world, group=None):
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def is_row_sharded(param_name):
    return param_name.endswith(ROW_SHARDED_SUFFIXES)


def allreduce_dense_grads(module, world, group=None):
    """Sum the gradients of the replicated (dense) parameters over ranks in one flat bucket."""
    if world == 1:
        return 0
    grads = [p.grad for n, p in module.named_parameters() if p.grad is not None and not is_row_sharded(n)]
    if not grads:
        return 0
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()
    return flat.numel()


class SlicedGraph(GraphHandle):
    """This rank's slice of the graph; ``exchange`` is the per-aggregation halo step."""

    def __init__(self, edge_index, num_nodes, rank, world, group=None, hub_chunk=0):
        lo, hi = slice_bounds(num_nodes, world, rank)
        super().__init__(edge_index, num_nodes, row_begin=lo, row_end=hi, hub_chunk=hub_chunk)
        self.rank, self.world, self.group = rank, world, group
        self.exchanged_bytes = 0

    def exchange(self, local_rows):
        full = exchange_rows(local_rows, self.num_nodes, self.world, self.group)
        if self.world > 1:
            self.exchanged_bytes += (self.world - 1) * rows_per_rank(self.num_nodes, self.world) * \
                local_rows.shape[1] * local_rows.element_size()
        return full

    def allreduce_sum(self, t):
        return allreduce_scalar_sum(t, self.world, self.group)


def attach_graph(teacher, graph):
    """Pre-seeds the layer stack's cached graph (the reference caches it in TricksComb.dglgraph,
    GCN.py:92-95) so that forward(x_local, edge_index=None) runs on this rank's slice."""
    teacher.model.model.dglgraph = graph
    return teacher
