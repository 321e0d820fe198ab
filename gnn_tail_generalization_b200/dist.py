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
import ctypes

import torch
import torch.distributed as dist

from . import _cabi as C
from .graph import GraphHandle, _DevArray

ROW_SHARDED_SUFFIXES = ('.le', 'embs')


PANEL_ROWS = 1 << C.CB_PANEL_SHIFT      # rows of one source-panel block (= the row tile of the producing GEMMs)


def rows_per_rank(num_nodes, world, align=1):
    """Rows of every rank but the trailing ones; ``align`` > 1 rounds it up to whole blocks of that many rows."""
    per = (num_nodes + world - 1) // world
    return -(-per // align) * align


def slice_bounds(num_nodes, world, rank, align=1):
    """[begin, end) of the nodes rank owns; trailing ranks may own fewer (or zero) rows."""
    per = rows_per_rank(num_nodes, world, align)
    lo = min(rank * per, num_nodes)
    return lo, min(lo + per, num_nodes)


def exchange_rows(local, num_nodes, world, group=None, per=None):
    """All-gather of equally sized row blocks -> the [num_nodes, d] matrix of every rank's rows.

    The last block is padded up to ``per`` (default ceil(N/P)) rows so that one all_gather_into_tensor moves
    everything; the returned tensor is the contiguous N-row prefix of the gathered buffer."""
    if world == 1:
        return local
    per = rows_per_rank(num_nodes, world) if per is None else per
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


def broadcast_dense_params(module, world, group=None, src=0):
    """Makes the replicated (dense) parameters identical on every rank.  With SE tables each rank draws
    ``randn(its_rows, d)`` between the layers' weight initialisations (GCN.py:177-182), so ranks that own a
    different number of rows leave the constructor with different dense weights."""
    if world == 1:
        return
    for n, p in module.named_parameters():
        if not is_row_sharded(n):
            dist.broadcast(p.data, src=src, group=group)


def need_masks(col_needed, num_nodes, world, rank, group=None, per=None):
    """Which local rows each peer gathers.

    col_needed: bool/uint8 [per*world] on this rank, 1 where this rank's CSR holds that global column id.
    Returns (mask uint8 [rows_per_rank]: bit j set = peer slot j gathers local row m, peers) where
    ``peers`` lists the remote ranks in slot order."""
    per = rows_per_rank(num_nodes, world) if per is None else per
    mine = col_needed.to(torch.uint8).contiguous()
    everyone = mine.new_empty((world, per * world))
    if world > 1:
        dist.all_gather_into_tensor(everyone, mine.reshape(1, -1), group=group)
    else:
        everyone[0] = mine
    peers = [r for r in range(world) if r != rank]
    lo = rank * per
    mask = torch.zeros(per, dtype=torch.uint8, device=mine.device)
    for j, r in enumerate(peers):
        mask |= everyone[r, lo:lo + per] << j
    return mask, peers


class PushSlot:
    """One use of an exchange buffer.  The kernel producing this rank's [rows, width] block writes it, one
    column panel per launch, into ``panel_local[p]`` (its rows of the [n_pad, panel_width] panel) and, through
    ``descs[p]`` (cb_peer_push_t), into the peers that gather those rows; ``pushed(p)`` then queues the stream
    barrier after which panel p is complete on every rank (``events[p]``).  Panels let the aggregation of
    panel p (HBM-bound, side stream) run while panel p+1 is still crossing NVLink.

    ``local`` is what the producer hands to autograd: [rows, width] for one panel, else the
    [rows, n_panels, panel_width] view of the panel-major buffer.

    Source-panel passes (``src_passes`` S > 1, graphs built with src_panels = S; one column panel): the producer is
    launched S times on the 128-row tiles of source panel 0, 1, ... (``descs[p]``.tile_first / tile_step), each
    launch followed by ``pushed(p)``; aggregation pass p (cb_agg_*_pass) needs only those rows, so it runs at FULL
    row width on the side stream while the tiles of panel p+1 are still being computed and pushed."""
    __slots__ = ('owner', 'width', 'local_rows', 'n_panels', 'panel_width', 'panel_full', 'panel_local', 'descs',
                 'local', 'events', 'pushed_rows', 'keep', 'dtype', 'src_passes')

    @property
    def n_launches(self):
        """Producer launches (= barriers = events) of this exchange."""
        return self.n_panels if self.src_passes == 1 else self.src_passes

    def pushed(self, p):
        dist.all_reduce(self.owner._flag, op=dist.ReduceOp.SUM, group=self.owner.group)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.owner.dev))
        self.events[p] = ev

    @property
    def side_stream(self):
        return self.owner.side_stream

    def rows(self, p, n):
        """Panel p of the exchanged matrix: [n, panel_width], every rank's rows."""
        return self.panel_full[p][:n]


class PeerExchange:
    """Peer-mapped exchange buffers of one rank (two, used alternately) and the per-row need masks.

    Why two buffers are enough: the push into buffer b of exchange #i starts after the stream barrier of
    exchange #i-1, which every rank enters only after its aggregation #i-2 -- the last reader of b -- is
    complete (the compute stream waits for the side stream that ran it)."""

    def __init__(self, graph, max_d, group=None, panels=1, push_ctas=0, elem_bytes=4, src_passes=True):
        """elem_bytes: bytes per element of the widest matrix exchanged (4: fp32 features, 2: a bf16-only model).
        src_passes: on a graph built with src_panels > 1, run every dense exchange as that many source-panel passes
        (False: the grouping of the neighbour lists is kept, the exchange is the one-launch / column-panel one)."""
        self.graph, self.group = graph, group
        self.src_passes = bool(src_passes) and graph.src_panels > 1
        self.world, self.rank = graph.world, graph.rank
        if not 2 <= self.world <= C.CB_MAX_PEERS + 1:
            raise ValueError(f'PeerExchange supports 2..{C.CB_MAX_PEERS + 1} ranks, got {self.world}')
        self.dev = graph.device
        self.per = graph.per
        self.n_pad = self.per * self.world
        self.max_d = int(max_d)
        self.panels, self.push_ctas = max(1, int(panels)), int(push_ctas)
        self.row_bytes = self.max_d * int(elem_bytes)
        nbytes = max(1, self.n_pad * self.row_bytes)
        self._mine, self._theirs, handles = [], [], []
        # Every rank goes through both collectives below whatever happened locally, so that a rank whose
        # allocation or mapping failed (no peer access, IPC disabled in the container ...) takes the whole
        # group to one clean error instead of leaving the others waiting in a collective.
        failure = None
        with torch.cuda.device(self.dev):
            try:
                for _ in range(2):
                    p = ctypes.c_void_p()
                    h = ctypes.create_string_buffer(C.CB_PEER_HANDLE_BYTES)
                    C.call('cb_peer_alloc', nbytes, ctypes.byref(p), h)
                    self._mine.append(p.value)
                    handles.append(bytes(h.raw))
            except Exception as e:      # noqa: BLE001 - reported to every rank below
                failure, handles = e, None
            everyone = [None] * self.world
            dist.all_gather_object(everyone, handles, group=group)
            self.peers = [r for r in range(self.world) if r != self.rank]
            if failure is None and all(h is not None for h in everyone):
                try:
                    for b in range(2):
                        ptrs = []
                        self._theirs.append(ptrs)
                        for r in self.peers:
                            q = ctypes.c_void_p()
                            C.call('cb_peer_open', ctypes.create_string_buffer(everyone[r][b], C.CB_PEER_HANDLE_BYTES),
                                   ctypes.byref(q))
                            ptrs.append(q.value)
                except Exception as e:  # noqa: BLE001
                    failure = e
            elif failure is None:
                failure = RuntimeError('a peer could not allocate its exchange buffers')
            ok = torch.tensor([0 if failure is not None else 1], dtype=torch.int32, device=self.dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok) == 0:
                self.close()
                raise RuntimeError(f'peer-mapped exchange buffers are not available on this node '
                                   f'(rank {self.rank}: {failure or "a peer failed"})')
            self.side_stream = torch.cuda.Stream(device=self.dev)
        # need masks: forward gathers sources (columns of the by-destination CSR), backward gathers
        # destinations (columns of the by-source CSR)
        self.need, self.pushed_rows = {}, {}
        for side in (C.CB_BY_DST, C.CB_BY_SRC):
            col = graph.csr(side)[1]
            needed = torch.zeros(self.n_pad, dtype=torch.uint8, device=self.dev)
            needed[col.long()] = 1
            mask, peers = need_masks(needed, graph.num_nodes, self.world, self.rank, group, per=self.per)
            assert peers == self.peers
            self.need[side] = mask
            rows = mask[:graph.rows].to(torch.int32)
            self.pushed_rows[side] = int(sum(((rows >> j) & 1).sum() for j in range(len(peers))))
        self._turn = 0
        self._flag = torch.zeros(1, dtype=torch.float32, device=self.dev)
        self._pending, self._views = {}, {}
        dist.barrier(group=group)   # every rank has mapped every buffer before the first push

    def slot(self, side, d, dtype=torch.float32, passes=True):
        """The next exchange buffer laid out for a [., d] matrix of ``dtype``; None if it does not fit.
        passes=False: one producer launch even on a source-panelled graph (the consumer gathers in one go)."""
        es = 4 if dtype == torch.float32 else 2
        vec = 16 // es                 # the aggregation reads rows with 16-byte accesses
        if dtype not in (torch.float32, torch.bfloat16) or d * es > self.row_bytes or d % vec:
            return None
        b = self._turn & 1
        self._turn += 1
        S = self.graph.src_panels if (passes and self.src_passes) else 1
        np_ = self.panels if (S == 1 and d % self.panels == 0 and (d // self.panels) % vec == 0) else 1
        pw = d // np_
        lo, rows = self.graph.row_begin, self.graph.rows
        s = PushSlot()
        s.owner, s.width, s.local_rows, s.n_panels, s.panel_width, s.dtype = self, d, rows, np_, pw, dtype
        s.src_passes = S
        full3 = self._views.get((b, d, np_, dtype))
        if full3 is None:
            if dtype == torch.float32:
                full3 = torch.as_tensor(_DevArray(self._mine[b], self.n_pad * d, '<f4'), device=self.dev)
            else:
                full3 = torch.as_tensor(_DevArray(self._mine[b], self.n_pad * d, '<i2'), device=self.dev).view(dtype)
            full3 = full3.view(np_, self.n_pad, pw)      # panel-major
            self._views[(b, d, np_, dtype)] = full3
        s.panel_full = [full3[p] for p in range(np_)]
        s.panel_local = [full3[p, lo:lo + rows] for p in range(np_)]
        s.local = s.panel_local[0] if np_ == 1 else full3[:, lo:lo + rows].permute(1, 0, 2)
        s.descs = []
        for p in range(np_ if S == 1 else S):
            desc = C.PeerPush()
            desc.n_peers = len(self.peers)
            desc.max_ctas = self.push_ctas if (np_ > 1 or S > 1) else 0
            for j, q in enumerate(self._theirs[b]):
                desc.peer[j] = q + (p * self.n_pad * pw * es if S == 1 else 0)
            desc.need = self.need[side].data_ptr()
            desc.row0, desc.ld = lo, pw
            if S > 1:
                # local tile t holds the rows of global block lo/128 + t (the slices are 128-row aligned), whose
                # source panel is that block index mod S
                desc.tile_first, desc.tile_step = (p - (lo >> C.CB_PANEL_SHIFT)) % S, S
            s.descs.append(desc)
        s.events = [None] * len(s.descs)
        s.pushed_rows, s.keep = self.pushed_rows[side], self.need[side]
        self._pending[s.local.data_ptr()] = s
        return s

    def take(self, local_rows):
        """The pending slot whose local view ``local_rows`` is (the producing kernel already pushed it)."""
        s = self._pending.pop(local_rows.data_ptr(), None)
        if s is not None and (local_rows.shape[0] != s.local_rows or local_rows.numel() != s.local_rows * s.width):
            s = None
        return s

    def close(self):
        lib = C.lib()
        for ptrs in self._theirs:
            for q in ptrs:
                lib.cb_peer_close(ctypes.c_void_p(q))
        self._theirs = []
        for p in self._mine:
            lib.cb_peer_free(ctypes.c_void_p(p))
        self._mine = []


class SlicedGraph(GraphHandle):
    """This rank's slice of the graph; ``exchange`` is the per-aggregation halo step."""

    def __init__(self, edge_index, num_nodes, rank, world, group=None, hub_chunk=0, local_out_edges=None,
                 src_panels=1):
        """edge_index: the whole edge list (every rank filters its slice), or -- with ``local_out_edges`` -- only this
        rank's in-edges, ``local_out_edges`` being its out-edges (see GraphHandle).
        src_panels > 1: neighbour lists grouped by source panel (GraphHandle) and slices of whole 128-row blocks, so
        that a row tile of the producing GEMM belongs to one panel (``slice_align``)."""
        self.slice_align = PANEL_ROWS if src_panels > 1 else 1
        self.per = rows_per_rank(num_nodes, world, self.slice_align)
        lo, hi = slice_bounds(num_nodes, world, rank, self.slice_align)
        super().__init__(edge_index, num_nodes, row_begin=lo, row_end=hi, hub_chunk=hub_chunk,
                         local_out_edges=local_out_edges, src_panels=src_panels)
        self.rank, self.world, self.group = rank, world, group
        self.exchanged_bytes = 0
        self.peer = None
        if world > 1 and dist.is_initialized():
            # the reference's check is global (GCN.py:187-197): every rank must raise together, or the ranks that
            # own no such node would walk into the exchange collectives alone
            flag = torch.tensor([int(self.has_zero_in_degree)], dtype=torch.int32, device=self.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
            self.has_zero_in_degree = bool(int(flag))

    def enable_push(self, max_d, panels=1, push_ctas=0, elem_bytes=4, src_passes=True):
        """Switch the exchange from an NCCL all-gather after the producing kernel to peer stores from
        inside it (PeerExchange).  ``max_d``: widest matrix that will be exchanged; ``panels`` > 1 pipelines
        the exchange by column panels against the aggregation, ``push_ctas`` caps the grid of a pushing
        kernel so that the aggregation of the previous panel finds free SMs; on a source-panelled graph the
        pipelining is by source panel at full row width instead (``src_passes``)."""
        if self.world > 1 and self.peer is None:
            self.peer = PeerExchange(self, max_d, self.group, panels, push_ctas, elem_bytes, src_passes)
        return self.peer

    def push_slot(self, side, d, dtype=torch.float32, passes=True):
        return self.peer.slot(side, d, dtype, passes) if self.peer is not None else None

    def exchange(self, local_rows):
        """[N, d] tensor of every rank's rows, or -- when the producer pushed them panel by panel -- the
        PushSlot whose panels become valid at its events."""
        s = self.peer.take(local_rows) if self.peer is not None else None
        if s is not None:
            self.exchanged_bytes += s.pushed_rows * s.width * (4 if s.dtype == torch.float32 else 2)
            # one launch: its barrier is already queued on this stream, stream order is enough
            return s.rows(0, self.num_nodes) if s.n_launches == 1 else s
        if local_rows.dim() == 3:
            local_rows = local_rows.reshape(local_rows.shape[0], -1)
        full = exchange_rows(local_rows, self.num_nodes, self.world, self.group, per=self.per)
        if self.world > 1:
            self.exchanged_bytes += (self.world - 1) * self.per * local_rows.shape[1] * local_rows.element_size()
        return full

    def exchange_flags(self, local_flags):
        if self.world == 1:
            return local_flags
        per = self.per
        send = local_flags
        if local_flags.shape[0] != per:
            send = local_flags.new_zeros(per)
            send[:local_flags.shape[0]] = local_flags
        full = local_flags.new_empty(per * self.world)
        dist.all_gather_into_tensor(full, send.contiguous(), group=self.group)
        return full[:self.num_nodes]

    def allreduce_sum(self, t):
        return allreduce_scalar_sum(t, self.world, self.group)


def attach_graph(teacher, graph):
    """Pre-seeds the layer stack's cached graph (the reference caches it in TricksComb.dglgraph,
    GCN.py:92-95) so that forward(x_local, edge_index=None) runs on this rank's slice."""
    teacher.model.model.dglgraph = graph
    return teacher
