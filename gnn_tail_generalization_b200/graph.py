"""Device-resident graph handle: the replacement for the ``dgl.graph`` object the reference's layer
walks (GNN_model/GCN.py:92-94, 186-246).

Only the DGL surface the TeacherGNN path touches is mirrored: ``in_degrees``, ``out_degrees``,
``number_of_edges``, ``number_of_nodes``, ``local_scope`` (a no-op: the handle is immutable).
Everything is built on the device by ``cb_graph_create`` from the int64 ``edge_index`` tensor.
"""
import contextlib
import ctypes

import torch

from . import _cabi as C


class _DevArray:
    """Zero-copy view of library-owned device memory through __cuda_array_interface__."""

    def __init__(self, addr, n, typestr):
        self.__cuda_array_interface__ = {'shape': (n,), 'typestr': typestr, 'data': (addr, False), 'version': 2}


class GraphHandle:
    """CSR-by-destination + CSR-by-source of one graph (or one node slice of it) in HBM."""

    world = 1

    def __init__(self, edge_index, num_nodes, row_begin=0, row_end=None, hub_chunk=0, local_out_edges=None,
                 src_panels=1):
        """edge_index [2, E] int64 on the device.  By default it is the whole edge list and the handle keeps the
        edges of the rows [row_begin, row_end).  With ``local_out_edges`` (cb_graph_create_local) ``edge_index`` holds
        only the edges whose DESTINATION is owned and ``local_out_edges`` those whose SOURCE is owned (for a symmetric
        graph: the same list with its rows swapped) -- the full list of a 10^9-edge graph never exists.
        src_panels (1, 2, 4): group every row's stored neighbours by source panel (cb_graph_create_panelled) so that an
        aggregation can run as that many passes, each needing only the source rows of its panel; the grouping (hence
        the summation order) is the same for every slicing of the graph."""
        def check(t, what):
            if not (torch.is_tensor(t) and t.is_cuda):
                raise ValueError(f'{what} must be a CUDA tensor (this path has no CPU implementation)')
            if t.dim() != 2 or t.shape[0] != 2:
                raise ValueError(f'{what} must be [2, E], got {tuple(t.shape)}')
            return t.to(torch.int64).contiguous()
        ei = check(edge_index, 'edge_index')
        self.device = ei.device
        self.num_nodes = int(num_nodes)
        row_end = self.num_nodes if row_end is None else int(row_end)
        self._h = ctypes.c_void_p()
        self.src_panels = int(src_panels)
        with torch.cuda.device(self.device):
            if self.src_panels > 1:
                eo = check(local_out_edges, 'local_out_edges') if local_out_edges is not None else ei
                C.call('cb_graph_create_panelled', C.ptr(ei), ei.shape[1], C.ptr(eo), eo.shape[1], self.num_nodes,
                       int(row_begin), row_end, int(hub_chunk), self.src_panels, int(local_out_edges is None),
                       C.stream_ptr(self.device), ctypes.byref(self._h))
            elif local_out_edges is None:
                C.call('cb_graph_create_sliced', C.ptr(ei), ei.shape[1], self.num_nodes, int(row_begin), row_end,
                       int(hub_chunk), C.stream_ptr(self.device), ctypes.byref(self._h))
            else:
                eo = check(local_out_edges, 'local_out_edges')
                C.call('cb_graph_create_local', C.ptr(ei), ei.shape[1], C.ptr(eo), eo.shape[1], self.num_nodes,
                       int(row_begin), row_end, int(hub_chunk), C.stream_ptr(self.device), ctypes.byref(self._h))
        self.row_begin, self.row_end = int(row_begin), row_end
        self.rows = row_end - int(row_begin)
        self.num_edges = self._qi(C.Q_NUM_EDGES)
        self.num_edges_by_src = self._qi(C.Q_SRC_NUM_EDGES)
        self.has_zero_in_degree = bool(self._qi(C.Q_HAS_ZERO_IN_DEG))
        self.hub_chunk = self._qi(C.Q_HUB_CHUNK)
        self.num_hub_chunks = (self._qi(C.Q_DST_NUM_HUB_CHUNKS), self._qi(C.Q_SRC_NUM_HUB_CHUNKS))
        # degree^-1/2 vectors stay owned by the handle; these tensors alias them (read-only use)
        self.din_inv_sqrt = self._view(C.Q_DIN_INV_SQRT, self.rows, '<f4')
        self.dout_inv_sqrt = self._view(C.Q_DOUT_INV_SQRT, self.rows, '<f4')
        self._ws = {}

    # ---- C-ABI plumbing -------------------------------------------------------------------
    @property
    def handle(self):
        if not self._h:
            raise RuntimeError('graph handle already destroyed')
        return self._h

    def _qi(self, what):
        v = ctypes.c_int64()
        C.call('cb_graph_query', self.handle, what, ctypes.byref(v))
        return int(v.value)

    def _view(self, what, n, typestr):
        p = ctypes.c_void_p()
        C.call('cb_graph_query', self.handle, what, ctypes.byref(p))
        if n == 0:
            dt = {'<f4': torch.float32, '<i4': torch.int32, '<i8': torch.int64}[typestr]
            return torch.empty(0, dtype=dt, device=self.device)
        return torch.as_tensor(_DevArray(p.value, n, typestr), device=self.device)

    def workspace(self, side, d):
        """Scratch for the hub-chunk partial rows of one aggregation (cached per (side, d))."""
        need = int(C.lib().cb_graph_workspace_bytes(self.handle, side, d))
        if need == 0:
            return None, 0
        key = (side, d)
        buf = self._ws.get(key)
        if buf is None or buf.numel() < need:
            buf = torch.empty(need, dtype=torch.uint8, device=self.device)
            self._ws[key] = buf
        return buf, need

    def carry(self, d):
        """fp32 [rows, d] buffer through which the source-panel passes of one aggregation hand on the row sums."""
        buf = self._ws.get(('carry', d))
        if buf is None:
            buf = torch.empty((self.rows, d), dtype=torch.float32, device=self.device)
            self._ws[('carry', d)] = buf
        return buf

    def rowptr_exp(self, side=C.CB_BY_DST):
        """int64 [rows*src_panels+1]: offsets of every (row, source panel) group (None when src_panels == 1)."""
        if self.src_panels == 1:
            return None
        return self._view(C.Q_DST_ROWPTR_EXP if side == C.CB_BY_DST else C.Q_SRC_ROWPTR_EXP,
                          self.rows * self.src_panels + 1, '<i8').clone()

    def close(self):
        if getattr(self, '_h', None):
            self.din_inv_sqrt = self.dout_inv_sqrt = None
            C.lib().cb_graph_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- structure access (copies; for tests and diagnostics) -----------------------------------
    def csr(self, side=C.CB_BY_DST):
        """(rowptr int64 [rows+1], col int32 [E], perm int32 [E]) as fresh tensors."""
        q = (C.Q_DST_ROWPTR, C.Q_DST_COL, C.Q_DST_PERM) if side == C.CB_BY_DST else \
            (C.Q_SRC_ROWPTR, C.Q_SRC_COL, C.Q_SRC_PERM)
        e = self.num_edges if side == C.CB_BY_DST else self.num_edges_by_src
        return (self._view(q[0], self.rows + 1, '<i8').clone(), self._view(q[1], e, '<i4').clone(),
                self._view(q[2], e, '<i4').clone())

    # ---- the DGL surface GCNConv.forward uses (GCN.py:186-246) ------------------------------------
    def in_degrees(self):
        return self._view(C.Q_IN_DEGREE, self.rows, '<i4').to(torch.int64)

    def out_degrees(self):
        return self._view(C.Q_OUT_DEGREE, self.rows, '<i4').to(torch.int64)

    def number_of_edges(self):
        return self.num_edges

    def number_of_nodes(self):
        return self.num_nodes

    @contextlib.contextmanager
    def local_scope(self):
        yield self

    def to(self, device):
        if torch.device(device) != self.device and torch.device(device).index is not None:
            raise ValueError('a GraphHandle lives on the device it was built on')
        return self

    # ---- halo exchange hook: identity on a whole graph, overridden by the node-sliced handle -----
    def exchange(self, local_rows):
        """Returns the matrix holding every source row the owned rows gather from."""
        return local_rows

    def exchange_flags(self, local_flags):
        """Per-row byte flags of every rank's rows (identity on a whole graph)."""
        return local_flags

    def push_slot(self, side, d, dtype=torch.float32, passes=True):
        """Exchange buffer the producing kernel should write into (node-sliced graphs with peer pushes)."""
        return None

    def allreduce_sum(self, t):
        """Sum of a small tensor over the ranks sharing the graph (identity for a whole graph)."""
        return t


def graph_from_edge_index(edge_index, num_nodes, hub_chunk=0):
    return GraphHandle(edge_index, num_nodes, hub_chunk=hub_chunk)
