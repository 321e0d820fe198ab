"""Graph preparation either side of the TeacherGNN path, on the device.

The reference prepares its graphs with O(E) Python loops over ``.tolist()``-ed edge lists and numpy calls on the host
(``utils.py:300-334`` ``graph_analyze``, ``utils.py:667-674`` ``ensure_symmetric``, ``utils.py:676-752``
``save_graph_analyze`` / ``craft_isolation_v2``, ``utils.py:910-943`` ``get_partial_sorted_idx``), which is what makes
ogbn-arxiv slow and the 10^8-edge configurations impossible through ``main.py``.  These are the same functions -- same
names, arguments, results and result ORDER -- on the integer kernels of ``csrc/cb_prep.cu`` (C ABI ``cb_prep_*``):
an atomic histogram pass, a 64-bit radix sort + adjacent-difference compaction, order statistics read off one sorted
copy, flag / scan / scatter compactions.  CUDA tensors only: there is no host implementation behind these names (the
CPU restatement the tests check them against lives in ``oracle/graph_prep_oracle.py``).
"""
import ctypes

import torch

from . import _cabi as C

_LEVELS = {'50': 1, '25': 2, '12': 3, '6': 4, '3': 5}


def _edges(edge_index, what='edge_index'):
    if not (torch.is_tensor(edge_index) and edge_index.is_cuda):
        raise ValueError(f'{what} must be a CUDA tensor (graph_prep has no CPU implementation)')
    if edge_index.dim() != 2 or edge_index.shape[0] != 2:
        raise ValueError(f'{what} must be [2, E], got {tuple(edge_index.shape)}')
    return edge_index.to(torch.int64).contiguous()


def _int_array(arr, what):
    arr = torch.as_tensor(arr)
    if not arr.is_cuda:
        raise ValueError(f'{what} must be a CUDA tensor (graph_prep has no CPU implementation)')
    if arr.is_floating_point() or arr.is_complex():
        raise TypeError(f'{what}: integer values expected (the reference passes degree counts), got {arr.dtype}')
    return arr.reshape(-1).to(torch.int64).contiguous()


def _ws(what, n, device):
    """Scratch of one cb_prep_* call, from torch's caching allocator (cudaMalloc / cudaFree of the sort buffers would
    cost more than the kernels)."""
    nbytes = int(C.lib().cb_prep_graph_workspace_bytes(what, int(n)))
    return torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device), nbytes


def graph_analyze(N_nodes, edge_index):
    """(degs_ori, degs_dst): edges per node as origin / as destination (utils.py:300-334), int64 tensors on the
    device of ``edge_index`` (cb_prep_degrees)."""
    ei = _edges(edge_index)
    n = int(N_nodes)
    ori = torch.empty(n, dtype=torch.int64, device=ei.device)
    dst = torch.empty(n, dtype=torch.int64, device=ei.device)
    with torch.cuda.device(ei.device):
        ws, nb = _ws(C.PREP_DEGREES, n, ei.device)
        C.call('cb_prep_degrees', C.ptr(ei), ei.shape[1], n, C.ptr(ori), C.ptr(dst), C.ptr(ws), nb,
               C.stream_ptr(ei.device))
    return ori, dst


def ensure_symmetric(edge_index):
    """Coalesced indices of A + A^T (utils.py:667-674): every edge and its reverse once, sorted by (row, col)
    (cb_prep_symmetrize)."""
    ei = _edges(edge_index)
    e = ei.shape[1]
    if e == 0:
        return ei
    out = torch.empty((2, 2 * e), dtype=torch.int64, device=ei.device)
    count = ctypes.c_int64()
    with torch.cuda.device(ei.device):
        ws, nb = _ws(C.PREP_SYMMETRIZE, e, ei.device)
        C.call('cb_prep_symmetrize', C.ptr(ei), e, C.ptr(out), ctypes.byref(count), C.ptr(ws), nb,
               C.stream_ptr(ei.device))
    return out[:, :count.value].contiguous()


def get_partial_sorted_idx(arr, mode='top25'):
    """Indices of the smallest ('top*') / largest ('bottom*') share of ``arr`` by the reference's repeated-median
    rule (utils.py:910-943); ascending index order, like ``np.where`` (cb_prep_partial_sorted_idx)."""
    a = _int_array(arr, 'arr')
    top = 'top' in mode
    levels = _LEVELS[mode.replace('top', '').replace('bottom', '')]
    idx = torch.empty(a.numel(), dtype=torch.int64, device=a.device)
    count = ctypes.c_int64()
    with torch.cuda.device(a.device):
        ws, nb = _ws(C.PREP_PARTIAL_SORTED_IDX, a.numel(), a.device)
        C.call('cb_prep_partial_sorted_idx', C.ptr(a), a.numel(), int(top), levels, C.ptr(idx), ctypes.byref(count),
               C.ptr(ws), nb, C.stream_ptr(a.device))
    return idx[:count.value]


def degree_stats(degs):
    """[N, sum, max, mean, median, % zeros] of a degree array: the Table-1 record of utils.py:676-678
    (cb_prep_degree_stats)."""
    d = _int_array(degs, 'degs')
    out = (ctypes.c_double * 6)()
    with torch.cuda.device(d.device):
        ws, nb = _ws(C.PREP_DEGREE_STATS, d.numel(), d.device)
        C.call('cb_prep_degree_stats', C.ptr(d), d.numel(), out, C.ptr(ws), nb, C.stream_ptr(d.device))
    return [int(out[0]), int(out[1]), int(out[2]), out[3], out[4], out[5]]


def sort_idx_by_value(arr, idx):
    """``idx[argsort(arr[idx])]`` with ties kept in the order of ``idx`` (cb_prep_sort_idx_by_value).

    numpy's default argsort (utils.py:703, introsort, SIMD-dispatched) orders equal keys in a way that depends on the
    numpy build, and the reference's split of the lowest-degree sixth into "isolated" and "small" halves inherits
    that; the stable order gives the same node set with the same degrees on each side (tests/test_graph_prep.py
    checks exactly that against the reference's output)."""
    a, i = _int_array(arr, 'arr'), _int_array(idx, 'idx')
    out = torch.empty_like(i)
    with torch.cuda.device(a.device):
        ws, nb = _ws(C.PREP_SORT_IDX_BY_VALUE, i.numel(), a.device)
        C.call('cb_prep_sort_idx_by_value', C.ptr(a), a.numel(), C.ptr(i), i.numel(), C.ptr(out), C.ptr(ws), nb,
               C.stream_ptr(a.device))
    return out


def mask_of(idx, N_nodes, device=None):
    """bool [N_nodes], True at ``idx`` (utils.py:694-697, 711-717; cb_prep_mask_from_idx)."""
    i = _int_array(idx, 'idx')
    m = torch.empty(int(N_nodes), dtype=torch.bool, device=i.device)
    with torch.cuda.device(i.device):
        ws, nb = _ws(C.PREP_MASK_FROM_IDX, 0, i.device)
        C.call('cb_prep_mask_from_idx', C.ptr(i), i.numel(), int(N_nodes), C.ptr(m), C.ptr(ws), nb,
               C.stream_ptr(i.device))
    return m if device is None else m.to(device)


def craft_isolation_v2(data):
    """Removes every non-self-loop edge touching a ``zero_deg_mask`` node, keeping the edge order
    (utils.py:732-752); sets ``data.edge_index_bkup`` and ``data.edge_index`` (cb_prep_drop_edges).  Returns the
    number of removed edges."""
    ei = _edges(data.edge_index, 'data.edge_index')
    z = data.zero_deg_mask.to(ei.device).to(torch.bool).contiguous()
    e = ei.shape[1]
    out = torch.empty((2, e), dtype=torch.int64, device=ei.device)
    kept = ctypes.c_int64()
    with torch.cuda.device(ei.device):
        ws, nb = _ws(C.PREP_DROP_EDGES, e, ei.device)
        C.call('cb_prep_drop_edges', C.ptr(ei), e, C.ptr(z), z.numel(), C.ptr(out), ctypes.byref(kept), C.ptr(ws), nb,
               C.stream_ptr(ei.device))
    data.edge_index_bkup = data.edge_index
    data.edge_index = out[:, :kept.value].contiguous()
    return e - kept.value


def save_graph_analyze(N_nodes, data, use_special_split):
    """Degree statistics and the head / tail / isolated node splits of utils.py:680-730 (without its plotting and
    its ``np.save`` side effect); returns the Table-1 statistics record."""
    data.N_nodes = N_nodes
    degs_ori, degs_dst = graph_analyze(N_nodes, data.edge_index)
    stats = degree_stats(degs_ori)
    dev = data.x.device
    if not use_special_split:
        data.small_deg_idx = get_partial_sorted_idx(degs_dst, 'top3')
        data.large_deg_idx = get_partial_sorted_idx(degs_dst, 'bottom3')
        data.small_deg_mask = mask_of(data.small_deg_idx, N_nodes, dev)
        data.large_deg_mask = mask_of(data.large_deg_idx, N_nodes, dev)
    else:
        idx = sort_idx_by_value(degs_dst, get_partial_sorted_idx(degs_dst, 'top6'))
        half = idx.numel() // 2
        data.zero_deg_idx, data.small_deg_idx = idx[:half], idx[half:]
        data.large_deg_idx = get_partial_sorted_idx(degs_dst, 'bottom3')
        data.zero_deg_mask = mask_of(data.zero_deg_idx, N_nodes, dev)
        data.small_deg_mask = mask_of(data.small_deg_idx, N_nodes, dev)
        data.large_deg_mask = mask_of(data.large_deg_idx, N_nodes, dev)
        craft_isolation_v2(data)
    return stats
