"""CPU restatement of the reference's graph-preparation helpers  --  TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/`` and ``scripts/`` may import this file; the product (``gnn_tail_generalization_b200/graph_prep.py``)
runs the same functions as integer CUDA kernels through the C ABI (``cb_prep_*``) and never imports anything here.

Restated (paths under /root/reference): ``utils.py:300-334`` ``graph_analyze``, ``utils.py:667-674``
``ensure_symmetric``, ``utils.py:676-730`` ``save_graph_analyze``, ``utils.py:732-752`` ``craft_isolation_v2``,
``utils.py:910-943`` ``get_partial_sorted_idx`` -- same names, arguments, results and result ORDER, as plain tensor
programs.  Pinned: ``tests/test_graph_prep.py`` checks every function against fixtures produced by executing the
reference's OWN functions (``tests/golden/make_golden_prep.py`` -> ``tests/golden/prep_cases.npz``).
"""
import torch


def graph_analyze(N_nodes, edge_index):
    """(degs_ori, degs_dst): edges per node as origin / as destination (utils.py:300-334), int64 tensors on the
    device of ``edge_index``."""
    ori, dst = edge_index[0].long(), edge_index[1].long()
    return torch.bincount(ori, minlength=N_nodes)[:N_nodes], torch.bincount(dst, minlength=N_nodes)[:N_nodes]


def ensure_symmetric(edge_index):
    """Coalesced indices of A + A^T (utils.py:667-674): every edge and its reverse once, sorted by (row, col)."""
    ei = edge_index.long()
    n = int(ei.max()) + 1 if ei.numel() else 0
    both = torch.cat([ei, ei.flip(0)], dim=1)
    key = torch.unique(both[0] * n + both[1])          # sorted: the order coalesce() yields
    return torch.stack([torch.div(key, n, rounding_mode='floor'), key % n]) if n else ei


def _np_median(v):
    """numpy's median (mean of the two middle values for an even count) of a 1-D tensor."""
    s = torch.sort(v.double()).values
    k = s.numel()
    return (s[(k - 1) // 2] + s[k // 2]) / 2


def get_partial_sorted_idx(arr, mode='top25'):
    """Indices of the smallest ('top*') / largest ('bottom*') share of ``arr`` by the reference's repeated-median
    rule (utils.py:910-943); ascending index order, like ``np.where``."""
    arr = torch.as_tensor(arr).reshape(-1)
    a = arr.double()
    top = 'top' in mode
    levels = {'50': 1, '25': 2, '12': 3, '6': 4, '3': 5}[mode.replace('top', '').replace('bottom', '')]
    idx = torch.arange(a.numel(), device=a.device)
    for _ in range(levels):
        m = _np_median(a[idx])
        idx = torch.nonzero(a <= m if top else a >= m).reshape(-1)
    return idx


def craft_isolation_v2(data):
    """Removes every non-self-loop edge touching a ``zero_deg_mask`` node, keeping the edge order
    (utils.py:732-752); sets ``data.edge_index_bkup`` and ``data.edge_index``."""
    ei = data.edge_index
    z = data.zero_deg_mask.to(ei.device)
    ori, dst = ei[0].long(), ei[1].long()
    drop = (ori != dst) & (z[ori] | z[dst])
    data.edge_index_bkup = ei
    data.edge_index = ei[:, ~drop]
    return int(drop.sum())


def save_graph_analyze(N_nodes, data, use_special_split):
    """Degree statistics and the head / tail / isolated node splits of utils.py:680-730 (without its plotting and
    its ``np.save`` side effect); returns the Table-1 statistics record."""
    data.N_nodes = N_nodes
    degs_ori, degs_dst = graph_analyze(N_nodes, data.edge_index)
    d = degs_ori.double()
    stats = [N_nodes, int(degs_ori.sum()), int(degs_ori.max()), float(d.mean()), float(_np_median(d)),
             float((degs_ori == 0).sum()) / N_nodes * 100]
    dev = data.x.device

    def mask_of(idx):
        m = torch.zeros(N_nodes, dtype=torch.bool, device=dev)
        m[idx.to(dev)] = True
        return m

    if not use_special_split:
        data.small_deg_idx = get_partial_sorted_idx(degs_dst, 'top3')
        data.large_deg_idx = get_partial_sorted_idx(degs_dst, 'bottom3')
        data.small_deg_mask, data.large_deg_mask = mask_of(data.small_deg_idx), mask_of(data.large_deg_idx)
    else:
        idx = get_partial_sorted_idx(degs_dst, 'top6')
        idx = idx[_np_argsort(degs_dst[idx])]
        half = idx.numel() // 2
        data.zero_deg_idx, data.small_deg_idx = idx[:half], idx[half:]
        data.large_deg_idx = get_partial_sorted_idx(degs_dst, 'bottom3')
        data.zero_deg_mask, data.small_deg_mask = mask_of(data.zero_deg_idx), mask_of(data.small_deg_idx)
        data.large_deg_mask = mask_of(data.large_deg_idx)
        craft_isolation_v2(data)
    return stats


def _np_argsort(v):
    """numpy's default argsort (introsort, not stable, SIMD-dispatched) orders equal keys in a way that depends on
    the numpy build, and the reference's split of the lowest-degree sixth into "isolated" and "small" halves
    (utils.py:702-706) inherits that.  A stable sort is used here: same node set, same degrees on each side, ties
    resolved by node id (tests/test_graph_prep.py checks exactly that against the reference's output)."""
    return torch.sort(v, stable=True).indices
