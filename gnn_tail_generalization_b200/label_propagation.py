"""Label propagation / outcome correlation on the gather kernel.

The reference's post-processing (Label_propagation_model/outcome_correlation.py:39-55, 128-156) iterates
``result = alpha * (adj @ result) + (1 - alpha) * y`` 50 times with ``adj`` a torch-sparse ``SparseTensor`` holding
``D^-1/2 A D^-1/2`` (or ``D^-1 A`` / ``A D^-1``) edge by edge.  The edge values are products of per-node factors, so the
product with ``adj`` is the unit-weight gather of this library between two row scalings:

    DAD x = d^-1/2 . A (d^-1/2 . x)        DA x = d^-1 . A x        AD x = A (d^-1 . x)

(``d`` = row sums of the adjacency, ``deg^p`` replaced by 0 where the degree is 0 -- outcome_correlation.py:47-48).
No edge-value array is stored or streamed: per iteration the kernel reads the column ids and the [N, c] matrix.
fp32 results agree with the edge-valued SpMM to rounding (the factors are applied per node instead of per edge).
"""
import torch

from . import _cabi as C, ops


def degree_factors(graph):
    """(d^-1/2, d^-1) with the reference's inf -> 0 rule; d = out-degree = row sums of ``SparseTensor(row, col)``."""
    deg = graph.out_degrees().to(torch.float32)
    dis = deg.pow(-0.5)
    dis[torch.isinf(dis)] = 0
    di = deg.pow(-1.0)
    di[torch.isinf(di)] = 0
    return dis, di


def normalized_adjacency_apply(graph, x, mode='DAD', factors=None):
    """``adj @ x`` for adj = DAD | DA | AD (gen_normalized_adjs, outcome_correlation.py:50-54) without materialising
    edge values.  x: float32 [N, c] on the graph's device."""
    dis, di = factors if factors is not None else degree_factors(graph)
    if mode == 'DAD':
        return ops.agg_gather_raw(graph, C.CB_BY_SRC, ops.row_scale_raw(x, dis), row_scale=dis)
    if mode == 'DA':
        return ops.agg_gather_raw(graph, C.CB_BY_SRC, x, row_scale=di)
    if mode == 'AD':
        return ops.agg_gather_raw(graph, C.CB_BY_SRC, ops.row_scale_raw(x, di))
    raise ValueError(f'unknown normalisation {mode!r}')


def propagate_step_raw(graph, x, y, row_scale, c_agg, c_y, clamp, out2_scale, want_out):
    """cb_agg_propagate: one fused iteration.  Returns (v or None, out2_scale * v or None)."""
    d = x.shape[1]
    out = torch.empty((graph.rows, d), dtype=torch.float32, device=x.device) if want_out else None
    out2 = torch.empty((graph.rows, d), dtype=torch.float32, device=x.device) if out2_scale is not None else None
    ws, ws_bytes = graph.workspace(C.CB_BY_SRC, d)
    lo, hi = clamp if clamp is not None else (0.0, 0.0)
    alg = ops.gather_alg_bytes(graph, C.CB_BY_SRC, d, int(want_out) + int(out2 is not None), 2, False,
                               int(row_scale is not None) + int(out2 is not None))
    with torch.cuda.device(x.device), ops._Timed('agg_propagate', alg, x.device):
        C.call('cb_agg_propagate', graph.handle, C.CB_BY_SRC, C.ptr(x), d, C.ptr(row_scale), C.ptr(y), float(c_agg),
               float(c_y), int(clamp is not None), float(lo), float(hi), C.ptr(out), C.ptr(out2_scale), C.ptr(out2),
               C.ptr(ws), ws_bytes, C.stream_ptr(x.device))
    return out, out2


def general_outcome_correlation(graph, y, alpha, num_propagations, post_step, alpha_term, mode='DAD', clamp=None):
    """outcome_correlation.py:128-147: ``res = alpha * adj @ res + ((1 - alpha) if alpha_term else 1) * y`` then
    ``post_step``, ``num_propagations`` times.

    post_step None (identity) or ``clamp=(lo, hi)``: ONE kernel per iteration (cb_agg_propagate) -- the source-side
    factor of adj rides on the iterate (the kernel also writes the pre-scaled copy the next step gathers), the
    destination-side factor, the axpy with y and the clamp sit in the gather's epilogue.  Any other ``post_step``
    callable keeps the iteration as gather + separate elementwise kernels."""
    factors = degree_factors(graph)
    dis, di = factors
    y = y.to(graph.device, torch.float32).contiguous()
    c_y = (1 - alpha) if alpha_term else 1.0
    if (post_step is None or clamp is not None) and num_propagations > 0:
        if mode not in ('DAD', 'DA', 'AD'):
            raise ValueError(f'unknown normalisation {mode!r}')
        pre = {'DAD': dis, 'DA': None, 'AD': di}[mode]        # factor carried by the gathered iterate
        post = {'DAD': dis, 'DA': di, 'AD': None}[mode]       # factor of the gathered sum
        it = ops.row_scale_raw(y, pre) if pre is not None else y
        result = None
        for k in range(num_propagations):
            last = k == num_propagations - 1
            result, scaled = propagate_step_raw(graph, it, y, post, alpha, c_y, clamp,
                                                pre if not last else None, want_out=last or pre is None)
            it = scaled if pre is not None else result
        return result
    result = y.clone()
    for _ in range(num_propagations):
        result = alpha * normalized_adjacency_apply(graph, result, mode, factors)
        result += c_y * y
        result = post_step(result) if post_step is not None else result
    return result


def label_propagation(graph, labels, label_idx, alpha, num_propagations, mode='DAD'):
    """outcome_correlation.py:149-158: propagate the one-hot training labels, clamped to [0, 1] after every step."""
    c = int(labels.max()) + 1
    y = torch.zeros((labels.shape[0], c), dtype=torch.float32, device=graph.device)
    y[label_idx] = torch.nn.functional.one_hot(labels[label_idx].reshape(-1), c).float()
    return general_outcome_correlation(graph, y, alpha, num_propagations, lambda t: torch.clamp(t, 0, 1), True, mode,
                                       clamp=(0.0, 1.0))
