"""Autograd bindings of the C-ABI kernels: what the replacement GCNConv calls where the reference
called DGL (GNN_model/GCN.py:198-253) and elementwise PyTorch ops.

Every function here requires CUDA tensors and raises otherwise; there is no CPU fallback.
"""
import ctypes
import os

import torch

from . import _cabi as C


# ---------------------------------------------------------------------------------------------
# optional per-launch timing (bench.py's roofline probe): CUDA events on the launching stream
# ---------------------------------------------------------------------------------------------
_timing_sink = None


def set_timing_sink(sink):
    """``sink`` is a list that receives (name, start_event, end_event, algorithmic_bytes, flops) per kernel
    call, or None to switch the probe off.  Events are recorded on the stream the kernel runs on."""
    global _timing_sink
    _timing_sink = sink


class _Timed:
    def __init__(self, name, alg_bytes, device, flops=0):
        self.name, self.bytes, self.device, self.flops = name, alg_bytes, device, flops

    def __enter__(self):
        if _timing_sink is not None:
            self.t0 = torch.cuda.Event(enable_timing=True)
            self.t1 = torch.cuda.Event(enable_timing=True)
            self.t0.record(torch.cuda.current_stream(self.device))
        return self

    def __exit__(self, *exc):
        if _timing_sink is not None and exc[0] is None:
            self.t1.record(torch.cuda.current_stream(self.device))
            _timing_sink.append((self.name, self.t0, self.t1, self.bytes, self.flops))
        return False


def gather_alg_bytes(graph, side, d, n_out_mats=1, n_in_mats=1, mask=False, scales=1, elem=4):
    """Algorithmic bytes of one aggregation launch: every feature matrix once, the CSR once."""
    e = graph.num_edges if side == C.CB_BY_DST else graph.num_edges_by_src
    b = graph.num_nodes * d * elem                    # gathered matrix, read once
    b += (n_in_mats - 1) * graph.rows * d * elem      # x0
    b += n_out_mats * graph.rows * d * elem           # outputs
    b += graph.rows * d if mask else 0
    b += e * 4 + (graph.rows + 1) * 8 + scales * graph.rows * 4 + d * 4
    return b


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError('gnn_tail_generalization_b200 kernels need CUDA tensors (no CPU fallback)')


def _f32c(t):
    if t is None:
        return None
    if t.dtype != torch.float32:
        raise TypeError(f'expected float32, got {t.dtype}')
    return t.contiguous()


# ---------------------------------------------------------------------------------------------
# raw (non-differentiable) kernel calls
# ---------------------------------------------------------------------------------------------

def _featc(t, dtype):
    """Feature-matrix operand in the aggregation's storage type (fp32 or bf16), contiguous."""
    if t is None:
        return None
    if t.dtype != dtype:
        raise TypeError(f'expected {dtype}, got {t.dtype}')
    return t.contiguous()


_STORAGE = (torch.float32, torch.bfloat16)


def _sfx(dtype):
    """C-ABI name suffix of the storage-type variant of a kernel."""
    return '' if dtype == torch.float32 else '_bf16'


def _pofs(t, elems):
    """Device address of element ``elems`` of a tensor (None stays NULL)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr() + elems * t.element_size())


def agg_forward_raw(graph, H, bias=None, x0=None, alpha=0.0, relu=False, want_out=True, want_scaled=False,
                    want_mask=False, outs=None, panel=None, src_pass=None):
    """One fused forward aggregation over the owned rows.  H holds every source row ([N_global, d]).

    H fp32 -> cb_agg_forward; H bf16 -> cb_agg_forward_bf16 (x0 and the outputs are then bf16 too; the sums, the
    bias and the epilogue stay fp32).
    panel=(c0, w): H is the [N_global, w] column panel c0..c0+w of a wider matrix; bias / x0 / the outputs are
    the full-width [.., D] tensors (``outs`` = (out, out_scaled, mask) preallocated) and only their columns
    c0..c0+w are read / written.
    src_pass=p: source-panel pass p of a graph built with src_panels > 1 (cb_agg_forward_pass): only the neighbours of
    that panel are added, the row sums travel through ``graph.carry(d)``; the last pass writes the outputs (``outs``
    preallocated by the caller, shared by all passes)."""
    _need_cuda(H, bias, x0)
    if H.dtype not in (torch.float32, torch.bfloat16):
        raise TypeError(f'H must be float32 or bfloat16, got {H.dtype}')
    st = H.dtype
    H, bias, x0 = H.contiguous(), _f32c(bias), _featc(x0, st)
    d = H.shape[1]
    if H.shape[0] != graph.num_nodes:
        raise ValueError(f'H has {H.shape[0]} rows, the graph has {graph.num_nodes} nodes')
    c0, D = (panel[0], None) if panel is not None else (0, d)
    if outs is not None:
        out, out_scaled, mask = outs
        D = (out if out is not None else out_scaled).shape[1]
    else:
        if panel is not None:
            raise ValueError('a column panel needs preallocated full-width outputs')
        out = torch.empty((graph.rows, d), dtype=st, device=H.device) if want_out else None
        out_scaled = torch.empty((graph.rows, d), dtype=st, device=H.device) if want_scaled else None
        mask = torch.empty((graph.rows, d), dtype=torch.uint8, device=H.device) if want_mask else None
    ws, ws_bytes = graph.workspace(C.CB_BY_DST, d)
    es = H.element_size()
    alg = gather_alg_bytes(graph, C.CB_BY_DST, d, int(out is not None) + int(out_scaled is not None),
                           1 + int(x0 is not None), mask is not None, 1 + int(out_scaled is not None), es)
    fn, name = ('cb_agg_forward', 'agg_forward') if st == torch.float32 else ('cb_agg_forward_bf16', 'agg_forward_bf16')
    if src_pass is not None:
        if panel is not None:
            raise ValueError('source-panel passes work at full row width')
        carry = graph.carry(d)
        with torch.cuda.device(H.device), _Timed(name + '_pass', alg // graph.src_panels + 2 * graph.rows * d * 4, H.device):
            C.call('cb_agg_forward_pass', graph.handle, C.CB_F32 if st == torch.float32 else C.CB_BF16, C.ptr(H), d, d,
                   C.ptr(bias), C.ptr(x0), float(alpha), C.CB_ACT_RELU if relu else C.CB_ACT_NONE, C.ptr(out),
                   C.ptr(out_scaled), C.ptr(mask), D, int(src_pass), C.ptr(carry), C.ptr(ws), ws_bytes,
                   C.stream_ptr(H.device))
        return out, out_scaled, mask
    with torch.cuda.device(H.device), _Timed(name, alg, H.device):
        C.call(fn, graph.handle, C.ptr(H), d, d, _pofs(bias, c0), _pofs(x0, c0), float(alpha),
               C.CB_ACT_RELU if relu else C.CB_ACT_NONE, _pofs(out, c0), _pofs(out_scaled, c0), _pofs(mask, c0), D,
               C.ptr(ws), ws_bytes, C.stream_ptr(H.device))
    return out, out_scaled, mask


def compact_live_raw(graph, side, live):
    """cb_graph_compact_live: the CSR of ``side`` restricted to the columns with live[s] != 0, built in a workspace
    cached on the graph handle (one per side; every use is ordered on the calling stream).  Returns the workspace."""
    _need_cuda(live)
    if live.dtype != torch.uint8 or live.numel() != graph.num_nodes:
        raise ValueError('live must be uint8 [num_nodes]')
    live = live.contiguous()
    need = int(C.lib().cb_graph_live_workspace_bytes(graph.handle, side))
    ws = graph._ws.get(('live', side))
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=graph.device)
        graph._ws[('live', side)] = ws
    e = graph.num_edges if side == C.CB_BY_DST else graph.num_edges_by_src
    alg = 2 * e * 4 + e // 4 + 2 * (graph.rows + 1) * 8 + graph.num_nodes
    with torch.cuda.device(graph.device), _Timed('live_compact', alg, graph.device):
        C.call('cb_graph_compact_live', graph.handle, side, C.ptr(live), C.ptr(ws), need, C.stream_ptr(graph.device))
    return ws


def agg_gather_raw(graph, side, X, row_scale=None, out=None, panel=None, live=None, live_ws=None, flag_walk=False,
                   src_pass=None):
    """out[r] = row_scale[r] * sum_{j in row r} X[col[j]]  over one CSR side of the graph (X fp32 or bf16).
    panel=(c0, w): X is a [N_global, w] column panel; ``out`` is the preallocated full-width result.
    live: uint8 [N_global], 0 where the row of X is known to be all-zero (not gathered): the side is first
    compacted to the live columns (compact_live_raw) and the gather walks only those.  live_ws: an already
    compacted workspace (several panels share one).  flag_walk: keep the full walk and test the flag per column
    (the A/B partner of the compacted path; same sums)."""
    _need_cuda(X, row_scale)
    if X.dtype not in (torch.float32, torch.bfloat16):
        raise TypeError(f'X must be float32 or bfloat16, got {X.dtype}')
    st = X.dtype
    X, row_scale = X.contiguous(), _f32c(row_scale)
    d = X.shape[1]
    if X.shape[0] != graph.num_nodes:
        raise ValueError(f'X has {X.shape[0]} rows, the graph has {graph.num_nodes} nodes')
    c0 = panel[0] if panel is not None else 0
    if out is None:
        if panel is not None:
            raise ValueError('a column panel needs a preallocated full-width output')
        out = torch.empty((graph.rows, d), dtype=st, device=X.device)
    elif out.dtype != st:
        raise TypeError('out must have the storage type of X')
    D = out.shape[1]
    ws, ws_bytes = graph.workspace(side, d)
    alg = gather_alg_bytes(graph, side, d, 1, 1, False, int(row_scale is not None), X.element_size())
    if live is not None and (live.dtype != torch.uint8 or live.numel() != graph.num_nodes):
        raise ValueError('live must be uint8 [num_nodes]')
    sparse = live is not None or live_ws is not None
    name = ('agg_gather_dst' if side == C.CB_BY_DST else 'agg_gather_src') + ('_rowsparse' if sparse else '') + \
        ('' if st == torch.float32 else '_bf16')
    if src_pass is not None:
        if sparse or panel is not None:
            raise ValueError('source-panel passes: dense gathers at full row width only')
        carry = graph.carry(d)
        with torch.cuda.device(X.device), _Timed(name + '_pass', alg // graph.src_panels + 2 * graph.rows * d * 4, X.device):
            C.call('cb_agg_gather_pass', graph.handle, side, C.CB_F32 if st == torch.float32 else C.CB_BF16, C.ptr(X), d,
                   d, C.ptr(row_scale), C.ptr(out), D, int(src_pass), C.ptr(carry), C.ptr(ws), ws_bytes,
                   C.stream_ptr(X.device))
        return out
    if sparse and not flag_walk:
        if live_ws is None:
            live_ws = compact_live_raw(graph, side, live)
        # bytes that must move whatever the live fraction is: the output, the row offsets (compacted + original)
        alg = graph.rows * d * X.element_size() + 2 * (graph.rows + 1) * 8 + int(row_scale is not None) * graph.rows * 4
        with torch.cuda.device(X.device), _Timed(name, alg, X.device):
            C.call('cb_agg_gather_compacted', graph.handle, side, C.CB_F32 if st == torch.float32 else C.CB_BF16,
                   C.ptr(X), d, d, C.ptr(row_scale), C.ptr(live_ws), _pofs(out, c0), D, C.ptr(ws), ws_bytes,
                   C.stream_ptr(X.device))
        return out
    with torch.cuda.device(X.device), _Timed(name + ('_flagwalk' if sparse else ''), alg, X.device):
        C.call('cb_agg_gather' if st == torch.float32 else 'cb_agg_gather_bf16', graph.handle, side, C.ptr(X), d, d,
               C.ptr(row_scale), C.ptr(live), _pofs(out, c0), D, C.ptr(ws), ws_bytes, C.stream_ptr(X.device))
    return out


def backward_prep_raw(graph, d_out, d_out_scaled, mask, relu_out, relu, mixed, alpha, want_bias, want_x0,
                      d_x0_accum=None, drop_keep=None, drop_scale=1.0, want_live=False):
    """d_x0_accum: an existing [rows, d] buffer that receives ``+= alpha * dtot`` instead of a fresh d_x0.
    drop_keep / drop_scale: ``d_out`` is the gradient of dropout(out); the boolean keep mask of torch.native_dropout
    and 1 / (1 - p) fold its backward into this pass (cb_agg_backward_prep_ex).  want_live: also return uint8 [rows]
    flags of the rows of G that hold a non-zero (returned as a 4th value)."""
    ref = d_out if d_out is not None else d_out_scaled
    _need_cuda(ref)
    st = ref.dtype
    if st not in _STORAGE:
        raise TypeError(f'expected float32 or bfloat16, got {st}')
    d_out, d_out_scaled, relu_out = _featc(d_out, st), _featc(d_out_scaled, st), _featc(relu_out, st)
    rows, d = ref.shape
    G = torch.empty((rows, d), dtype=st, device=ref.device)
    d_bias = torch.empty(d, dtype=torch.float32, device=ref.device) if want_bias else None
    accumulate = int(want_x0 and d_x0_accum is not None)
    d_x0 = (d_x0_accum if accumulate else torch.empty((rows, d), dtype=st, device=ref.device)) \
        if want_x0 else None
    ws_bytes = int(C.lib().cb_prep_workspace_bytes(rows, d)) if want_bias else 0
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=ref.device) if ws_bytes else None
    mats = int(d_out is not None) + int(d_out_scaled is not None) + 1 + int(want_x0) + accumulate + \
        int(relu_out is not None)
    alg = mats * rows * d * ref.element_size() + (rows * d if mask is not None else 0) + 2 * rows * 4
    if drop_keep is not None or want_live:
        if drop_keep is not None:
            if drop_keep.dtype != torch.bool or drop_keep.shape != ref.shape or d_out is None or d_out_scaled is not None:
                raise ValueError('drop_keep: a bool mask of the plain output, whose gradient d_out must be the only one')
            drop_keep = drop_keep.contiguous()
            alg += rows * d
        live = torch.zeros(rows, dtype=torch.uint8, device=ref.device) if want_live else None
        with torch.cuda.device(ref.device), _Timed('backward_prep' + _sfx(st), alg, ref.device):
            C.call('cb_agg_backward_prep_ex', graph.handle, C.CB_F32 if st == torch.float32 else C.CB_BF16, C.ptr(d_out),
                   C.ptr(d_out_scaled), d, C.ptr(mask), C.ptr(relu_out), C.CB_ACT_RELU if relu else C.CB_ACT_NONE,
                   int(bool(mixed)), float(alpha), C.ptr(drop_keep), float(drop_scale), C.ptr(G), C.ptr(d_bias),
                   C.ptr(d_x0), accumulate, C.ptr(live), C.ptr(ws), ws_bytes, C.stream_ptr(ref.device))
        return (G, d_bias, d_x0, live) if want_live else (G, d_bias, d_x0)
    with torch.cuda.device(ref.device), _Timed('backward_prep' + _sfx(st), alg, ref.device):
        C.call('cb_agg_backward_prep' + _sfx(st), graph.handle, C.ptr(d_out), C.ptr(d_out_scaled), d, C.ptr(mask),
               C.ptr(relu_out), C.CB_ACT_RELU if relu else C.CB_ACT_NONE, int(bool(mixed)), float(alpha),
               C.ptr(G), C.ptr(d_bias), C.ptr(d_x0), accumulate, C.ptr(ws), ws_bytes, C.stream_ptr(ref.device))
    return G, d_bias, d_x0


def row_scale_raw(x, s):
    _need_cuda(x, s)
    x, s = _f32c(x), _f32c(s)
    if x.dim() != 2 or s.shape[0] != x.shape[0]:
        raise ValueError('row_scale: x must be [rows, d] and s [rows]')
    y = torch.empty_like(x)
    with torch.cuda.device(x.device), _Timed('row_scale', 2 * x.numel() * 4 + x.shape[0] * 4, x.device):
        C.call('cb_row_scale', C.ptr(x), C.ptr(s), x.shape[0], x.shape[1], C.ptr(y), C.stream_ptr(x.device))
    return y


def sumsq_raw(x):
    _need_cuda(x)
    x = _f32c(x)
    out = torch.empty(1, dtype=torch.float32, device=x.device)
    nb = int(C.lib().cb_sumsq_workspace_bytes())
    ws = torch.empty(nb, dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        C.call('cb_sumsq', C.ptr(x), x.numel(), C.ptr(out), C.ptr(ws), nb, C.stream_ptr(x.device))
    return out


class SplitWeight:
    """The weight operand of cb_gemm_rows: K-major [N, K] TF32 hi/lo halves (hi + lo ~= W to 2^-22), or -- for the
    bf16 transform -- one bf16 [N, K] matrix (``lo`` is None)."""

    __slots__ = ('hi', 'lo', 'n', 'k')

    def __init__(self, hi, lo):
        self.hi, self.lo = hi, lo
        self.n, self.k = hi.shape

    @property
    def dtype(self):
        return self.hi.dtype


def split_weight(W, transpose, dtype=torch.float32):
    """transpose=False: W is already [N, K] (nn.Linear.weight; or the GCNConv weight for the adjoint).
    transpose=True: W is [K, N] (GCNConv.weight used forward) and is transposed while splitting.
    dtype: storage type of the matrix the weight will multiply (fp32 master weights are rounded to bf16 per call)."""
    _need_cuda(W)
    W = _f32c(W.detach())
    n, k = (W.shape[1], W.shape[0]) if transpose else W.shape
    if dtype == torch.bfloat16:
        out = torch.empty((n, k), dtype=torch.bfloat16, device=W.device)
        with torch.cuda.device(W.device):
            C.call('cb_gemm_weight_to_bf16', C.ptr(W), n, k, int(bool(transpose)), C.ptr(out), C.stream_ptr(W.device))
        return SplitWeight(out, None)
    buf = torch.empty((2, n, k), dtype=torch.float32, device=W.device)
    with torch.cuda.device(W.device):
        C.call('cb_gemm_split_weight', C.ptr(W), n, k, int(bool(transpose)), C.ptr(buf[0]), C.ptr(buf[1]),
               C.stream_ptr(W.device))
    return SplitWeight(buf[0], buf[1])


def gemm_supported(M, N, K, dtype=torch.float32):
    """Shape test of cb_gemm_rows / cb_gemm_rows_grad.  M == 0 (a rank of a node-sliced graph that owns no rows)
    counts as supported: the raw wrappers then launch nothing but still take part in the exchange barriers."""
    if dtype not in _STORAGE:
        return False
    return bool(getattr(C.lib(), 'cb_gemm_rows_supported' + _sfx(dtype))(max(int(M), 1), int(N), int(K)))


def _weight_args(wt, c0, K):
    """The weight operand(s) of a rows GEMM starting at output column c0: (hi, lo) for fp32, (Bt,) for bf16."""
    if wt.lo is None:
        return (_pofs(wt.hi, c0 * K),)
    return (_pofs(wt.hi, c0 * K), _pofs(wt.lo, c0 * K))


def gemm_rows_raw(A, wt, row_scale=None, bias=None, add=None, relu=False, out2_scale=None, want_out=True,
                  want_out2=False, push=None, want_relu_mask=False):
    """act(row_scale * (A @ W^T) + bias + add) on the tcgen05 tensor cores: fp32 A -> 3xTF32 (fp32-class accuracy),
    bf16 A -> kind::f16 on the operands as stored (add / out / out2 then bf16 too, epilogue in fp32).
    Returns out, or (out, out2) when want_out2 (out2 = out2_scale[:,None] * out).

    push (dist.PushSlot, multi-GPU): ``out`` is written into the exchange buffer -- one launch per column
    panel of the slot, each storing its rows also into the peers that gather them and followed by the
    slot's stream barrier -- and the slot's local view is returned ([M, N], or [M, panels, N/panels]).
    want_relu_mask (no push): also returns, last, uint8 [M, N] with 1 where the activated output is positive
    (cb_gemm_rows_masked): the byte gate of the backward pass."""
    _need_cuda(A, row_scale, bias, add, out2_scale)
    st = A.dtype
    if st not in _STORAGE or wt.dtype != st:
        raise TypeError(f'A is {st}, the weight operand {wt.dtype}')
    A, row_scale, bias, add, out2_scale = A.contiguous(), _f32c(row_scale), _f32c(bias), _featc(add, st), _f32c(out2_scale)
    es = A.element_size()
    M, K = A.shape
    if K != wt.k:
        raise ValueError(f'A is [{M},{K}] but the weight operand is [{wt.n},{wt.k}]')
    N = wt.n
    if push is not None and (not want_out or push.width != N or push.local_rows != M or push.dtype != st):
        raise ValueError('push: the exchange slot does not match the kernel output')
    out = None
    if push is None and want_out:
        out = torch.empty((M, N), dtype=st, device=A.device)
    out2 = torch.empty((M, N), dtype=st, device=A.device) if want_out2 else None
    act = C.CB_ACT_RELU if relu else C.CB_ACT_NONE
    fn, mma = 'cb_gemm_rows' + _sfx(st), (6 if st == torch.float32 else 2)
    wbytes = (2 if wt.lo is not None else 1) * es
    if M == 0:
        if push is not None:      # a rank that owns no rows still takes part in every panel's barrier
            for p in range(push.n_launches):
                push.pushed(p)
            out = push.local
        return (out, out2) if want_out2 else out
    if push is not None and push.src_passes > 1:
        # source-panel passes: one launch per panel on that panel's 128-row tiles (desc.tile_first / tile_step), all
        # columns; the rows of panel p are complete everywhere after barrier p
        S = push.src_passes
        alg = (es * (M * K + M * N * (1 + int(want_out2) + int(add is not None)) + push.pushed_rows * N)) // S + wbytes * N * K
        for p in range(S):
            with torch.cuda.device(A.device), _Timed('gemm_rows_push' + _sfx(st), alg, A.device, flops=mma * M * N * K // S):
                C.call(fn, C.ptr(A), M, K, K, *_weight_args(wt, 0, K), N, C.ptr(row_scale), C.ptr(bias), C.ptr(add), N,
                       act, C.ptr(push.local), N, C.ptr(out2_scale), C.ptr(out2), N, ctypes.byref(push.descs[p]),
                       C.stream_ptr(A.device))
            push.pushed(p)
        return (push.local, out2) if want_out2 else push.local
    if want_relu_mask:
        if push is not None or M == 0:
            raise ValueError('want_relu_mask: a plain launch on a non-empty matrix only')
        rmask = torch.empty((M, N), dtype=torch.uint8, device=A.device)
        alg = es * (M * K + M * N * (int(want_out) + int(want_out2) + int(add is not None))) + wbytes * N * K + M * N
        wa = _weight_args(wt, 0, K)
        with torch.cuda.device(A.device), _Timed('gemm_rows' + _sfx(st), alg, A.device, flops=mma * M * N * K):
            C.call('cb_gemm_rows_masked', C.CB_F32 if st == torch.float32 else C.CB_BF16, C.ptr(A), M, K, K, wa[0],
                   wa[1] if len(wa) > 1 else None, N, C.ptr(row_scale), C.ptr(bias), C.ptr(add), N, act, C.ptr(out), N,
                   C.ptr(out2_scale), C.ptr(out2), N, C.ptr(rmask), N, None, C.stream_ptr(A.device))
        return (out, out2, rmask) if want_out2 else (out, rmask)
    if push is None:
        alg = es * (M * K + M * N * (int(want_out) + int(want_out2) + int(add is not None))) + wbytes * N * K
        with torch.cuda.device(A.device), _Timed('gemm_rows' + _sfx(st), alg, A.device, flops=mma * M * N * K):
            C.call(fn, C.ptr(A), M, K, K, *_weight_args(wt, 0, K), N, C.ptr(row_scale), C.ptr(bias),
                   C.ptr(add), N, act, C.ptr(out), N, C.ptr(out2_scale), C.ptr(out2), N, None,
                   C.stream_ptr(A.device))
        return (out, out2) if want_out2 else out
    pw = push.panel_width
    for p in range(push.n_panels):
        c0 = p * pw
        alg = es * (M * K + M * pw * (1 + int(want_out2) + int(add is not None)) + push.pushed_rows * pw) + wbytes * pw * K
        with torch.cuda.device(A.device), _Timed('gemm_rows_push' + _sfx(st), alg, A.device, flops=mma * M * pw * K):
            C.call(fn, C.ptr(A), M, K, K, *_weight_args(wt, c0, K), pw,
                   C.ptr(row_scale), _pofs(bias, c0), _pofs(add, c0), N, act, C.ptr(push.panel_local[p]), pw,
                   C.ptr(out2_scale), _pofs(out2, c0), N, ctypes.byref(push.descs[p]), C.stream_ptr(A.device))
        push.pushed(p)
    return (push.local, out2) if want_out2 else push.local


def row_any_nonzero_raw(x):
    """uint8 [rows]: 1 where the row of x (fp32 or bf16, [rows, d]) holds a non-zero element (cb_row_any_nonzero)."""
    _need_cuda(x)
    if x.dtype not in _STORAGE or x.dim() != 2:
        raise TypeError('row_any_nonzero: a 2-D float32 / bfloat16 matrix')
    x = x.contiguous()
    flags = torch.empty(x.shape[0], dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device), _Timed('row_any_nonzero', x.numel() * x.element_size() + x.shape[0], x.device):
        C.call('cb_row_any_nonzero', C.ptr(x), C.CB_F32 if x.dtype == torch.float32 else C.CB_BF16, x.shape[0],
               x.shape[1], x.shape[1], C.ptr(flags), C.stream_ptr(x.device))
    return flags


def gemm_rows_grad_raw(A, wt, row_scale=None, add=None, gate_u8=None, gate_f32=None, mixed=False, alpha=0.0,
                       d_x0=None, accumulate_x0=False, want_x0=False, post_scale=None, want_col_sum=False,
                       push=None, row_live=None, push_live=None, a_live=None, x0_valid=None):
    """cb_gemm_rows_grad: the adjoint GEMM with the backward prologue of the layer below in its epilogue.
    Returns (out, col_sum or None, d_x0 or None); with ``push`` the output goes to the exchange slot
    (see gemm_rows_raw) and its local view is returned.  row_live: zeroed uint8 [M] that receives 1 for
    every output row holding a non-zero element.  push_live: uint8 [M], 0 for rows known to come out all-zero
    (their A row is zero): those are not pushed to the peers.  gate_f32: the relu output in the storage type of A
    (gate = value > 0).  a_live: uint8 [M] from row_any_nonzero_raw(A): rows with 0 are skipped entirely -- their
    rows of ``out`` / ``d_x0`` are NOT written (the consumers go by the same flags).  x0_valid: uint8 [M] flags of an
    accumulating ``d_x0`` whose first writer skipped rows that way (0 = read as zero)."""
    _need_cuda(A, row_scale, add, gate_u8, gate_f32, d_x0, post_scale)
    st = A.dtype
    if st not in _STORAGE or wt.dtype != st:
        raise TypeError(f'A is {st}, the weight operand {wt.dtype}')
    A, row_scale, add, gate_f32, post_scale = A.contiguous(), _f32c(row_scale), _featc(add, st), _featc(gate_f32, st), \
        _f32c(post_scale)
    if d_x0 is not None and d_x0.dtype != st:
        raise TypeError('d_x0 must have the storage type of A')
    es = A.element_size()
    M, K = A.shape
    if K != wt.k:
        raise ValueError(f'A is [{M},{K}] but the weight operand is [{wt.n},{wt.k}]')
    N = wt.n
    if push is not None and (push.width != N or push.local_rows != M or push.dtype != st):
        raise ValueError('push: the exchange slot does not match the kernel output')
    out = torch.empty((M, N), dtype=st, device=A.device) if push is None else push.local
    col_sum = torch.empty(N, dtype=torch.float32, device=A.device) if want_col_sum else None
    if want_x0 and d_x0 is None:
        d_x0, accumulate_x0 = torch.empty((M, N), dtype=st, device=A.device), False
    gate = gate_u8 if gate_u8 is not None else gate_f32
    ws_bytes = int(C.lib().cb_gemm_rows_grad_workspace_bytes(M, N)) if want_col_sum else 0
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=A.device) if ws_bytes else None
    extra = int(add is not None) + int(gate_f32 is not None) + int(d_x0 is not None) * (1 + int(bool(accumulate_x0)))
    if M == 0:
        if push is not None:      # a rank that owns no rows still takes part in every panel's barrier
            for p in range(push.n_launches):
                push.pushed(p)
        return out, (col_sum.zero_() if col_sum is not None else None), d_x0
    S = push.src_passes if push is not None else 1
    if S > 1:       # source-panel passes: every launch covers all columns of its panel's row tiles
        panels = [(0, N, push.descs[p], out) for p in range(S)]
    elif push is None:
        panels = [(0, N, None, out)]
    else:
        panels = [(p * push.panel_width, push.panel_width, push.descs[p], push.panel_local[p]) for p in range(push.n_panels)]
    fn, mma = 'cb_gemm_rows_grad' + _sfx(st), (6 if st == torch.float32 else 2)
    wbytes = (2 if wt.lo is not None else 1) * es
    col_parts = []
    for p, (c0, w, desc, dst) in enumerate(panels):
        if desc is not None:
            desc.row_live = push_live.data_ptr() if push_live is not None else None
        alg = (es * (M * K + M * w * (1 + extra)) + (M * w if gate_u8 is not None else 0) +
               (push.pushed_rows * w * es if push is not None else 0)) // S + wbytes * w * K
        # the column sums of a launch cover its row tiles only: one [N] vector per pass, added below in pass order
        col_p = torch.empty(N, dtype=torch.float32, device=A.device) if (col_sum is not None and S > 1) else col_sum
        with torch.cuda.device(A.device), _Timed(('gemm_rows_grad_push' if push is not None else 'gemm_rows_grad') +
                                                 _sfx(st), alg, A.device, flops=mma * M * w * K // S):
            C.call(fn, C.ptr(A), M, K, K, *_weight_args(wt, c0, K), w,
                   C.ptr(row_scale), _pofs(add, c0), N, _pofs(gate_u8, c0), _pofs(gate_f32, c0),
                   N if gate is not None else 0, int(bool(mixed)), float(alpha), _pofs(d_x0, c0), N,
                   int(bool(accumulate_x0)), C.ptr(post_scale), C.ptr(dst), dst.shape[1], _pofs(col_p, c0),
                   C.ptr(row_live), C.ptr(a_live), C.ptr(x0_valid), C.ptr(ws), ws_bytes,
                   ctypes.byref(desc) if desc is not None else None, C.stream_ptr(A.device))
        if push is not None:
            push.pushed(p)
        if col_p is not col_sum:
            col_parts.append(col_p)
    if col_parts:
        col_sum.copy_(torch.stack(col_parts).sum(0))
    return out, col_sum, d_x0


def gemm_tn_supported(M, Ka, Nb, dtype=torch.float32):
    if dtype not in _STORAGE:
        return False
    return bool(getattr(C.lib(), 'cb_gemm_tn_supported' + _sfx(dtype))(max(int(M), 1), int(Ka), int(Nb)))


def gemm_tn_raw(A, B, a_row_scale=None, b_row_scale=None):
    """A^T @ B for row-major A [M, Ka], B [M, Nb] (the weight gradient), fp32 result: fp32 operands -> 3xTF32, bf16
    operands -> kind::f16 as stored.  A per-row scale s[m] of either operand (sum_m s[m] A[m,:]^T B[m,:]) is applied
    while the fp32 operand is split in shared memory (the bf16 variant takes pre-scaled operands)."""
    _need_cuda(A, B, a_row_scale, b_row_scale)
    st = A.dtype
    if st not in _STORAGE or B.dtype != st:
        raise TypeError(f'operands must both be float32 or bfloat16, got {A.dtype} / {B.dtype}')
    A, B = A.contiguous(), B.contiguous()
    if a_row_scale is not None and b_row_scale is not None:
        raise ValueError('one row scale at most')
    scale, scale_b = (_f32c(b_row_scale), 1) if b_row_scale is not None else (_f32c(a_row_scale), 0)
    M, Ka = A.shape
    if B.shape[0] != M:
        raise ValueError(f'A has {M} rows, B has {B.shape[0]}')
    Nb = B.shape[1]
    if M == 0:
        return torch.zeros((Ka, Nb), dtype=torch.float32, device=A.device)
    out = torch.empty((Ka, Nb), dtype=torch.float32, device=A.device)
    nb = int(getattr(C.lib(), 'cb_gemm_tn_workspace_bytes' + _sfx(st))(M, Ka, Nb))
    ws = torch.empty(nb, dtype=torch.uint8, device=A.device)
    alg = A.element_size() * (M * Ka + M * Nb) + 4 * Ka * Nb
    mma = 6 if st == torch.float32 else 2
    with torch.cuda.device(A.device), _Timed('gemm_tn' + _sfx(st), alg, A.device, flops=mma * M * Ka * Nb):
        if st == torch.float32:
            C.call('cb_gemm_tn', C.ptr(A), Ka, C.ptr(B), Nb, M, Ka, Nb, C.ptr(scale), scale_b, C.ptr(out), Nb,
                   C.ptr(ws), nb, C.stream_ptr(A.device))
        else:
            if scale is not None:
                raise ValueError('the bf16 weight-gradient kernel takes pre-scaled operands')
            C.call('cb_gemm_tn_bf16', C.ptr(A), Ka, C.ptr(B), Nb, M, Ka, Nb, C.ptr(out), Nb, C.ptr(ws), nb,
                   C.stream_ptr(A.device))
    return out


def to_bf16_raw(x):
    """bf16(x) for an fp32 tensor (cb_to_bf16, round to nearest even)."""
    _need_cuda(x)
    x = _f32c(x)
    y = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        C.call('cb_to_bf16', C.ptr(x), x.numel(), C.ptr(y), C.stream_ptr(x.device))
    return y


# ---------------------------------------------------------------------------------------------
# differentiable ops
# ---------------------------------------------------------------------------------------------

class _RowScale(torch.autograd.Function):
    """y = s[:,None] * x with s constant (GCN.py:205-213); the adjoint is the same kernel."""

    @staticmethod
    def forward(ctx, x, s):
        ctx.save_for_backward(s)
        return row_scale_raw(x, s)

    @staticmethod
    def backward(ctx, dy):
        (s,) = ctx.saved_tensors
        return row_scale_raw(dy, s), None


def row_scale(x, s):
    return _RowScale.apply(x, s)


class _FrobNorm(torch.autograd.Function):
    """||E||_F (GCN.py:232 th.norm(self.le)); gradient g * E / ||E||, zero at E = 0 like torch.norm."""

    @staticmethod
    def forward(ctx, e, graph):
        ss = sumsq_raw(e)
        if graph is not None:
            ss = graph.allreduce_sum(ss)        # row-sharded table: the norm is over every rank's rows
        n = ss.sqrt_().reshape(())
        ctx.save_for_backward(e, n)
        return n

    @staticmethod
    def backward(ctx, g):
        e, n = ctx.saved_tensors
        scale = torch.where(n > 0, g / n, torch.zeros_like(n))
        return e * scale, None


def frob_norm(e, graph=None):
    return _FrobNorm.apply(e, graph)


# ---------------------------------------------------------------------------------------------
# gradient sink: the Initial residual x0 feeds every layer (res_tricks.py:23), so autograd would add
# one [N, d] gradient per consumer with separate full passes.  The fused backward kernels instead
# accumulate their contributions in place into one buffer, which the last consumer folds into its own
# GEMM epilogue (or, failing that, the hub adds once).
# ---------------------------------------------------------------------------------------------
class GradSink:
    """``buf``: the accumulated gradient.  ``valid``: None, or uint8 [rows] flags left by a row-sparse first writer
    (cb_gemm_rows_grad with a_live): rows with 0 were never written and count as zero.  The next cb_gemm_rows_grad
    reads the buffer through those flags; every other consumer gets it materialised."""
    __slots__ = ('buf', 'valid')

    def __init__(self):
        self.buf = self.valid = None

    def dense(self):
        """The buffer with the never-written rows zeroed (no-op when every row is valid)."""
        if self.buf is not None and self.valid is not None:
            self.buf = torch.where(self.valid.bool()[:, None], self.buf, torch.zeros((), dtype=self.buf.dtype,
                                                                                     device=self.buf.device))
            self.valid = None
        return self.buf

    def take(self):
        b = self.dense()
        self.buf = self.valid = None
        return b


_backward_fusion = True


def set_backward_fusion(on):
    """False: every backward prologue runs as its own kernel(s) (cb_agg_backward_prep, torch relu/bias
    backward) instead of inside the neighbouring dX GEMM -- the A/B switch of tests and bench."""
    global _backward_fusion
    _backward_fusion = bool(on)


_dropout_fusion = True


def set_dropout_fusion(on):
    """False: the dropouts between the layers stay separate F.dropout autograd nodes (the A/B partner of the layers
    that own their output's dropout, see fused_aggregate(dropout_p=...)); same masks, same gradients."""
    global _dropout_fusion
    _dropout_fusion = bool(on)


def dropout_fusion():
    return _dropout_fusion


def new_plan():
    return BwdPlan() if (_backward_fusion and torch.is_grad_enabled()) else None


class BwdPlan:
    """Backward hand-off between two neighbouring autograd nodes.

    The op that PRODUCED an activation y (a Linear+relu, or a fused aggregation) starts its backward with
    an elementwise prologue on dL/dy (relu mask, residual split, degree scale, bias column sums).  When y
    has exactly one consumer and that consumer's backward produces dL/dy with a GEMM, the prologue runs in
    that GEMM's epilogue instead (cb_gemm_rows_grad): the producer fills the plan in its forward, the
    consumer's backward executes it and leaves ``result``; the producer's backward then finds its prologue
    already done.  The caller (TricksComb) creates a plan only where the single-consumer condition holds by
    construction."""
    __slots__ = ('kind', 'graph', 'gate_u8', 'gate_f32', 'relu', 'mixed', 'alpha', 'want_bias', 'want_x0',
                 'x0_sink', 'slot', 'result', 'row_sparse_hint', 'dy_live')

    def __init__(self):
        self.kind = None       # 'relu_bias' (Linear + relu) | 'prep' (fused aggregation)
        self.result = None     # set by the consumer's backward: {'d_bias': tensor or None}
        self.graph = self.gate_u8 = self.gate_f32 = self.x0_sink = None
        self.relu = self.mixed = self.want_bias = self.want_x0 = False
        self.alpha = 0.0
        self.slot = None       # 'out' | 'out_scaled': which output of the aggregation the consumer reads
        # set by the caller for the layer under the output head: its gradient is row-sparse when the loss reads
        # the train rows only, so the GEMM also records which rows of G are non-zero and the gather skips the rest
        self.row_sparse_hint = False
        self.dy_live = None    # left by run(): uint8 [M] flags of the non-zero rows of the consumer's incoming gradient

    def run(self, dtot_in, wb, row_scale, add):
        """Called from the consumer's backward: dX GEMM + this plan's prologue.  ``row_scale``/``add`` are
        the consumer's own epilogue terms (its out-degree scale, parked residual gradients)."""
        if self.kind == 'relu_bias':
            out, col, _ = gemm_rows_grad_raw(dtot_in, wb, row_scale=row_scale, add=add, gate_u8=self.gate_u8,
                                             gate_f32=self.gate_f32, want_col_sum=self.want_bias)
            self.result = {'d_bias': col}
            return out
        g = self.graph
        rs = row_scale
        if self.slot == 'out_scaled':      # dtot = dout^-1/2 * d(out_scaled)
            if rs is not None:
                raise RuntimeError('BwdPlan: a pre-scaled input cannot carry a second row scale')
            rs = g.dout_inv_sqrt
        sink = self.x0_sink if self.want_x0 else None
        if add is not None and sink is not None and sink.buf is not None:
            return None     # the kernel keeps one [M, N] epilogue input: the caller runs the two-kernel path
        # G is what the transposed aggregation gathers; a row-sparse G is gathered over the compacted lists in one go
        # (no source-panel passes)
        slot = g.push_slot(C.CB_BY_SRC, wb.n, dtot_in.dtype, passes=not self.row_sparse_hint)
        live = push_live = live_full = a_live = kernel_live = None
        if self.row_sparse_hint and add is None:
            # A zero row of the incoming gradient gives a zero row of dtot, of G and of the d_x0 contribution -- known
            # BEFORE the GEMM (one pass over the [M, K] gradient, cb_row_any_nonzero).  Such rows are not loaded, not
            # stored, not pushed to the peers and not gathered by anyone: the GEMM, the pushes and the compacted
            # gather all go by the same flags, and the x0 sink remembers which of its rows were never written.
            # Measured at the bench shape (scripts/grad_sparse_bench.py): 7.48 -> 5.75 + 0.90 ms for this GEMM and
            # 12.2 -> 10.2 ms for the next one, which no longer reads the 90 % of d_x0 that would have been zeros.
            a_live = row_any_nonzero_raw(dtot_in)
            self.dy_live = a_live          # the consumer's weight gradient skips the same rows (_Dense.backward)
            if slot is not None:
                # multi-GPU: the flags of every rank are all-gathered HERE, ahead of the first panel's GEMM: the
                # side-stream gathers wait only for their panel's event, which is recorded after this point, so
                # they can never read a half-written flag array.
                push_live = a_live
                live_full = compact_live_raw(g, C.CB_BY_SRC, g.exchange_flags(a_live))   # the compacted workspace
            else:
                live = a_live
        elif self.row_sparse_hint:
            # `add` makes rows non-zero that the incoming gradient does not: the kernel reports what it stored
            live = kernel_live = torch.zeros(dtot_in.shape[0], dtype=torch.uint8, device=dtot_in.device)
        accumulate = sink is not None and sink.buf is not None
        out, col, d_x0 = gemm_rows_grad_raw(
            dtot_in, wb, row_scale=rs, add=add, gate_u8=self.gate_u8, gate_f32=self.gate_f32 if self.relu else None,
            mixed=self.mixed, alpha=self.alpha, d_x0=sink.buf if sink is not None else None,
            accumulate_x0=accumulate, want_x0=self.want_x0,
            post_scale=g.din_inv_sqrt, want_col_sum=self.want_bias, push=slot, row_live=kernel_live,
            push_live=push_live, a_live=a_live, x0_valid=sink.valid if accumulate else None)
        if sink is not None:
            if a_live is None:
                sink.valid = None                    # a dense writer: every row holds its sum now
            elif not accumulate:
                sink.valid = a_live                  # first writer skipped the dead rows: they count as zero
            elif sink.valid is not None:
                sink.valid = sink.valid | a_live     # rows written now or before
            sink.buf, d_x0 = d_x0, None
        self.result = {'d_bias': col, 'd_x0': d_x0, 'G': out, 'live': live, 'live_full': live_full}
        if out.dim() == 3:
            # a panelled slot cannot be viewed as [M, N]: G travels in the plan, autograd gets a placeholder
            return out.new_zeros(()).expand(dtot_in.shape[0], wb.n)
        return out

    def take_result(self):
        r, self.result = self.result, None
        return r


class _SinkHub(torch.autograd.Function):
    """Identity on the shared tensor; its backward adds whatever is still parked in the sink."""

    @staticmethod
    def forward(ctx, x, sink):
        ctx.sink = sink
        ctx.set_materialize_grads(False)
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        parked = ctx.sink.take()
        if parked is not None:
            g = parked if g is None else g + parked
        return g, None


def sink_hub(x, sink):
    return _SinkHub.apply(x, sink)


# ---------------------------------------------------------------------------------------------
# dense transform: act(row_scale * (x @ W) + bias + add) with the tcgen05 kernels, cuBLAS for the
# shapes they do not cover (N or K not a multiple of 4; weight gradient: not a multiple of 32)
# ---------------------------------------------------------------------------------------------
_dense_backend = 'tcgen05'


def set_dense_backend(name):
    """'tcgen05' (default): 3xTF32 tensor-core kernels of this library wherever the shape allows;
    'cublas': torch.matmul in fp32 everywhere (the reference's own GEMM path, for A/B comparisons)."""
    global _dense_backend
    if name not in ('tcgen05', 'cublas'):
        raise ValueError(name)
    _dense_backend = name


def _w_as_kn(weight, layout):
    return weight if layout == 'kn' else weight.t()


def _dense_composite(x, weight, layout, bias, add, relu, row_scale, out2_scale, want_out, want_out2):
    """Library-GEMM path for the shapes the tcgen05 kernels do not cover; bf16 inputs: bf16 GEMM, fp32 epilogue."""
    st = x.dtype
    acc = x @ _w_as_kn(weight, layout).to(st)
    if st != torch.float32:
        acc = acc.float()
    if row_scale is not None:
        acc = acc * row_scale[:, None]
    if bias is not None:
        acc = acc + bias
    if add is not None:
        acc = acc + add
    if relu:
        acc = torch.relu(acc)
    out2 = (acc * out2_scale[:, None]).to(st) if want_out2 else None
    return (acc.to(st) if want_out else None), out2


class _Dense(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, add, row_scale, out2_scale, layout, relu, want_out, want_out2, dx_sink,
                my_plan, dx_plan, push_graph, add_sink):
        ctx.dx_sink, ctx.add_sink = dx_sink, add_sink
        ctx.my_plan, ctx.dx_plan = my_plan, dx_plan
        wt = split_weight(weight, transpose=(layout == 'kn'), dtype=x.dtype)
        # multi-GPU: the output is what the next aggregation gathers -> write it into the exchange buffer and
        # into the peers from the epilogue (graph.exchange() then only has to wait for everyone's pushes)
        slot = push_graph.push_slot(C.CB_BY_DST, wt.n, x.dtype) if (push_graph is not None and want_out) else None
        need = any(ctx.needs_input_grad[:4]) or add_sink is not None
        # the relu/bias backward of this op will ride on its consumer's dX GEMM (see below): that GEMM then gates with
        # a byte mask written here instead of re-reading the activations
        hands_off_relu = (my_plan is not None and relu and want_out and not want_out2 and add is None and need and
                          slot is None and x.shape[0] > 0)
        rmask = None
        if hands_off_relu:
            out, rmask = gemm_rows_raw(x, wt, row_scale, bias, add, relu, out2_scale, want_out, want_out2,
                                       want_relu_mask=True)
            out2 = None
        else:
            res = gemm_rows_raw(x, wt, row_scale, bias, add, relu, out2_scale, want_out, want_out2, push=slot)
            out, out2 = res if want_out2 else (res, None)
        ctx.layout, ctx.relu = layout, relu
        ctx.has_bias, ctx.has_add = bias is not None, add is not None
        keep_y = (out if out is not None else out2) if (relu and need) else None
        ctx.save_for_backward(x if need else None, weight if need else None, row_scale, out2_scale, keep_y)
        ctx.set_materialize_grads(False)
        if my_plan is not None:
            my_plan.kind = None
            if relu and want_out and not want_out2 and add is None and need:
                # detached alias: a plan must not keep the autograd graph alive (the graph owns the plan)
                my_plan.kind = 'relu_bias'
                my_plan.gate_u8, my_plan.gate_f32 = (rmask, None) if rmask is not None else (None, out.detach())
                my_plan.want_bias = bias is not None and ctx.needs_input_grad[2]
        empty = x.new_empty(0)
        return (out if out is not None else empty), (out2 if out2 is not None else empty)

    @staticmethod
    def backward(ctx, dy, dy2):
        x, weight, row_scale, out2_scale, y = ctx.saved_tensors
        if dy is not None and dy.dim() == 3:      # the output was a panelled exchange slot [M, panels, N/panels]
            dy = dy.reshape(dy.shape[0], -1)
        # the unused output slot was a 1-D empty placeholder; a real [0, N] gradient (a rank that owns no rows)
        # must still run, so that this rank enters the exchange barriers of the backward pass
        if dy is not None and dy.dim() == 1:
            dy = None
        if dy2 is not None and dy2.dim() == 1:
            dy2 = None
        if dy is None and dy2 is None:
            return (None,) * 15
        st = x.dtype
        done = ctx.my_plan.take_result() if ctx.my_plan is not None else None
        if done is not None:
            # the consumer's dX GEMM already applied the relu mask and summed the bias gradient
            dtot, d_bias = dy.contiguous(), done['d_bias']
        else:
            if dy2 is not None:
                dy2 = (dy2 * out2_scale[:, None]).to(st)
            dtot = dy2 if dy is None else (dy if dy2 is None else dy + dy2)
            if ctx.relu:
                dtot = torch.where(y > 0, dtot, torch.zeros((), dtype=dtot.dtype, device=dtot.device))
            dtot = dtot.contiguous()
            d_bias = dtot.sum(0, dtype=torch.float32) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        d_add = dtot if (ctx.has_add and ctx.needs_input_grad[3]) else None
        if ctx.add_sink is not None:
            # h = (D X) W + E: dL/dE is dL/dh itself.  Handed to the table's optimizer by reference (no autograd
            # accumulation pass, no fp32 copy of a bf16 gradient): se_optim.FusedSEAdam reads it.
            ctx.add_sink.grad = dtot
        M, K = x.shape
        N = dtot.shape[1]
        dx = dw = None
        if ctx.needs_input_grad[0]:
            parked = ctx.dx_sink.take() if ctx.dx_sink is not None else None   # other consumers' share of dx
            plan = ctx.dx_plan
            if gemm_supported(M, K, N, st):
                wb = split_weight(weight, transpose=(ctx.layout == 'nk'), dtype=st)
                dx = plan.run(dtot, wb, row_scale, parked) if (plan is not None and plan.kind is not None) else None
                if dx is None:
                    dx = gemm_rows_raw(dtot, wb, row_scale=row_scale, add=parked)
            else:
                dx = dtot @ _w_as_kn(weight, ctx.layout).t().to(st)
                if row_scale is not None:
                    dx = (dx * row_scale[:, None]).to(st)
                if parked is not None:
                    dx = dx + parked
        dy_live = None
        if ctx.dx_plan is not None:
            dy_live, ctx.dx_plan.dy_live = ctx.dx_plan.dy_live, None
        rows_kept = None
        if (ctx.needs_input_grad[1] and dy_live is not None and M >= _COMPACT_DW_MIN_ROWS and
                not torch.cuda.is_current_stream_capturing()):
            # Row-sparse incoming gradient (the head under a loss over the train rows): dW = X^T dY only has terms from
            # the live rows, wherever they sit.  Their indices are read back once (the one host sync of the step) and
            # the weight-gradient GEMM runs on the compacted operands -- unless most rows are live anyway.
            rows_kept = torch.nonzero(dy_live).squeeze(1)
            if rows_kept.numel() * 2 > M:
                rows_kept = None
        if rows_kept is not None:
            x = x.index_select(0, rows_kept)
            dtot_w = dtot.index_select(0, rows_kept)
            row_scale = row_scale.index_select(0, rows_kept) if row_scale is not None else None
            M = int(rows_kept.numel())
        else:
            dtot_w = dtot
        if ctx.needs_input_grad[1]:
            dtot = dtot_w       # (dx above is done with the full gradient)
            if gemm_tn_supported(M, K, N, st):
                if st != torch.float32 and row_scale is not None:
                    # the bf16 kernel takes operands as stored: one extra [M, K] pass for the layer that reads an
                    # unscaled input (layer 0 of the residual topologies); fp32 scale, one rounding
                    xs, rs = (x.float() * row_scale[:, None]).to(st), None
                else:
                    xs, rs = x, row_scale
                dw = gemm_tn_raw(xs, dtot, a_row_scale=rs) if ctx.layout == 'kn' else gemm_tn_raw(dtot, xs, b_row_scale=rs)
            else:
                group = 32 if st == torch.float32 else 64
                Kp, Np = -(-K // group) * group, -(-N // group) * group
                xs = x if row_scale is None else (x.float() * row_scale[:, None]).to(st)
                if gemm_tn_supported(M, Kp, Np, st):
                    # whole 128-byte feature groups for the tensor-core kernel: zero columns, sliced off the result
                    xs = torch.nn.functional.pad(xs, (0, Kp - K)) if Kp != K else xs
                    dp = torch.nn.functional.pad(dtot, (0, Np - N)) if Np != N else dtot
                    dw = gemm_tn_raw(xs, dp)[:K, :N] if ctx.layout == 'kn' else gemm_tn_raw(dp, xs)[:N, :K]
                else:
                    dw = (xs.t() @ dtot if ctx.layout == 'kn' else dtot.t() @ xs).float()
        return dx, dw, d_bias, d_add, None, None, None, None, None, None, None, None, None, None, None


# a row-sparse weight gradient is compacted (one host sync) only where the GEMM it saves is worth it
_COMPACT_DW_MIN_ROWS = 1 << 18


class GradSlot:
    """Receives the gradient of a non-autograd ``add`` operand of dense() (the SE table in fused-optimizer mode)."""
    __slots__ = ('grad',)

    def __init__(self):
        self.grad = None


def dense(x, weight, layout, bias=None, add=None, relu=False, row_scale=None, out2_scale=None, want_out=True,
          want_out2=False, dx_sink=None, my_plan=None, dx_plan=None, push_graph=None, add_sink=None):
    """act(row_scale[:,None] * (x @ W) + bias + add); also out2_scale[:,None] * that when want_out2.
    layout 'kn': weight is [in, out] (GCNConv.weight, GCN.py:170); 'nk': [out, in] (nn.Linear.weight).
    my_plan / dx_plan: BwdPlan of this op's own backward prologue / of the op that produced ``x``.
    add_sink: GradSlot that receives dL/d(add) when ``add`` is not an autograd leaf (fused SE optimizer).
    push_graph: node-sliced graph whose next aggregation gathers ``out`` (the exchange rides on the epilogue).
    Returns (out, out2); the one not asked for is None."""
    M, K = x.shape
    N = weight.shape[1] if layout == 'kn' else weight.shape[0]
    tc = _dense_backend == 'tcgen05' and x.is_cuda and x.dtype in _STORAGE
    group = 32 if x.dtype == torch.float32 else 64      # feature group (128 bytes) of the weight-gradient kernel
    if tc and gemm_supported(M, N, K, x.dtype):
        out, out2 = _Dense.apply(x, weight, bias, add, row_scale, out2_scale, layout, bool(relu), bool(want_out),
                                 bool(want_out2), dx_sink, my_plan, dx_plan, push_graph, add_sink)
        return (out if want_out else None), (out2 if want_out2 else None)
    Kp, Np = -(-K // group) * group, -(-N // group) * group
    if tc and add_sink is None and gemm_supported(M, Np, Kp, x.dtype):
        # Ragged widths (Cora F = 1433 / C = 7, Pubmed C = 3): TMA needs 16-byte row pitches and the epilogue 4-element
        # accesses, so the operands are zero-padded to the next 128-byte feature group (autograd slices the gradients
        # back) and the SAME tensor-core kernels run -- no library GEMM.  The padded layer keeps its own backward
        # prologue (no hand-off plans, no epilogue push).  Widths the rows kernels take but the weight-gradient kernel
        # does not (Pubmed F = 500, ogbn-arxiv C = 40) are padded inside _Dense.backward only.
        F = torch.nn.functional
        xp = F.pad(x, (0, Kp - K)) if Kp != K else x
        wp = F.pad(weight, (0, Np - N, 0, Kp - K)) if layout == 'kn' else F.pad(weight, (0, Kp - K, 0, Np - N))
        bp = F.pad(bias, (0, Np - N)) if (bias is not None and Np != N) else bias
        ap = F.pad(add, (0, Np - N)) if (add is not None and Np != N) else add
        if my_plan is not None:
            my_plan.kind = None
        out, out2 = _Dense.apply(xp, wp, bp, ap, row_scale, out2_scale, layout, bool(relu), bool(want_out),
                                 bool(want_out2), None, None, None, None, None)
        return (out[:, :N] if want_out else None), (out2[:, :N] if want_out2 else None)
    _need_cuda(x)
    if add_sink is not None:
        raise RuntimeError('the fused SE optimizer needs the tcgen05 transform (shape not covered by cb_gemm_rows)')
    if my_plan is not None:   # library GEMM path: no fused backward prologue
        my_plan.kind = None
    return _dense_composite(x, weight, layout, bias, add, relu, row_scale, out2_scale, want_out, want_out2)


def _panelled(graph, ex, make_outputs, run_panel, after=None):
    """Runs ``run_panel(p, rows_of_panel_p, outputs)`` for every panel of an exchange slot on its side stream,
    each as soon as that panel's barrier has passed, while the compute stream goes on pushing the next panels.

    The outputs are allocated ON the side stream: a block the compute-stream allocator hands out may still
    be read by kernels queued there (e.g. the weight-gradient GEMM of the layer above reading the previous
    dH), and the side stream deliberately does not wait for those."""
    cur = torch.cuda.current_stream(graph.device)
    side = ex.side_stream
    if os.environ.get('CB_PANEL_SERIAL'):       # debugging aid: no overlap, everything on the compute stream
        outputs = make_outputs()
        for p in range(ex.n_launches):
            run_panel(p, ex.rows(p if ex.src_passes == 1 else 0, graph.num_nodes), outputs)
        return outputs
    with torch.cuda.stream(side):
        if after is not None:
            side.wait_event(after)
        outputs = make_outputs()
        for p in range(ex.n_launches):
            side.wait_event(ex.events[p])
            run_panel(p, ex.rows(p if ex.src_passes == 1 else 0, graph.num_nodes), outputs)
    cur.wait_stream(side)
    for t in outputs:
        if t is not None:
            t.record_stream(cur)
    return outputs


def aggregate_forward(graph, H, bias, x0, alpha, relu, want_out, want_scaled, want_mask):
    """Exchange + fused forward aggregation; panel-pipelined when H was pushed in column panels."""
    ex = graph.exchange(H)
    if torch.is_tensor(ex):
        return agg_forward_raw(graph, ex, bias, x0, alpha, relu, want_out, want_scaled, want_mask)
    shape, dev, pw = (graph.rows, ex.width), graph.device, ex.panel_width

    def make():
        return (torch.empty(shape, dtype=H.dtype, device=dev) if want_out else None,
                torch.empty(shape, dtype=H.dtype, device=dev) if want_scaled else None,
                torch.empty(shape, dtype=torch.uint8, device=dev) if want_mask else None)

    if ex.src_passes > 1:     # source-panel passes at full row width, the sums carried from pass to pass
        return _panelled(graph, ex, make, lambda p, Hf, outs: agg_forward_raw(
            graph, Hf, bias, x0, alpha, relu, outs=outs, src_pass=p))
    return _panelled(graph, ex, make, lambda p, Hp, outs: agg_forward_raw(
        graph, Hp, bias, x0, alpha, relu, outs=outs, panel=(p * pw, pw)))


def aggregate_gather(graph, side, X, row_scale=None, live=None, live_full=None):
    """Exchange + plain gather-reduce over one side; panel-pipelined when X was pushed in column panels.
    live: uint8 [rows] flags of the local rows of X (0 = all-zero row), exchanged like the rows themselves.
    live_full: the compacted workspace built from the already exchanged flags (ordered before the pushes -- see
    BwdPlan.run)."""
    flags_event = None
    live_ws = live_full
    if live_ws is None and live is not None:
        live_ws = compact_live_raw(graph, side, graph.exchange_flags(live))
        if graph.world > 1:
            # built on the compute stream AFTER the pushes: a side-stream gather must wait for it explicitly
            flags_event = torch.cuda.Event()
            flags_event.record(torch.cuda.current_stream(graph.device))
    ex = graph.exchange(X)
    if torch.is_tensor(ex):
        return agg_gather_raw(graph, side, ex, row_scale, live_ws=live_ws)
    pw = ex.panel_width
    make = lambda: (torch.empty((graph.rows, ex.width), dtype=X.dtype, device=graph.device),)   # noqa: E731
    if ex.src_passes > 1:
        if live_ws is not None:
            raise RuntimeError('a row-sparse gather takes a one-launch exchange slot (push_slot(passes=False))')
        (out,) = _panelled(graph, ex, make, lambda p, Xf, outs: agg_gather_raw(graph, side, Xf, row_scale, out=outs[0],
                                                                                src_pass=p))
        return out
    (out,) = _panelled(graph, ex, make,
                       lambda p, Xp, outs: agg_gather_raw(graph, side, Xp, row_scale, out=outs[0], panel=(p * pw, pw),
                                                          live_ws=live_ws), after=flags_event)
    return out


class _FusedAggregate(torch.autograd.Function):
    """out = mix(act(din^-1/2 * A^T-sum(H) + b), x0), optionally also dout^-1/2 * out.

    forward : halo exchange of H (identity on one GPU) -> cb_agg_forward
    backward: cb_agg_backward_prep -> halo exchange of G -> cb_agg_gather over the by-source CSR
    """

    @staticmethod
    def forward(ctx, H, bias, x0, graph, alpha, relu, want_out, want_scaled, x0_sink, my_plan, row_sparse_hint,
                dropout_p):
        ctx.x0_sink = x0_sink
        ctx.my_plan = my_plan
        ctx.row_sparse_hint = row_sparse_hint
        mixed = x0 is not None
        need_grad = any(ctx.needs_input_grad[:3])
        owns_dropout = dropout_p > 0.0
        if owns_dropout and (want_scaled or not want_out):
            raise ValueError('fused_aggregate: the dropout applies to the plain output only')
        # relu mask source for backward: the plain relu output doubles as the mask when nothing was
        # mixed into it, otherwise a byte mask is written by the kernel
        use_out_as_mask = relu and not mixed and want_out
        want_mask = relu and need_grad and not use_out_as_mask
        out, out_scaled, mask = aggregate_forward(graph, H, bias, x0, alpha, relu, want_out, want_scaled, want_mask)
        ctx.h_shape = H.shape
        ctx.graph, ctx.alpha, ctx.relu, ctx.mixed = graph, alpha, relu, mixed
        ctx.has_bias, ctx.use_out_as_mask = bias is not None, use_out_as_mask
        ctx.drop_scale, keep = 1.0, None
        pre_drop = out
        if owns_dropout:
            # The dropout that follows the layer (GCN.py:104,110,133) is drawn HERE with the same call F.dropout makes,
            # so the Philox stream and the mask are the reference's; the layer keeps the mask and folds the dropout's
            # backward into its own prologue (cb_agg_backward_prep_ex) instead of a separate pass over [N, d].
            out, keep = torch.native_dropout(out, dropout_p, True)
            ctx.drop_scale = 1.0 / (1.0 - dropout_p)
        ctx.save_for_backward(mask if want_mask else None, pre_drop if (use_out_as_mask and need_grad) else None, keep)
        ctx.set_materialize_grads(False)
        if my_plan is not None:
            my_plan.kind = None
            if need_grad and (want_out != want_scaled) and not owns_dropout:
                p = my_plan
                p.kind, p.graph, p.slot = 'prep', graph, 'out' if want_out else 'out_scaled'
                p.gate_u8 = mask if want_mask else None
                p.gate_f32 = out.detach() if use_out_as_mask else None
                p.relu, p.mixed, p.alpha = relu, mixed, alpha
                p.want_bias = bias is not None and ctx.needs_input_grad[1]
                p.want_x0 = mixed and ctx.needs_input_grad[2]
                p.x0_sink = x0_sink
        if out is not None and out_scaled is not None:
            return out, out_scaled
        # a Function must return tensors; the unused slot gets an empty placeholder
        empty = H.new_empty(0)
        return (out if out is not None else empty), (out_scaled if out_scaled is not None else empty)

    @staticmethod
    def backward(ctx, d_out, d_out_scaled):
        mask, relu_out, drop_keep = ctx.saved_tensors
        graph = ctx.graph
        if d_out is not None and d_out.dim() == 1:          # 1-D empty placeholder of the unused slot
            d_out = None
        if d_out_scaled is not None and d_out_scaled.dim() == 1:
            d_out_scaled = None
        if d_out is None and d_out_scaled is None:
            return (None,) * 12
        done = ctx.my_plan.take_result() if ctx.my_plan is not None else None
        if done is not None:
            # the consumer's dX GEMM ran the prologue in its epilogue: what arrived is G itself
            G = done.get('G')
            if G is None:
                G = (d_out if d_out is not None else d_out_scaled).contiguous()
            d_bias, d_x0, live, live_full = done['d_bias'], done['d_x0'], done.get('live'), done.get('live_full')
        else:
            live = live_full = None
            want_bias = ctx.has_bias and ctx.needs_input_grad[1]
            want_x0 = ctx.mixed and ctx.needs_input_grad[2]
            sink = ctx.x0_sink if want_x0 else None
            # The layer under the output head, reached without the fused hand-off (dropout in between, training with the
            # reference's defaults): under a loss over the train rows only most rows of G are zero.  The prologue flags
            # the others while it writes G; the gather then walks the compacted lists (same sums).
            want_live = ctx.row_sparse_hint and ctx.needs_input_grad[0] and graph.rows > 0
            res = backward_prep_raw(graph, d_out, d_out_scaled, mask, relu_out, ctx.relu, ctx.mixed,
                                    ctx.alpha, want_bias, want_x0,
                                    d_x0_accum=sink.dense() if sink is not None else None,
                                    drop_keep=drop_keep, drop_scale=ctx.drop_scale, want_live=want_live)
            G, d_bias, d_x0 = res[:3]
            live = res[3] if want_live else None
            if sink is not None:   # parked for the hub / the last consumer's GEMM epilogue
                sink.buf, d_x0 = d_x0, None
        dH = None
        if ctx.needs_input_grad[0]:
            dH = aggregate_gather(graph, C.CB_BY_SRC, G, live=live, live_full=live_full).view(ctx.h_shape)
        return dH, d_bias, d_x0, None, None, None, None, None, None, None, None, None


def fused_aggregate(H, graph, bias=None, x0=None, alpha=0.0, relu=False, want_out=True, want_scaled=False,
                    x0_sink=None, my_plan=None, row_sparse_hint=False, dropout_p=0.0):
    """Returns (out, out_scaled); the one not asked for is None.
    row_sparse_hint: the gradient of this output is expected to be zero on most rows (the layer under an output head
    whose loss reads the train rows only): the backward pass then looks for all-zero rows and gathers over the
    compacted lists.  A hint only -- the rows are found from the data, the sums are the same.
    dropout_p > 0: the returned ``out`` is F.dropout(out, dropout_p, training=True) -- the same torch call, hence the
    same mask -- and the dropout's backward runs inside this op's own backward prologue."""
    if not (want_out or want_scaled):
        raise ValueError('fused_aggregate: nothing requested')
    out, out_scaled = _FusedAggregate.apply(H, bias, x0, graph, float(alpha), bool(relu), bool(want_out),
                                            bool(want_scaled), x0_sink, my_plan, bool(row_sparse_hint), float(dropout_p))
    return (out if want_out else None), (out_scaled if want_scaled else None)


class _CopySum(torch.autograd.Function):
    """rst[v] = sum_{(u->v)} h[u]: the bare update_all(copy_src, sum) of GCN.py:238."""

    @staticmethod
    def forward(ctx, H, graph):
        ctx.graph, ctx.h_shape = graph, H.shape
        return aggregate_gather(graph, C.CB_BY_DST, H)

    @staticmethod
    def backward(ctx, d):
        return aggregate_gather(ctx.graph, C.CB_BY_SRC, d.contiguous()).view(ctx.h_shape), None


def copy_sum(H, graph):
    return _CopySum.apply(H, graph)


# ---------------------------------------------------------------------------------------------
# edge-weighted aggregation: graph.update_all(fn.u_mul_e('h', '_edge_weight', 'm'), fn.sum('m', 'h')), GCN.py:199-202
# ---------------------------------------------------------------------------------------------
def sort_edge_values_raw(graph, side, values):
    """Per-edge fp32 values in the order of the caller's edge list -> the stored order of one CSR side."""
    _need_cuda(values)
    v = values.reshape(-1).to(torch.float32).contiguous()
    e = graph.num_edges if side == C.CB_BY_DST else graph.num_edges_by_src
    out = torch.empty(e, dtype=torch.float32, device=v.device)
    with torch.cuda.device(v.device):
        C.call('cb_graph_sort_edge_values', graph.handle, side, C.ptr(v), v.numel(), C.ptr(out), C.stream_ptr(v.device))
    return out


def agg_gather_weighted_raw(graph, side, X, edge_val, row_scale=None):
    """out[r] = row_scale[r] * sum_j X[col[j]] * edge_val[j] (edge_val in stored order, cb_agg_gather_weighted)."""
    _need_cuda(X, edge_val, row_scale)
    if X.dtype not in _STORAGE:
        raise TypeError(f'X must be float32 or bfloat16, got {X.dtype}')
    X, row_scale = X.contiguous(), _f32c(row_scale)
    d = X.shape[1]
    if X.shape[0] != graph.num_nodes:
        raise ValueError(f'X has {X.shape[0]} rows, the graph has {graph.num_nodes} nodes')
    out = torch.empty((graph.rows, d), dtype=X.dtype, device=X.device)
    ws, ws_bytes = graph.workspace(side, d)
    alg = gather_alg_bytes(graph, side, d, 1, 1, False, int(row_scale is not None), X.element_size()) + 4 * edge_val.numel()
    with torch.cuda.device(X.device), _Timed('agg_gather_weighted', alg, X.device):
        C.call('cb_agg_gather_weighted', graph.handle, side, C.CB_F32 if X.dtype == torch.float32 else C.CB_BF16,
               C.ptr(X), d, d, C.ptr(edge_val), C.ptr(row_scale), C.ptr(out), d, C.ptr(ws), ws_bytes,
               C.stream_ptr(X.device))
    return out


def edge_dot_raw(graph, side, X, Y, num_values):
    """fp32 [num_values]: <X[col[j]], Y[row]> at the caller's position of every stored edge of ``side`` (cb_agg_edge_dot)."""
    _need_cuda(X, Y)
    if X.dtype not in _STORAGE or Y.dtype != X.dtype or X.shape[1] != Y.shape[1]:
        raise TypeError('edge_dot: X and Y must share a float32 / bfloat16 type and their width')
    X, Y = X.contiguous(), Y.contiguous()
    out = torch.zeros(int(num_values), dtype=torch.float32, device=X.device)
    with torch.cuda.device(X.device):
        C.call('cb_agg_edge_dot', graph.handle, side, C.CB_F32 if X.dtype == torch.float32 else C.CB_BF16, C.ptr(X),
               X.shape[1], C.ptr(Y), Y.shape[1], X.shape[1], C.ptr(out), C.stream_ptr(X.device))
    return out


class _WeightedSum(torch.autograd.Function):
    """rst[v] = sum_{e = (u->v)} w[e] * h[u]; backward: dh[u] = sum_{e = (u->v)} w[e] * drst[v], dw[e] = <h[u], drst[v]>."""

    @staticmethod
    def forward(ctx, H, w, graph):
        if graph.world != 1:
            raise NotImplementedError('edge_weight on a node-sliced graph: the per-edge gradient is not assembled '
                                      'across ranks (the TeacherGNN path never passes edge weights)')
        w32 = w.detach().reshape(-1).to(torch.float32).contiguous()
        ctx.graph, ctx.w_shape, ctx.w_dtype = graph, w.shape, w.dtype
        ctx.save_for_backward(H, w32)
        return agg_gather_weighted_raw(graph, C.CB_BY_DST, H, sort_edge_values_raw(graph, C.CB_BY_DST, w32))

    @staticmethod
    def backward(ctx, d):
        H, w32 = ctx.saved_tensors
        g = ctx.graph
        d = d.contiguous()
        dH = dw = None
        if ctx.needs_input_grad[0]:
            dH = agg_gather_weighted_raw(g, C.CB_BY_SRC, d, sort_edge_values_raw(g, C.CB_BY_SRC, w32))
        if ctx.needs_input_grad[1]:
            dw = edge_dot_raw(g, C.CB_BY_DST, H, d, w32.numel()).to(ctx.w_dtype).view(ctx.w_shape)
        return dH, dw, None


def weighted_sum(H, edge_weight, graph):
    """The bare ``update_all(u_mul_e, sum)`` of GCN.py:199-202,238 with autograd in H and in the edge weights."""
    if edge_weight.reshape(-1).shape[0] != graph.num_edges or (edge_weight.dim() > 1 and edge_weight.numel() != graph.num_edges):
        raise ValueError(f'edge_weight must hold one value per edge ({graph.num_edges}), got {tuple(edge_weight.shape)}')
    return _WeightedSum.apply(H, edge_weight, graph)
