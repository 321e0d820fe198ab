"""SEMLP "virtual neighbour" replacement.

The student's ``replacement`` (MLP_model/__init__.py:143-156) consumes the teacher's ``collect_SE`` output: for every
node it multiplies one row of guessed structural embeddings with the whole teacher table (a [1, N] matmul), keeps
the top-K scores, soft-maxes them and averages the K teacher rows -- a Python loop over nodes.

Here: the scores of a block of query rows against a tile of teacher rows come from the tcgen05 transform
(``cb_gemm_rows``, 3xTF32 = fp32-class scores, the teacher table as its weight operand) into a buffer sized for L2,
``cb_topk_merge`` folds the tile into running per-row top-K lists, and ``cb_topk_softmax_mix`` does the soft-max and the
weighted sum.  The [B, N] score matrix never exists; no library GEMM / top-k runs (K <= 32; larger K falls back to the
blocked torch formulation below).  Requires CUDA tensors.
"""
import ctypes

import torch
import torch.nn.functional as F

from . import _cabi as C, ops

QUERY_BLOCK = 1024      # x TILE x 4 bytes = 32 MB of scores: stays in the 126 MB L2 between the GEMM and the merge
TILE = 8192


def replacement(teacherSE, le_guess, topK_2_replace, node_idx=None, block=4096):
    """[len(node_idx), d]: for node i, softmax(top-K of le_guess[i] . teacherSE^T) . teacherSE[top-K rows]."""
    le_guess = le_guess.detach()
    teacherSE = teacherSE.detach()
    if node_idx is None:
        node_idx = torch.arange(le_guess.shape[0], device=le_guess.device)
    node_idx = torch.as_tensor(node_idx, device=le_guess.device).long()
    k = min(int(topK_2_replace), teacherSE.shape[0])
    if k > 32 or not teacherSE.is_cuda or teacherSE.dtype != torch.float32:
        if not teacherSE.is_cuda:
            raise RuntimeError('gnn_tail_generalization_b200 kernels need CUDA tensors (no CPU fallback)')
        return _replacement_blocked(teacherSE, le_guess, k, node_idx, block)
    table = teacherSE.contiguous()
    n, d = table.shape
    d4, n4 = -(-d // 4) * 4, -(-n // 4) * 4
    # the GEMM wants K % 4 == 0 and whole groups of 4 output columns: zero columns change no score, zero teacher rows
    # are never looked at (the merge is told the real tile width)
    tpad = F.pad(table, (0, d4 - d, 0, n4 - n)) if (d4 != d or n4 != n) else table
    wt = ops.split_weight(tpad, transpose=False)
    nq = node_idx.numel()
    dev = table.device
    out = torch.empty((nq, d), dtype=torch.float32, device=dev)
    top_v = torch.empty((nq, 32), dtype=torch.float32, device=dev)
    top_i = torch.empty((nq, 32), dtype=torch.int32, device=dev)
    tile = min(TILE, n4)
    scores = torch.empty((min(QUERY_BLOCK, max(nq, 1)), tile), dtype=torch.float32, device=dev)
    st = C.stream_ptr(dev)
    with torch.cuda.device(dev):
        for b0 in range(0, nq, QUERY_BLOCK):
            idx = node_idx[b0:b0 + QUERY_BLOCK]
            q = le_guess[idx].float()
            q = (F.pad(q, (0, d4 - d)) if d4 != d else q).contiguous()
            m = q.shape[0]
            for t0 in range(0, n, tile):
                width = min(tile, n - t0)
                w4 = -(-width // 4) * 4
                with ops._Timed('vn_scores_gemm', 4 * (m * d4 + 2 * w4 * d4 + m * w4), dev, flops=6 * m * w4 * d4):
                    C.call('cb_gemm_rows', C.ptr(q), m, d4, d4, ops._pofs(wt.hi, t0 * d4), ops._pofs(wt.lo, t0 * d4), w4,
                           None, None, None, 0, C.CB_ACT_NONE, C.ptr(scores), tile, None, None, 0, None, st)
                with ops._Timed('vn_topk_merge', 4 * m * width, dev):
                    C.call('cb_topk_merge', C.ptr(scores), m, width, tile, t0, k, ops._pofs(top_v, b0 * 32),
                           ops._pofs(top_i, b0 * 32), int(t0 == 0), st)
        C.call('cb_topk_softmax_mix', C.ptr(top_v), C.ptr(top_i), nq, k, C.ptr(table), d, d, C.ptr(out), st)
    return out


def _replacement_blocked(teacherSE, le_guess, k, node_idx, block):
    """K > 32: one [B, d] x [d, N] GEMM, one ``topk``, one soft-max, one gathered weighted sum per block of nodes."""
    table_t = teacherSE.t().contiguous()
    out = torch.empty((node_idx.numel(), teacherSE.shape[1]), dtype=teacherSE.dtype, device=teacherSE.device)
    for b0 in range(0, node_idx.numel(), block):
        idx = node_idx[b0:b0 + block]
        scores = le_guess[idx] @ table_t                         # [B, N]
        top, sel = torch.topk(scores, k, dim=1)                  # K largest per row
        attn = F.softmax(top, dim=1)
        out[b0:b0 + idx.numel()] = torch.einsum('bk,bkd->bd', attn, teacherSE[sel])
    return out
