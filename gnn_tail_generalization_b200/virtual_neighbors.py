"""SEMLP "virtual neighbour" replacement, batched.

The student's ``replacement`` (MLP_model/__init__.py:143-156) consumes the teacher's ``collect_SE`` output: for every
node it multiplies one row of guessed structural embeddings with the whole teacher table (a [1, N] matmul), keeps
the top-K scores, soft-maxes them and averages the K teacher rows -- a Python loop over nodes.  Here the same
thing runs for a block of nodes at a time: one [B, d] x [d, N] GEMM, one ``topk``, one soft-max, one gathered
weighted sum; ``block`` bounds the [B, N] score matrix (B x N x 4 bytes).
The result equals the loop's up to the order in which the K terms are added (the loop adds them by ascending
score, ``topk`` returns them descending).
"""
import torch
import torch.nn.functional as F


def replacement(teacherSE, le_guess, topK_2_replace, node_idx=None, block=4096):
    """[len(node_idx), d]: for node i, softmax(top-K of le_guess[i] . teacherSE^T) . teacherSE[top-K rows]."""
    le_guess = le_guess.detach()
    teacherSE = teacherSE.detach()
    if node_idx is None:
        node_idx = torch.arange(le_guess.shape[0], device=le_guess.device)
    node_idx = torch.as_tensor(node_idx, device=le_guess.device).long()
    table_t = teacherSE.t().contiguous()
    out = torch.empty((node_idx.numel(), teacherSE.shape[1]), dtype=teacherSE.dtype, device=teacherSE.device)
    k = min(int(topK_2_replace), teacherSE.shape[0])
    for b0 in range(0, node_idx.numel(), block):
        idx = node_idx[b0:b0 + block]
        scores = le_guess[idx] @ table_t                         # [B, N]
        top, sel = torch.topk(scores, k, dim=1)                  # K largest per row
        attn = F.softmax(top, dim=1)
        out[b0:b0 + idx.numel()] = torch.einsum('bk,bkd->bd', attn, teacherSE[sel])
    return out
