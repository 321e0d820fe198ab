"""Launches each hot kernel a few times at the bench shapes (for ncu captures of single kernels)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_tail_generalization_b200 import ops  # noqa: E402

M, K, N = 10_000_000, 256, 256
A = torch.randn(M, K, device='cuda')
B = torch.randn(M, N, device='cuda')
W = torch.randn(N, K, device='cuda') / 16
rs = torch.rand(M, device='cuda')
wt = ops.split_weight(W, False)
for _ in range(3):
    ops.gemm_rows_raw(A, wt, row_scale=rs)
    ops.gemm_tn_raw(A, B, a_row_scale=rs)
torch.cuda.synchronize()
