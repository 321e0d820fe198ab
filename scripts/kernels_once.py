"""Launches each tensor-core kernel at the bench shapes (for single-kernel ncu captures).

    ncu --set full --clock-control none --import-source on --kernel-name regex:k_gemm -s 6 -c 3 \
        -o gpurun_out/prof_gemm python scripts/kernels_once.py
"""
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
mask = (torch.rand(M, N, device='cuda') > 0.5).to(torch.uint8)
dx0 = torch.randn(M, N, device='cuda')
wt = ops.split_weight(W, False)
for _ in range(3):   # launches per round: k_gemm_rows<256,0>, k_gemm_rows<256,1>, k_gemm_tn (+ small reduce kernels)
    ops.gemm_rows_raw(A, wt, row_scale=rs)
    ops.gemm_rows_grad_raw(A, wt, row_scale=rs, gate_u8=mask, mixed=True, alpha=0.1, d_x0=dx0, accumulate_x0=True,
                           post_scale=rs, want_col_sum=True)
    ops.gemm_tn_raw(A, B, a_row_scale=rs)
torch.cuda.synchronize()
