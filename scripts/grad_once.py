import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_tail_generalization_b200 import ops
M, N, K = 4_000_000, 256, 256
A = torch.randn(M, K, device='cuda'); W = torch.randn(N, K, device='cuda') / 16
wt = ops.split_weight(W, False)
rs = torch.rand(M, device='cuda'); mask = (torch.rand(M, N, device='cuda') > 0.5).to(torch.uint8)
dx0 = torch.randn(M, N, device='cuda')
for _ in range(2):
    ops.gemm_rows_grad_raw(A, wt)
    ops.gemm_rows_grad_raw(A, wt, row_scale=rs, gate_u8=mask, mixed=True, alpha=0.1, d_x0=dx0, accumulate_x0=True, post_scale=rs, want_col_sum=True)
torch.cuda.synchronize()
