"""The adjoint launches of the bench step, once each (for ncu)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_tail_generalization_b200 import ops
M, N = 10_000_000, 256
g = torch.Generator(device='cuda').manual_seed(0)
rs = torch.rand(M, device='cuda', generator=g)
mask = (torch.rand(M, N, device='cuda', generator=g) > 0.5).to(torch.uint8)
dx0 = torch.randn(M, N, device='cuda', generator=g)
A = torch.randn(M, 64, device='cuda', generator=g); A[M // 10:] = 0
wt = ops.split_weight(torch.randn(N, 64, device='cuda', generator=g) / 8, False)
kw = dict(gate_u8=mask, mixed=True, alpha=0.1, post_scale=rs, want_col_sum=True)
ops.gemm_rows_grad_raw(A, wt, d_x0=dx0, accumulate_x0=False, **kw)            # (1)
ops.gemm_rows_grad_raw(A, wt)                                                  # (1) plain
del A
A = torch.randn(M, 256, device='cuda', generator=g)
wt = ops.split_weight(torch.randn(N, 256, device='cuda', generator=g) / 16, False)
ops.gemm_rows_grad_raw(A, wt, d_x0=dx0, accumulate_x0=True, **kw)             # (2)
ops.gemm_rows_raw(A, wt)                                                       # forward, plain
torch.cuda.synchronize()
