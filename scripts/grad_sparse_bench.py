"""The three adjoint launches of the bench step (cfg4) one by one, with and without the row-sparsity flags."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_tail_generalization_b200 import ops
M, N = 10_000_000, 256
def t(fn, n=4):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
g = torch.Generator(device='cuda').manual_seed(0)
rs = torch.rand(M, device='cuda', generator=g)
mask = (torch.rand(M, N, device='cuda', generator=g) > 0.5).to(torch.uint8)
dx0 = torch.randn(M, N, device='cuda', generator=g)
# (1) output-head adjoint + layer-1 prologue: K = 64, 10 % live rows, first writer of d_x0
A = torch.randn(M, 64, device='cuda', generator=g); A[M // 10:] = 0
wt = ops.split_weight(torch.randn(N, 64, device='cuda', generator=g) / 8, False)
flags = ops.row_any_nonzero_raw(A)
kw = dict(gate_u8=mask, mixed=True, alpha=0.1, post_scale=rs, want_col_sum=True)
print('(1) head adjoint, dense            %.2f ms' % t(lambda: ops.gemm_rows_grad_raw(A, wt, d_x0=dx0, accumulate_x0=False, **kw)))
print('(1) head adjoint, a_live           %.2f ms' % t(lambda: ops.gemm_rows_grad_raw(A, wt, d_x0=dx0, accumulate_x0=False, a_live=flags, **kw)))
print('    row_any_nonzero [M, 64]        %.2f ms' % t(lambda: ops.row_any_nonzero_raw(A)))
print('    kernel row_live (round 1)      %.2f ms' % t(lambda: ops.gemm_rows_grad_raw(A, wt, d_x0=dx0, accumulate_x0=False, row_live=torch.zeros(M, dtype=torch.uint8, device='cuda'), **kw)))
del A
# (2) GCNConv-1 adjoint + layer-0 prologue: K = 256, dense, accumulating d_x0
A = torch.randn(M, 256, device='cuda', generator=g)
wt = ops.split_weight(torch.randn(N, 256, device='cuda', generator=g) / 16, False)
print('(2) conv adjoint, accumulate       %.2f ms' % t(lambda: ops.gemm_rows_grad_raw(A, wt, d_x0=dx0, accumulate_x0=True, **kw)))
print('(2) conv adjoint, x0_valid         %.2f ms' % t(lambda: ops.gemm_rows_grad_raw(A, wt, d_x0=dx0, accumulate_x0=True, x0_valid=flags, **kw)))
# (3) GCNConv-0 adjoint + input-Linear relu/bias backward
gate = torch.randn(M, N, device='cuda', generator=g)
print('(3) conv adjoint + relu/bias       %.2f ms' % t(lambda: ops.gemm_rows_grad_raw(A, wt, row_scale=rs, add=dx0, gate_f32=gate, want_col_sum=True)))
print('    plain gemm_rows K=256          %.2f ms' % t(lambda: ops.gemm_rows_raw(A, wt)))
