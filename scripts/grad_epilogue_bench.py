"""Which part of the gradient epilogue of cb_gemm_rows_grad costs what (bench shapes, one B200)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_tail_generalization_b200 import ops  # noqa: E402

M, N = 10_000_000, 256


def t(fn, n=4):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for K in (256, 64):
    A = torch.randn(M, K, device='cuda')
    W = torch.randn(N, K, device='cuda') / K ** 0.5
    wt = ops.split_weight(W, False)
    rs = torch.rand(M, device='cuda')
    mask = (torch.rand(M, N, device='cuda') > 0.5).to(torch.uint8)
    gate32 = torch.randn(M, N, device='cuda')
    dx0 = torch.randn(M, N, device='cuda')
    print(f'K={K}')
    print('  plain gemm_rows            %.2f ms' % t(lambda: ops.gemm_rows_raw(A, wt)))
    print('  grad, nothing              %.2f ms' % t(lambda: ops.gemm_rows_grad_raw(A, wt)))
    print('  grad + col_sum             %.2f ms' % t(lambda: ops.gemm_rows_grad_raw(A, wt, want_col_sum=True)))
    print('  grad + row/post scale      %.2f ms' % t(lambda: ops.gemm_rows_grad_raw(A, wt, row_scale=rs, post_scale=rs)))
    print('  grad + gate u8             %.2f ms' % t(lambda: ops.gemm_rows_grad_raw(A, wt, gate_u8=mask)))
    print('  grad + gate f32            %.2f ms' % t(lambda: ops.gemm_rows_grad_raw(A, wt, gate_f32=gate32)))
    print('  grad + d_x0 write          %.2f ms' % t(lambda: ops.gemm_rows_grad_raw(A, wt, d_x0=dx0, accumulate_x0=False, mixed=True, alpha=0.1)))
    print('  grad + d_x0 accumulate     %.2f ms' % t(lambda: ops.gemm_rows_grad_raw(A, wt, d_x0=dx0, accumulate_x0=True, mixed=True, alpha=0.1)))
    print('  grad, all (prep layer 0)   %.2f ms' % t(lambda: ops.gemm_rows_grad_raw(A, wt, row_scale=rs, gate_u8=mask, mixed=True, alpha=0.1, d_x0=dx0, accumulate_x0=True, post_scale=rs, want_col_sum=True)), flush=True)
    del A, mask, gate32, dx0
