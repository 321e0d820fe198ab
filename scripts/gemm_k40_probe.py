import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_tail_generalization_b200 import ops
g = torch.Generator(device='cuda').manual_seed(0)
M = 169343
for K in (32, 40, 64, 72, 256):
    for scale in (1.0, 1e-5):
        A = torch.randn(M, K, device='cuda', generator=g) * scale
        W = torch.randn(256, K, device='cuda', generator=g) / K ** 0.5
        wt = ops.split_weight(W, transpose=False)
        out = ops.gemm_rows_raw(A, wt)
        ref = A.double() @ W.double().t()
        sab = A.double().abs() @ W.double().abs().t()
        err = (out.double() - ref).abs()
        # transposed-split weight (what the adjoint uses)
        wt2 = ops.split_weight(W.t().contiguous(), transpose=True)
        out2 = ops.gemm_rows_raw(A, wt2)
        err2 = (out2.double() - ref).abs()
        print(f'K={K} scale={scale}: max err/max|ref| {float(err.max() / ref.abs().max()):.2e}  max err/sum|ab| {float((err / sab).max()):.2e}'
              f'   transposed split: {float(err2.max() / ref.abs().max()):.2e}', flush=True)
