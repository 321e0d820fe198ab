import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_tail_generalization_b200 import ops
torch.backends.cuda.matmul.allow_tf32 = False
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for (M, K, N) in [(10_000_000, 256, 256), (10_000_000, 256, 64), (10_000_000, 64, 256), (2_000_000, 128, 128)]:
    A = torch.randn(M, K, device='cuda'); W = torch.randn(N, K, device='cuda') / K ** 0.5
    wt = ops.split_weight(W, False)
    out = ops.gemm_rows_raw(A, wt)
    ref = A[:4096].double() @ W.double().t()
    err = float((out[:4096].double() - ref).abs().max()); errc = float(((A[:4096] @ W.t()).double() - ref).abs().max())
    ms = t(lambda: ops.gemm_rows_raw(A, wt)); msc = t(lambda: A @ W.t())
    if N == 256 and K == 256:
        add = torch.randn(M, N, device='cuda'); rs = torch.rand(M, device='cuda')
        print('   with row_scale+add: %.2f ms; with bias+relu: %.2f ms' % (t(lambda: ops.gemm_rows_raw(A, wt, row_scale=rs, add=add)), t(lambda: ops.gemm_rows_raw(A, wt, bias=rs[:N].contiguous(), relu=True))), flush=True)
        del add
    fl = 2 * M * K * N
    print(f'M={M} K={K} N={N}: ours {ms:.2f} ms ({fl/ms/1e9:.1f} TF/s fp32-equiv, {4*(M*K+M*N)/ms/1e6:.0f} GB/s) cublas fp32 {msc:.2f} ms  err ours {err:.2e} cublas {errc:.2e}', flush=True)
    del A, out

for (M, Ka, Nb) in [(10_000_000, 256, 256), (10_000_000, 256, 64)]:
    A = torch.randn(M, Ka, device='cuda'); B = torch.randn(M, Nb, device='cuda')
    out = ops.gemm_tn_raw(A, B)
    ref = A[:200000].double().t() @ B[:200000].double()
    o2 = ops.gemm_tn_raw(A[:200000], B[:200000])
    err = float((o2.double() - ref).abs().max()); errc = float(((A[:200000].t() @ B[:200000]).double() - ref).abs().max())
    ms = t(lambda: ops.gemm_tn_raw(A, B)); msc = t(lambda: A.t() @ B)
    fl = 2 * M * Ka * Nb
    print(f'TN M={M} Ka={Ka} Nb={Nb}: ours {ms:.2f} ms ({fl/ms/1e9:.1f} TF/s fp32-equiv, {4*(M*Ka+M*Nb)/ms/1e6:.0f} GB/s) cublas fp32 {msc:.2f} ms  err(200k rows) ours {err:.2e} cublas {errc:.2e}', flush=True)
    del A, B
