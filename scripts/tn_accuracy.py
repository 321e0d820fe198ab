import sys, os, subprocess, torch
sys.path.insert(0, '/root/repo')
if len(sys.argv) > 1:
    from gnn_tail_generalization_b200 import ops
    g = torch.Generator(device='cuda').manual_seed(0)
    M = 169_343
    A = torch.randn(M, 256, device='cuda', generator=g) * torch.exp(3 * torch.randn(M, 1, device='cuda', generator=g))
    B = torch.randn(M, 128, device='cuda', generator=g)
    ref = A.double().t() @ B.double()
    sab = A.abs().double().t() @ B.abs().double()
    for name, out in (('ours', ops.gemm_tn_raw(A, B)), ('cublas', A.t() @ B)):
        err = (out.double() - ref).abs()
        print(f'seg={os.environ.get("CB_TN_SEG_CHUNKS")} {name}: max err/max|ref| {float(err.max() / ref.abs().max()):.3e}  max err/sum|ab| {float((err / sab).max()):.3e}')
    A = torch.rand(M, 256, device='cuda', generator=g); B = torch.rand(M, 128, device='cuda', generator=g)
    ref = A.double().t() @ B.double()
    for name, out in (('ours', ops.gemm_tn_raw(A, B)), ('cublas', A.t() @ B)):
        print(f'   positive data {name}: max rel err {float(((out.double() - ref).abs() / ref).max()):.3e}')
else:
    for s in ('4', '16', '72', '1000'):
        subprocess.run([sys.executable, __file__, 'x'], env=dict(os.environ, CB_TN_SEG_CHUNKS=s))
