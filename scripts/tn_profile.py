import torch, sys
sys.path.insert(0, '/root/repo')
from gnn_tail_generalization_b200 import ops
from torch.profiler import profile, ProfilerActivity
M, Ka, Nb = 10_000_000, 256, 256
A = torch.randn(M, Ka, device='cuda'); B = torch.randn(M, Nb, device='cuda')
W = torch.randn(256, 256, device='cuda') / 16
wt = ops.split_weight(W, False)
for _ in range(2):
    ops.gemm_tn_raw(A, B); ops.gemm_rows_raw(A, wt)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        ops.gemm_tn_raw(A, B)
        ops.gemm_rows_raw(A, wt)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=8, max_name_column_width=60))
