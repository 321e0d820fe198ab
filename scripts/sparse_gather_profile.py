"""Row-sparse transposed gather at the bench shape (10 % live rows): timings with CUDA events, or -- under ncu -- the
kernels to capture.  CB_SPARSE_LEAN=0 selects the general kernel over the compacted lists."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_tail_generalization_b200 import _cabi as C, graph as G, ops, synth
N, d = 10_000_000, 256
ei = synth.powerlaw_graph(N, 45_000_000, seed=0, device='cuda')
g = G.GraphHandle(ei, N); del ei
X = synth.features(N, d, 1, 'cuda'); X[N // 10:] = 0
live = torch.zeros(N, dtype=torch.uint8, device='cuda'); live[:N // 10] = 1
ws = ops.compact_live_raw(g, C.CB_BY_SRC, live)
out = torch.empty(N, d, device='cuda')
def t(fn, n=4):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
ts = t(lambda: ops.agg_gather_raw(g, C.CB_BY_SRC, X, live_ws=ws, out=out))
tc = t(lambda: ops.compact_live_raw(g, C.CB_BY_SRC, live))
print(f'CB_SPARSE_LEAN={os.environ.get("CB_SPARSE_LEAN", "1")}: compacted gather {ts:.2f} ms, compaction {tc:.2f} ms, '
      f'hub chunks {g.num_hub_chunks}', flush=True)
