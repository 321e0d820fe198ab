"""Row-sparse transposed gather at the bench shape: 10% live rows (the first tenth of the nodes)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_tail_generalization_b200 import _cabi as C, graph as G, ops, synth
N, d = 10_000_000, 256
ei = synth.powerlaw_graph(N, 45_000_000, seed=0, device='cuda')
g = G.GraphHandle(ei, N); del ei
X = synth.features(N, d, 1, 'cuda'); X[N // 10:] = 0
live = torch.zeros(N, dtype=torch.uint8, device='cuda'); live[:N // 10] = 1
def t(fn, n=4):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): out = fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n, out
td, od = t(lambda: ops.agg_gather_raw(g, C.CB_BY_SRC, X))
ts, os_ = t(lambda: ops.agg_gather_raw(g, C.CB_BY_SRC, X, live=live))
print(f'cfg {os.environ.get("CB_AGG_LIVE_CFG", "1")}: dense {td:.2f} ms, row-sparse {ts:.2f} ms, identical {torch.equal(od, os_)}')
