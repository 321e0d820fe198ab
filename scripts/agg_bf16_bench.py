"""Aggregation kernels in bf16 storage against fp32, at the cfg4 graph and at 128 columns (cfg5's width)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_tail_generalization_b200 import _cabi as C, graph as G, ops, synth
N = 10_000_000
ei = synth.powerlaw_graph(N, 45_000_000, seed=0, device='cuda')
g = G.GraphHandle(ei, N); E = ei.shape[1]; del ei
def t(fn, n=4):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for d in (256, 128):
    X = synth.features(N, d, 1, 'cuda'); x0 = synth.features(N, d, 2, 'cuda'); bias = torch.randn(d, device='cuda')
    for dt in (torch.float32, torch.bfloat16):
        Xd, x0d = X.to(dt), x0.to(dt)
        tg = t(lambda: ops.agg_gather_raw(g, C.CB_BY_SRC, Xd))
        tf = t(lambda: ops.agg_forward_raw(g, Xd, bias, x0d, 0.1, True, True, True, True))
        es = Xd.element_size()
        print(f'd={d} {str(dt)[6:]}: gather {tg:.2f} ms ({E / tg / 1e6:.1f} G edges/s, no-reuse model {E * (d * es + 4) / tg / 1e6 + N * d * es / tg / 1e6:.0f} GB/s), '
              f'fused forward {tf:.2f} ms ({E / tf / 1e6:.1f} G edges/s)', flush=True)
        del Xd, x0d
    del X, x0
