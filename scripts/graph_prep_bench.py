"""Device-side graph preparation (graph_prep) at the ogbn-arxiv and cfg4 sizes; the reference's own
Python loops (utils.py:300-334, 732-752) timed on a sample beside it and extrapolated per edge."""
import os, sys, time
from types import SimpleNamespace
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_tail_generalization_b200 import graph_prep as P, synth

def ref_graph_analyze_loop(n, ei):          # the loop of utils.py:300-334, as written there (dict of lists)
    ei = ei.cpu().numpy(); d = {}
    for ie in range(ei.shape[1]):
        o, t = ei[:, ie]
        if not d.get(o): d[o] = [0, 0]
        if not d.get(t): d[t] = [0, 0]
        d[o][0] += 1; d[t][1] += 1
    return d

for n, und in ((169_343, 1_157_799), (10_000_000, 45_000_000)):
    ei = synth.powerlaw_graph(n, und, seed=0, device='cuda')[:, : 2 * und]        # without the self loops
    half = ei[:, :und].contiguous()                                                # one direction only
    torch.cuda.synchronize(); t0 = time.perf_counter()
    sym = P.ensure_symmetric(half)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    do, dd = P.graph_analyze(n, sym)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    data = SimpleNamespace(x=torch.zeros(n, 1, device='cuda'), edge_index=torch.cat([sym, torch.arange(n, device='cuda').repeat(2, 1)], 1))
    P.save_graph_analyze(n, data, 1)
    torch.cuda.synchronize(); t3 = time.perf_counter()
    sample = sym[:, :200_000]
    t4 = time.perf_counter(); ref_graph_analyze_loop(n, sample); t5 = time.perf_counter()
    per_edge = (t5 - t4) / sample.shape[1]
    print(f'N={n} E={sym.shape[1]}: ensure_symmetric {1e3 * (t1 - t0):.1f} ms, graph_analyze {1e3 * (t2 - t1):.2f} ms, '
          f'save_graph_analyze(special split + isolation) {1e3 * (t3 - t2):.1f} ms on the device; the reference\'s '
          f'graph_analyze loop: {per_edge * 1e6:.2f} us/edge on the host = {per_edge * sym.shape[1]:.0f} s at this size', flush=True)
    del ei, half, sym, data
