"""Device-side graph preparation (graph_prep on the cb_prep_* kernels) at the ogbn-arxiv and cfg4 sizes;
beside it the same functions as torch tensor programs on the same GPU (the oracle restatement, round 1's
implementation) and the reference's own Python loops (utils.py:300-334, 732-752) timed on a sample and extrapolated
per edge."""
import os, sys, time
from types import SimpleNamespace
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_tail_generalization_b200 import graph_prep as P, synth
from oracle import graph_prep_oracle as PO      # timing comparator only

def ref_graph_analyze_loop(n, ei):          # the loop of utils.py:300-334, as written there (dict of lists)
    ei = ei.cpu().numpy(); d = {}
    for ie in range(ei.shape[1]):
        o, t = ei[:, ie]
        if not d.get(o): d[o] = [0, 0]
        if not d.get(t): d[t] = [0, 0]
        d[o][0] += 1; d[t][1] += 1
    return d

def timed_all(mod, n, half):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    sym = mod.ensure_symmetric(half)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    mod.graph_analyze(n, sym)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    data = SimpleNamespace(x=torch.zeros(n, 1, device='cuda'), edge_index=torch.cat([sym, torch.arange(n, device='cuda').repeat(2, 1)], 1))
    mod.save_graph_analyze(n, data, 1)
    torch.cuda.synchronize(); t3 = time.perf_counter()
    return 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), sym, data.edge_index

for n, und in ((169_343, 1_157_799), (10_000_000, 45_000_000)):
    half0 = synth.powerlaw_graph(n, und, seed=0, device='cuda')[:, :und].contiguous()
    for mod in (P, PO):
        timed_all(mod, n, half0)                                               # warm-up (lazy init, allocator)
    a = timed_all(P, n, half0); b = timed_all(PO, n, half0)
    assert torch.equal(a[3], b[3]) and torch.equal(a[4], b[4])
    print(f'N={n}: kernels (cb_prep_*) ensure_symmetric {a[0]:.2f} ms, graph_analyze {a[1]:.2f} ms, save_graph_analyze '
          f'{a[2]:.2f} ms | torch tensor programs {b[0]:.2f} / {b[1]:.2f} / {b[2]:.2f} ms; identical results', flush=True)
    del a, b, half0
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
