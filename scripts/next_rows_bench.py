"""Timing of the two "next" rows built on top of the path on one B200."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_tail_generalization_b200 import graph as G, label_propagation as LP, synth
from gnn_tail_generalization_b200.virtual_neighbors import replacement

def sync_time(fn, n=3):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n, out

# --- label propagation: 50 iterations of alpha * DAD @ r + (1 - alpha) * y, clamp ----------------------------------
for n, und, c in ((169_343, 1_157_799, 40), (10_000_000, 45_000_000, 64)):
    ei = synth.powerlaw_graph(n, und, seed=0, device='cuda')[:, : 2 * und].contiguous()       # undirected, no self loops
    g = G.GraphHandle(ei, n)
    labels = torch.randint(0, c, (n, 1), device='cuda'); idx = torch.arange(n // 2, device='cuda')
    t_ours, res = sync_time(lambda: LP.label_propagation(g, labels, idx, 0.8, 50))
    # the reference's formulation on the same GPU: edge-valued sparse matrix x dense (cuSPARSE through torch.sparse)
    deg = torch.bincount(ei[0], minlength=n).float(); dis = deg.pow(-0.5); dis[torch.isinf(dis)] = 0
    adj = torch.sparse_coo_tensor(ei, dis[ei[0]] * dis[ei[1]], (n, n)).coalesce().to_sparse_csr()
    y = torch.zeros(n, c, device='cuda'); y[idx] = torch.nn.functional.one_hot(labels[idx].reshape(-1), c).float()
    def ref():
        r = y.clone()
        for _ in range(50):
            r = torch.clamp(0.8 * (adj @ r) + 0.2 * y, 0, 1)
        return r
    t_ref, want = sync_time(ref)
    print(f'label propagation N={n} E={ei.shape[1]} c={c}, 50 iterations: gather kernel {t_ours * 1e3:.1f} ms, '
          f'edge-valued cuSPARSE SpMM {t_ref * 1e3:.1f} ms, max |diff| {float((res - want).abs().max()):.2e}', flush=True)
    del ei, g, adj, y, res, want

# --- virtual-neighbour replacement ----------------------------------------------------------------------------------
n, d, k = 19_717, 256, 10
table = torch.randn(n, d, device='cuda'); guess = torch.randn(n, d, device='cuda')
t_b, out = sync_time(lambda: replacement(table, guess, k))
def loop(m=500):
    tt = table.t(); res = []
    for i in range(m):
        a = guess[[i]] @ tt; sel = a.argsort()[0][-k:]
        res.append(torch.softmax(a[:, sel], 1) @ table[sel])
    return torch.cat(res)
t_l, _ = sync_time(loop, 1)
print(f'virtual-neighbour replacement N={n} d={d} K={k}: batched {t_b * 1e3:.1f} ms for all nodes; the reference loop '
      f'{t_l / 500 * 1e3:.3f} ms per node = {t_l / 500 * n:.1f} s for all nodes on the same GPU', flush=True)
