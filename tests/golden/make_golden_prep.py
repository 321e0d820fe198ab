"""Fixtures for gnn_tail_generalization_b200/graph_prep.py, produced by the REFERENCE's own functions.

/root/reference/utils.py cannot be imported here (it imports dgl / torch_geometric at module level), so the
functions under test are cut out of its source by name with ``ast`` and executed unmodified in a namespace that
holds only numpy and torch.  Runs in the build container only; the .npz it writes is committed.

    python tests/golden/make_golden_prep.py
"""
import ast
import os
from types import SimpleNamespace

import numpy as np
import torch

REF = '/root/reference/utils.py'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'prep_cases.npz')
NAMES = ['graph_analyze', 'ensure_symmetric', 'get_partial_sorted_idx', 'craft_isolation_v2', 'save_graph_analyze',
         'gen_rec_for_table1_stats', 'tonp']


def load_reference_functions():
    src = open(REF).read()
    tree = ast.parse(src)
    ns = {'np': np, 'torch': torch, 'th': torch, 'plot_dist': lambda *a, **k: None, 'print': lambda *a, **k: None}
    ns['np'] = SimpleNamespace(**{k: getattr(np, k) for k in dir(np) if not k.startswith('__')})
    ns['np'].save = lambda *a, **k: None          # save_graph_analyze dumps a .npy into the cwd
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in NAMES:
            exec(compile(ast.Module([node], []), REF, 'exec'), ns)
    return ns


def main():
    ref = load_reference_functions()
    out = {}
    g = torch.Generator().manual_seed(0)
    for case, (n, e) in enumerate([(40, 120), (300, 1500), (1000, 3000), (64, 64)]):
        w = torch.arange(1, n + 1, dtype=torch.float64).pow(-0.8)
        ei = torch.stack([torch.multinomial(w, e, True, generator=g), torch.multinomial(w, e, True, generator=g)])
        ei[:, 0] = n - 1                                         # the largest id is present (N = max + 1)
        out[f'c{case}/edge_index'] = ei.numpy()
        do, dd = ref['graph_analyze'](n, ei)
        out[f'c{case}/degs_ori'], out[f'c{case}/degs_dst'] = np.asarray(do), np.asarray(dd)
        sym = ref['ensure_symmetric'](ei)
        out[f'c{case}/symmetric'] = sym.numpy()
        for mode in ['top50', 'bottom50', 'top25', 'bottom25', 'top12', 'bottom12', 'top6', 'bottom6', 'top3', 'bottom3']:
            arr = np.asarray(dd)
            # the reference returns None for the 50 % modes (no branch assigns a return for them): call guarded
            res = ref['get_partial_sorted_idx'](arr, mode)
            if res is not None:
                out[f'c{case}/partial_{mode}'] = np.asarray(res)
        for special in (0, 1):
            loops = torch.arange(n).repeat(2, 1)
            data = SimpleNamespace(x=torch.zeros(n, 1), edge_index=torch.cat([sym, loops], 1))
            ref['save_graph_analyze'](n, data, special)
            out[f'c{case}/s{special}/small_idx'] = np.asarray(data.small_deg_idx)
            out[f'c{case}/s{special}/large_idx'] = np.asarray(data.large_deg_idx)
            if special:
                out[f'c{case}/s{special}/zero_idx'] = np.asarray(data.zero_deg_idx)
                out[f'c{case}/s{special}/crafted'] = data.edge_index.numpy()
    np.savez_compressed(OUT, **out)
    print(f'wrote {OUT}: {len(out)} arrays')


if __name__ == '__main__':
    main()
