"""graph_prep against fixtures produced by the reference's own utils.py functions
(tests/golden/make_golden_prep.py): same values and the same ORDER; on the GPU the same tensor programs must give
what they give on the CPU."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from gnn_tail_generalization_b200 import graph_prep as P

Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'prep_cases.npz'))
CASES = sorted({k.split('/')[0] for k in Z.files})
MODES = ['top50', 'bottom50', 'top25', 'bottom25', 'top12', 'bottom12', 'top6', 'bottom6', 'top3', 'bottom3']


def _run_all(case, device):
    ei = torch.from_numpy(Z[f'{case}/edge_index']).to(device)
    n = int(ei.max()) + 1
    out = {}
    out['degs_ori'], out['degs_dst'] = P.graph_analyze(n, ei)
    sym = P.ensure_symmetric(ei)
    out['symmetric'] = sym
    for mode in MODES:
        out[f'partial_{mode}'] = P.get_partial_sorted_idx(out['degs_dst'], mode)
    for special in (0, 1):
        loops = torch.arange(n, device=device).repeat(2, 1)
        data = SimpleNamespace(x=torch.zeros(n, 1, device=device), edge_index=torch.cat([sym, loops], 1))
        stats = P.save_graph_analyze(n, data, special)
        assert stats[0] == n and stats[1] == data.edge_index_bkup.shape[1] if special else True
        out[f's{special}/small_idx'], out[f's{special}/large_idx'] = data.small_deg_idx, data.large_deg_idx
        if special:
            out[f's{special}/zero_idx'], out[f's{special}/crafted'] = data.zero_deg_idx, data.edge_index
    return {k: v.cpu().numpy() for k, v in out.items()}


@pytest.mark.parametrize('case', CASES)
def test_graph_prep_matches_reference_fixtures(case):
    got = _run_all(case, 'cpu')
    want = {k[len(case) + 1:]: Z[k] for k in Z.files if k.startswith(case + '/') and not k.endswith('edge_index')}
    assert set(want) <= set(got)
    ei = torch.from_numpy(Z[f'{case}/edge_index'])
    n = len(got['degs_dst'])
    sym = P.ensure_symmetric(ei)
    full = torch.cat([sym, torch.arange(n).repeat(2, 1)], 1)      # the graph save_graph_analyze was given
    deg = P.graph_analyze(n, full)[1].numpy()
    for k, v in want.items():
        if k in ('s1/zero_idx', 's1/small_idx', 's1/crafted'):
            continue        # depend on how numpy's (unstable) argsort orders equal degrees: checked below
        assert np.array_equal(got[k], v), k
    # the special split halves the lowest-degree sixth after a sort by degree; equal degrees may land on either
    # side (numpy's default argsort is not stable, its tie order depends on the numpy build): same node set, same
    # degrees on each side, nothing of a lower degree on the "small" side than on the "isolated" side
    assert set(got['s1/zero_idx']) | set(got['s1/small_idx']) == set(want['s1/zero_idx']) | set(want['s1/small_idx'])
    assert len(got['s1/zero_idx']) == len(want['s1/zero_idx'])
    assert np.array_equal(np.sort(deg[got['s1/zero_idx']]), np.sort(deg[want['s1/zero_idx']]))
    assert deg[got['s1/zero_idx']].max() <= deg[got['s1/small_idx']].min()
    # given the reference's own choice of isolated nodes, the crafted edge list is identical, order included
    mask = torch.zeros(n, dtype=torch.bool)
    mask[torch.from_numpy(want['s1/zero_idx'])] = True
    data = SimpleNamespace(edge_index=full, zero_deg_mask=mask)
    P.craft_isolation_v2(data)
    assert np.array_equal(data.edge_index.numpy(), want['s1/crafted'])


def test_graph_prep_edge_cases():
    ei = torch.tensor([[0, 2, 2, 1], [2, 0, 2, 1]])
    do, dd = P.graph_analyze(4, ei)                       # node 3 does not occur
    assert do.tolist() == [1, 1, 2, 0] and dd.tolist() == [1, 1, 2, 0]
    assert P.ensure_symmetric(ei).tolist() == [[0, 1, 2, 2], [2, 1, 0, 2]]
    data = SimpleNamespace(edge_index=ei, zero_deg_mask=torch.tensor([True, False, False, False]))
    assert P.craft_isolation_v2(data) == 2                # (0,2) and (2,0) go; self loops stay
    assert data.edge_index.tolist() == [[2, 1], [2, 1]] and data.edge_index_bkup is ei


@pytest.mark.gpu
@pytest.mark.parametrize('case', CASES)
def test_graph_prep_gpu_equals_cpu(case):
    cpu, gpu = _run_all(case, 'cpu'), _run_all(case, 'cuda:0')
    for k in cpu:
        assert np.array_equal(cpu[k], gpu[k]), k
