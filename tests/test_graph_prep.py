"""Graph preparation against fixtures produced by the reference's own utils.py functions
(tests/golden/make_golden_prep.py): same values and the same ORDER.

CPU (-m "not gpu"): the oracle restatement (oracle/graph_prep_oracle.py) against the fixtures.
GPU (-m gpu): the product -- gnn_tail_generalization_b200/graph_prep.py on the cb_prep_* kernels -- against the same
fixtures directly, and against the oracle on larger random graphs and the edge cases (empty lists, ids the node count
does not cover, duplicates, all-equal degrees).  Index work: every comparison is exact.
"""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import graph_prep_oracle as PO

Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'prep_cases.npz'))
CASES = sorted({k.split('/')[0] for k in Z.files})
MODES = ['top50', 'bottom50', 'top25', 'bottom25', 'top12', 'bottom12', 'top6', 'bottom6', 'top3', 'bottom3']


def _product():
    from gnn_tail_generalization_b200 import graph_prep
    return graph_prep


def _run_all(P, ei, device):
    ei = ei.to(device)
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
        out[f's{special}/stats'] = torch.tensor(stats, dtype=torch.float64)
        assert stats[0] == n
        if special:
            assert stats[1] == data.edge_index_bkup.shape[1]
            for name in ('zero', 'small', 'large'):
                out[f's{special}/{name}_mask'] = getattr(data, f'{name}_deg_mask')
        out[f's{special}/small_idx'], out[f's{special}/large_idx'] = data.small_deg_idx, data.large_deg_idx
        if special:
            out[f's{special}/zero_idx'], out[f's{special}/crafted'] = data.zero_deg_idx, data.edge_index
    return {k: v.cpu().numpy() for k, v in out.items()}


def _check_against_fixture(P, case, device):
    ei = torch.from_numpy(Z[f'{case}/edge_index'])
    got = _run_all(P, ei, device)
    want = {k[len(case) + 1:]: Z[k] for k in Z.files if k.startswith(case + '/') and not k.endswith('edge_index')}
    assert set(want) <= set(got)
    n = len(got['degs_dst'])
    sym = PO.ensure_symmetric(ei)
    full = torch.cat([sym, torch.arange(n).repeat(2, 1)], 1)      # the graph save_graph_analyze was given
    deg = PO.graph_analyze(n, full)[1].numpy()
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
    data = SimpleNamespace(edge_index=full.to(device), zero_deg_mask=mask.to(device))
    P.craft_isolation_v2(data)
    assert np.array_equal(data.edge_index.cpu().numpy(), want['s1/crafted'])
    return got


@pytest.mark.parametrize('case', CASES)
def test_oracle_matches_reference_fixtures(case):
    _check_against_fixture(PO, case, 'cpu')


def test_oracle_edge_cases():
    ei = torch.tensor([[0, 2, 2, 1], [2, 0, 2, 1]])
    do, dd = PO.graph_analyze(4, ei)                       # node 3 does not occur
    assert do.tolist() == [1, 1, 2, 0] and dd.tolist() == [1, 1, 2, 0]
    assert PO.ensure_symmetric(ei).tolist() == [[0, 1, 2, 2], [2, 1, 0, 2]]
    data = SimpleNamespace(edge_index=ei, zero_deg_mask=torch.tensor([True, False, False, False]))
    assert PO.craft_isolation_v2(data) == 2                # (0,2) and (2,0) go; self loops stay
    assert data.edge_index.tolist() == [[2, 1], [2, 1]] and data.edge_index_bkup is ei


def test_product_has_no_host_path():
    P = _product()
    with pytest.raises(ValueError):
        P.graph_analyze(4, torch.tensor([[0, 1], [1, 0]]))
    with pytest.raises(ValueError):
        P.ensure_symmetric(torch.tensor([[0, 1], [1, 0]]))
    with pytest.raises(ValueError):
        P.get_partial_sorted_idx(torch.arange(5))


# ------------------------------------------------------------------------------------------------------------------
# the CUDA product
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize('case', CASES)
def test_kernels_match_reference_fixtures(case):
    from gnn_tail_generalization_b200 import _cabi
    before = _cabi.launch_count()
    got = _check_against_fixture(_product(), case, 'cuda:0')
    assert _cabi.launch_count() > before, 'no kernel of libcoldbrew_b200.so ran'
    # and everything else the fixture does not hold (statistics record, masks, the tie-dependent halves) against the
    # oracle, which uses the same stable tie rule
    want = _run_all(PO, torch.from_numpy(Z[f'{case}/edge_index']), 'cpu')
    assert set(got) == set(want)
    for k in want:
        assert np.array_equal(got[k], want[k]), k


@pytest.mark.gpu
@pytest.mark.parametrize('n,e,skew,seed', [(5000, 60000, 0.9, 1), (200000, 2500000, 0.7, 2), (1000, 200000, 1.2, 3),
                                            (70000, 70000, 0.0, 4)])
def test_kernels_match_oracle_on_random_graphs(n, e, skew, seed):
    P = _product()
    g = torch.Generator().manual_seed(seed)
    w = torch.arange(1, n + 1, dtype=torch.float64).pow(-skew)
    ei = torch.stack([torch.multinomial(w, e, True, generator=g), torch.multinomial(w, e, True, generator=g)])
    ei = torch.randperm(n, generator=g)[ei]
    ei[:, 0] = n - 1
    got, want = _run_all(P, ei, 'cuda:0'), _run_all(PO, ei, 'cpu')
    assert set(got) == set(want)
    for k in want:
        assert np.array_equal(got[k], want[k]), k


@pytest.mark.gpu
def test_kernels_edge_cases():
    P = _product()
    from gnn_tail_generalization_b200 import _cabi
    dev = 'cuda:0'
    ei = torch.tensor([[0, 2, 2, 1], [2, 0, 2, 1]], device=dev)
    do, dd = P.graph_analyze(4, ei)                       # node 3 does not occur
    assert do.tolist() == [1, 1, 2, 0] and dd.tolist() == [1, 1, 2, 0]
    do, dd = P.graph_analyze(2, ei)                       # ids the node count does not cover are ignored
    assert do.tolist() == [1, 1] and dd.tolist() == [1, 1]
    assert P.ensure_symmetric(ei).tolist() == [[0, 1, 2, 2], [2, 1, 0, 2]]
    data = SimpleNamespace(edge_index=ei, zero_deg_mask=torch.tensor([True, False, False, False], device=dev))
    assert P.craft_isolation_v2(data) == 2                # (0,2) and (2,0) go; self loops stay
    assert data.edge_index.tolist() == [[2, 1], [2, 1]] and data.edge_index_bkup is ei
    # empty inputs
    empty = torch.zeros(2, 0, dtype=torch.int64, device=dev)
    do, dd = P.graph_analyze(3, empty)
    assert do.tolist() == [0, 0, 0] and dd.tolist() == [0, 0, 0]
    assert P.ensure_symmetric(empty).shape == (2, 0)
    assert P.get_partial_sorted_idx(torch.zeros(0, dtype=torch.int64, device=dev), 'top3').numel() == 0
    data = SimpleNamespace(edge_index=empty, zero_deg_mask=torch.zeros(3, dtype=torch.bool, device=dev))
    assert P.craft_isolation_v2(data) == 0 and data.edge_index.shape == (2, 0)
    # all-equal values: every level keeps everything, both ways
    same = torch.full((37,), 5, dtype=torch.int64, device=dev)
    for mode in MODES:
        assert P.get_partial_sorted_idx(same, mode).tolist() == list(range(37))
    # negative ids are an error, not an index from the end
    with pytest.raises(_cabi.ColdBrewError) as err:
        P.graph_analyze(4, torch.tensor([[0, -1], [1, 0]], device=dev))
    assert err.value.code == -2
    with pytest.raises(_cabi.ColdBrewError):
        P.ensure_symmetric(torch.tensor([[0, -1], [1, 0]], device=dev))
    # negative VALUES are fine in the selection (any integer array)
    vals = torch.tensor([3, -7, 0, 12, -7, 5, 1], device=dev)
    for mode in MODES:
        assert P.get_partial_sorted_idx(vals, mode).tolist() == PO.get_partial_sorted_idx(vals.cpu(), mode).tolist()
