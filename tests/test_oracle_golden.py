"""CPU: the oracle restatement against (1) the hand-derived KAT of SURVEY 3.3 and (2) fixtures produced by
running the reference's own GNN_model code (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import coldbrew_oracle as O
from tests.helpers import golden_args, golden_cases, load_golden, load_params

TOY = torch.tensor([[0, 0, 1, 1, 1, 2], [0, 1, 0, 1, 2, 2]])          # reference utils.py:1096


def test_kat_hand_derived():
    # dout=[2,3,1], din=[2,2,2]  ->  values computed by hand from GCN.py:205-253
    rst, reg = O.gcn_conv(torch.eye(3), TOY, 3, torch.eye(3), torch.zeros(3))
    want = torch.tensor([[0.5, 0.408248, 0.0], [0.5, 0.408248, 0.0], [0.0, 0.408248, 0.707107]])
    assert reg is None and torch.allclose(rst, want, atol=1e-6)
    le = torch.arange(9.).view(3, 3) / 10
    rst, reg = O.gcn_conv(torch.eye(3), TOY, 3, torch.eye(3), torch.zeros(3), le)
    want = torch.tensor([[0.712132, 0.761802, 0.494975], [0.712132, 0.761802, 0.494975],
                         [0.636396, 1.186066, 1.626346]])
    assert torch.allclose(rst, want, atol=1e-6) and abs(float(reg) - 1.428286) < 1e-6


def test_kat_reference_code():
    z = load_golden('kat_toy')
    rst, _ = O.gcn_conv(torch.eye(3), TOY, 3, torch.eye(3), torch.zeros(3))
    assert np.array_equal(rst.numpy(), z['rst_se0'])
    rst, reg = O.gcn_conv(torch.eye(3), TOY, 3, torch.eye(3), torch.zeros(3), torch.arange(9.).view(3, 3) / 10)
    assert np.array_equal(rst.numpy(), z['rst_se1']) and float(reg) == float(z['se_reg'])


def test_zero_in_degree_raises():
    ei = torch.tensor([[0, 1], [1, 1]])                              # node 0 has no in-edge
    with pytest.raises(RuntimeError):
        O.gcn_conv(torch.eye(2), ei, 2, torch.eye(2), None)
    O.gcn_conv(torch.eye(2), ei, 2, torch.eye(2), None, allow_zero_in_degree=True)


@pytest.mark.parametrize('name', golden_cases())
def test_oracle_matches_reference_fixture(name):
    z = load_golden(name)
    torch.set_num_threads(1)
    a = golden_args(z, O.make_args)
    model = O.OracleTeacherGNN(a, None)
    load_params(model, z)
    model.eval() if name.startswith('exact_batchnorm') else model.train()
    x, ei = torch.from_numpy(z['x']), torch.from_numpy(z['edge_index'])
    y, mask = torch.from_numpy(z['y']), torch.from_numpy(z['train_mask'])
    res = model.get_3_embs(x, ei, mask)
    # same ops in the same order on the same backend: bit-for-bit
    assert np.array_equal(res.emb4classi_full.detach().numpy(), z['logits'])
    if np.isnan(z['se_reg_all']):
        assert model.se_reg_all is None
    else:
        assert float(model.se_reg_all) == pytest.approx(float(z['se_reg_all']), rel=1e-6)
    loss = F.nll_loss(F.log_softmax(res.emb4classi, 1), y[mask])
    if int(z['reg_in_loss']):
        loss = loss + 0.5 * model.se_reg_all
    assert float(loss) == pytest.approx(float(z['loss']), rel=1e-6)
    loss.backward()
    grads = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
    want = {k[5:]: v for k, v in z.items() if k.startswith('grad/')}
    assert set(grads) == set(want)
    for k, v in want.items():
        np.testing.assert_allclose(grads[k].numpy(), v, rtol=1e-5, atol=1e-7, err_msg=k)
    xin = x if a.dim_learnable_input == 0 else model.embs
    _, _, les = model.model.model(xin, ei, want_les=True)
    assert np.array_equal(les.detach().numpy(), z['les'])


def test_csr_stable_and_c_oracle_bit_exact():
    ei = O.powerlaw_graph(3000, 12000, seed=3)
    n = 3000
    rowptr, cols, perm = O.build_csr(ei[1].numpy(), ei[0].numpy(), n)
    # stability: inside a row, stored order == original edge order
    for r in np.random.default_rng(0).integers(0, n, 50):
        seg = perm[rowptr[r]:rowptr[r + 1]]
        assert np.all(np.diff(seg) > 0)
    h = torch.randn(n, 40, generator=torch.Generator().manual_seed(0))
    a = O.aggregate_sum_csr_ordered(h.numpy(), rowptr, cols, threads=1)
    b = O.aggregate_sum_csr_ordered(h.numpy(), rowptr, cols, threads=4)
    c = O.aggregate_sum(h, ei, n).numpy()
    assert np.array_equal(a, b) and np.array_equal(a, c)
    # chunked association differs only on rows longer than the chunk
    d = O.aggregate_sum_csr_ordered(h.numpy(), rowptr, cols, hub_chunk=64)
    deg = np.diff(rowptr)
    assert np.array_equal(a[deg <= 64], d[deg <= 64])
    np.testing.assert_allclose(a, d, rtol=1e-5, atol=2e-4)   # hub rows: a few thousand terms, re-associated


def test_powerlaw_graph_is_canonical():
    n = 2000
    ei = O.powerlaw_graph(n, 9000, seed=1)
    src, dst = ei
    key = src * n + dst
    assert key.unique().numel() == key.numel()                       # no duplicate edges
    assert int((src == dst).sum()) == n                              # exactly one self loop per node
    assert set((dst * n + src).tolist()) == set(key.tolist())        # symmetric
    assert ei.shape[1] == 2 * 9000 + n
    assert not O.has_zero_in_degree(ei, n)


def test_edge_weighted_oracle_forms_agree():
    """GCN.py:199-202 (u_mul_e): unit weights give the unweighted layer; the in-order C restatement and the
    index_add_ restatement agree (to fp32 reassociation); the toy graph by hand."""
    import numpy as np
    rst, _ = O.gcn_conv(torch.eye(3), TOY, 3, torch.eye(3), torch.zeros(3))
    rst_w, _ = O.gcn_conv(torch.eye(3), TOY, 3, torch.eye(3), torch.zeros(3), edge_weight=torch.ones(TOY.shape[1]))
    assert torch.equal(rst, rst_w)
    g = torch.Generator().manual_seed(0)
    n, e, d = 300, 4000, 16
    ei = torch.randint(0, n, (2, e), generator=g)
    x, w = torch.randn(n, d, generator=g), torch.randn(e, generator=g)
    rp, cl, pm = O.build_csr(ei[1].numpy(), ei[0].numpy(), n)
    a = O.aggregate_mul_sum_csr_ordered(x.numpy(), rp, cl, w.numpy()[pm])
    b = O.aggregate_mul_sum(x, ei, w, n).numpy()
    assert np.allclose(a, b, rtol=1e-5, atol=1e-5)
    assert np.array_equal(O.aggregate_mul_sum_csr_ordered(x.numpy(), rp, cl, np.ones(e, np.float32)),
                          O.aggregate_sum_csr_ordered(x.numpy(), rp, cl))
    tiny = O.aggregate_mul_sum_csr_ordered(np.array([[1., 2.], [3., 4.]], np.float32), np.array([0, 2, 3]), np.array([1, 0, 1]),
                                           np.array([0.5, 2., 3.], np.float32))
    assert tiny.tolist() == [[3.5, 6.0], [9.0, 12.0]]
