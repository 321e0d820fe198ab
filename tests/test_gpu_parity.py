"""GPU parity: the CUDA path, called through the C ABI, against the CPU oracle (same seeded inputs),
the reference-generated golden fixtures, and size-independent properties at larger sizes.

Tolerances: integer/index work bit-exact; the bare aggregation bit-exact against the in-order C oracle
(same association, hub chunks included); fp32 model outputs within 1e-4 (north_star) -- observed ~1e-6.
"""
import numpy as np
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import coldbrew_oracle as O
from tests.helpers import golden_args, golden_cases, load_golden, load_params

pytestmark = pytest.mark.gpu

DEV = 'cuda:0'
LOGIT_TOL = 1e-4          # north_star: fp32 logits within 1e-4 of the reference path


def _pkg():
    from gnn_tail_generalization_b200 import _cabi, graph, ops
    return _cabi, graph, ops


def _multigraph(n, e, seed, skew=True):
    """Unsorted edge list with duplicates and self loops (DGL multigraph semantics)."""
    g = torch.Generator().manual_seed(seed)
    if skew:
        w = torch.arange(1, n + 1, dtype=torch.float64).pow(-0.9)
        src = torch.multinomial(w, e, replacement=True, generator=g)
        dst = torch.multinomial(w, e, replacement=True, generator=g)
        p = torch.randperm(n, generator=g)
        src, dst = p[src], p[dst]
    else:
        src = torch.randint(0, n, (e,), generator=g)
        dst = torch.randint(0, n, (e,), generator=g)
    return torch.stack([src, dst])


# ------------------------------------------------------------------------------------------------
# graph construction: bit-exact indexing
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('n,e,seed', [(1, 1, 0), (7, 0, 1), (50, 400, 2), (3000, 40000, 3), (100000, 1200000, 4)])
def test_graph_build_bit_exact(n, e, seed):
    C, G, _ = _pkg()
    ei = _multigraph(n, e, seed) if e else torch.zeros(2, 0, dtype=torch.int64)
    h = G.GraphHandle(ei.to(DEV), n, hub_chunk=64)
    for side, key, val in ((C.CB_BY_DST, ei[1], ei[0]), (C.CB_BY_SRC, ei[0], ei[1])):
        rowptr, col, perm = h.csr(side)
        rp, cl, pm = O.build_csr(key.numpy(), val.numpy(), n)
        assert np.array_equal(rowptr.cpu().numpy(), rp)
        assert np.array_equal(col.cpu().numpy().astype(np.int64), cl)
        assert np.array_equal(perm.cpu().numpy().astype(np.int64), pm)
    assert np.array_equal(h.in_degrees().cpu().numpy(), np.bincount(ei[1].numpy(), minlength=n))
    assert np.array_equal(h.out_degrees().cpu().numpy(), np.bincount(ei[0].numpy(), minlength=n))
    dout, din = O.degree_inv_sqrt(ei, n)
    assert torch.allclose(h.din_inv_sqrt.cpu(), din, rtol=2e-7, atol=0)
    assert torch.allclose(h.dout_inv_sqrt.cpu(), dout, rtol=2e-7, atol=0)
    assert h.has_zero_in_degree == O.has_zero_in_degree(ei, n)
    assert h.number_of_edges() == e and h.number_of_nodes() == n
    h.close()


def test_graph_build_sliced_matches_whole():
    C, G, _ = _pkg()
    n, e = 5000, 60000
    ei = _multigraph(n, e, 11)
    rp, cl, pm = O.build_csr(ei[1].numpy(), ei[0].numpy(), n)
    for lo, hi in ((0, 1250), (1250, 3000), (3000, 5000), (777, 777)):
        h = G.GraphHandle(ei.to(DEV), n, row_begin=lo, row_end=hi, hub_chunk=32)
        rowptr, col, perm = h.csr(C.CB_BY_DST)
        assert np.array_equal(rowptr.cpu().numpy(), rp[lo:hi + 1] - rp[lo])
        assert np.array_equal(col.cpu().numpy(), cl[rp[lo]:rp[hi]])
        assert np.array_equal(perm.cpu().numpy(), pm[rp[lo]:rp[hi]])
        assert h.rows == hi - lo and h.num_edges == rp[hi] - rp[lo]


def test_graph_build_rejects_bad_ids():
    C, G, _ = _pkg()
    ei = torch.tensor([[0, 1, 5], [1, 0, 2]])
    with pytest.raises(C.ColdBrewError) as err:
        G.GraphHandle(ei.to(DEV), 5)
    assert err.value.code == -2
    with pytest.raises(C.ColdBrewError):
        G.GraphHandle(torch.tensor([[0, -1], [1, 0]]).to(DEV), 5)


# ------------------------------------------------------------------------------------------------
# the bare aggregation: bit-exact against the in-order C oracle
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('d', [1, 2, 3, 4, 7, 8, 10, 16, 20, 40, 64, 100, 128, 130, 256, 512, 700])
def test_aggregate_bit_exact_all_widths(d):
    C, G, ops = _pkg()
    n, e, hub = 2000, 30000, 48
    ei = _multigraph(n, e, 100 + d)
    x = torch.randn(n, d, generator=torch.Generator().manual_seed(d))
    h = G.GraphHandle(ei.to(DEV), n, hub_chunk=hub)
    assert h.num_hub_chunks[0] > 0 and h.num_hub_chunks[1] > 0
    for side, key, val in ((C.CB_BY_DST, ei[1], ei[0]), (C.CB_BY_SRC, ei[0], ei[1])):
        rp, cl, _ = O.build_csr(key.numpy(), val.numpy(), n)
        want = O.aggregate_sum_csr_ordered(x.numpy(), rp, cl, hub_chunk=hub)
        got = ops.agg_gather_raw(h, side, x.to(DEV)).cpu().numpy()
        assert np.array_equal(got, want)


@pytest.mark.parametrize('n,e,d,hub', [(1, 1, 8, 0), (10, 0, 16, 0), (300, 300 * 40, 64, 16), (200000, 2400000, 256, 0),
                                        (50000, 900000, 128, 128)])
def test_aggregate_bit_exact_shapes(n, e, d, hub):
    C, G, ops = _pkg()
    ei = _multigraph(n, e, n + e) if e else torch.zeros(2, 0, dtype=torch.int64)
    x = torch.randn(n, d, generator=torch.Generator().manual_seed(5))
    h = G.GraphHandle(ei.to(DEV), n, hub_chunk=hub)
    rp, cl, _ = O.build_csr(ei[1].numpy(), ei[0].numpy(), n)
    want = O.aggregate_sum_csr_ordered(x.numpy(), rp, cl, hub_chunk=h.hub_chunk)
    got = ops.agg_gather_raw(h, C.CB_BY_DST, x.to(DEV)).cpu().numpy()
    assert np.array_equal(got, want)
    # run-to-run determinism (no atomics anywhere on the path)
    assert torch.equal(ops.agg_gather_raw(h, C.CB_BY_DST, x.to(DEV)).cpu(), torch.from_numpy(got))


def test_isolated_rows_get_bias_and_zero():
    C, G, ops = _pkg()
    ei = torch.tensor([[0, 1], [1, 1]])                      # node 0 and 2 have no in-edge
    h = G.GraphHandle(ei.to(DEV), 3)
    assert h.has_zero_in_degree
    x = torch.arange(12.).view(3, 4).to(DEV)
    b = torch.tensor([1., 2., 3., 4.]).to(DEV)
    out, _, _ = ops.agg_forward_raw(h, x, bias=b)
    assert torch.equal(out[0], b) and torch.equal(out[2], b)
    want1 = (x[0] + x[1]) * (2 ** -0.5) + b
    assert torch.allclose(out[1], want1, rtol=1e-6)


# ------------------------------------------------------------------------------------------------
# fused epilogue and the backward kernels against the oracle's op-by-op arithmetic
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('d', [5, 16, 64, 256])
@pytest.mark.parametrize('relu,mix', [(False, False), (True, False), (True, True), (False, True)])
def test_fused_forward_matches_oracle_chain(d, relu, mix):
    C, G, ops = _pkg()
    n, e = 1500, 20000
    ei = O.canonicalize_planetoid(_multigraph(n, e, d), n)
    gen = torch.Generator().manual_seed(d)
    hmat, bias, x0 = torch.randn(n, d, generator=gen), torch.randn(d, generator=gen), torch.randn(n, d, generator=gen)
    alpha = 0.1
    h = G.GraphHandle(ei.to(DEV), n)
    out, out_s, mask = ops.agg_forward_raw(h, hmat.to(DEV), bias.to(DEV), x0.to(DEV) if mix else None, alpha, relu,
                                           True, True, True)
    dout_is, din_is = O.degree_inv_sqrt(ei, n)
    z = O.aggregate_sum(hmat, ei, n) * din_is[:, None] + bias
    r = F.relu(z) if relu else z
    want = (1 - alpha) * r + alpha * x0 if mix else r
    # the aggregation is bit-exact and the epilogue rounds like the op-by-op chain; the only slack is
    # the last bit of degree^-1/2 (computed in double on the device)
    assert torch.allclose(out.cpu(), want, rtol=1e-6, atol=3e-6)
    assert torch.allclose(out_s.cpu(), want * dout_is[:, None], rtol=1e-6, atol=3e-6)
    zc = z.abs() > 1e-5
    assert torch.equal(mask.cpu().bool()[zc], (z > 0)[zc])


@pytest.mark.parametrize('d', [6, 32, 256])
@pytest.mark.parametrize('relu,mix', [(False, False), (True, False), (True, True)])
def test_fused_aggregate_autograd_matches_oracle(d, relu, mix):
    C, G, ops = _pkg()
    n, e, alpha = 800, 9000, 0.2
    ei = O.canonicalize_planetoid(_multigraph(n, e, 7 * d), n)
    gen = torch.Generator().manual_seed(d)
    base = [torch.randn(n, d, generator=gen), torch.randn(d, generator=gen), torch.randn(n, d, generator=gen)]
    wout, wsc = torch.randn(n, d, generator=gen), torch.randn(n, d, generator=gen)
    dout_is, din_is = O.degree_inv_sqrt(ei, n)

    def oracle(hm, b, x0):
        z = O.aggregate_sum(hm, ei, n) * din_is[:, None] + b
        r = F.relu(z) if relu else z
        o = (1 - alpha) * r + alpha * x0 if mix else r
        return o, o * dout_is[:, None]

    cpu = [t.clone().requires_grad_() for t in base]
    o, os_ = oracle(*cpu)
    ((o * wout).sum() + (os_ * wsc).sum()).backward()

    h = G.GraphHandle(ei.to(DEV), n)
    gpu = [t.clone().to(DEV).requires_grad_() for t in base]
    o2, os2 = ops.fused_aggregate(gpu[0], h, gpu[1], gpu[2] if mix else None, alpha, relu, True, True)
    ((o2 * wout.to(DEV)).sum() + (os2 * wsc.to(DEV)).sum()).backward()
    assert torch.allclose(o2.detach().cpu(), o.detach(), rtol=1e-6, atol=1e-6)
    for a, b_, name in zip(gpu, cpu, 'H bias x0'.split()):
        if name == 'x0' and not mix:
            assert a.grad is None
            continue
        scale = float(b_.grad.abs().max()) + 1e-12
        assert float((a.grad.cpu() - b_.grad).abs().max()) <= 2e-6 * scale + 1e-6, name

    # only the scaled output consumed (the fused layer-to-layer hand-off)
    gpu = [t.clone().to(DEV).requires_grad_() for t in base]
    _, os3 = ops.fused_aggregate(gpu[0], h, gpu[1], gpu[2] if mix else None, alpha, relu, False, True)
    (os3 * wsc.to(DEV)).sum().backward()
    cpu = [t.clone().requires_grad_() for t in base]
    (oracle(*cpu)[1] * wsc).sum().backward()
    assert torch.allclose(gpu[0].grad.cpu(), cpu[0].grad, rtol=1e-5, atol=1e-5)


def test_row_sparse_hint_on_the_unfused_backward_changes_nothing():
    """The layer under the output head when dropout sits in between (no hand-off plan): with the hint the backward
    looks for the all-zero rows of G and gathers over the compacted lists -- same gradients, bit for bit."""
    C, G, ops = _pkg()
    n, e, d, alpha = 6000, 80000, 64, 0.1
    ei = O.canonicalize_planetoid(_multigraph(n, e, 321), n)
    gen = torch.Generator().manual_seed(9)
    base = [torch.randn(n, d, generator=gen), torch.randn(d, generator=gen), torch.randn(n, d, generator=gen)]
    w = torch.randn(n, d, generator=gen)
    w[n // 10:] = 0                                  # the loss reads the first tenth of the rows only
    h = G.GraphHandle(ei.to(DEV), n, hub_chunk=64)
    grads, names = [], []
    for hint in (False, True):
        gpu = [t.clone().to(DEV).requires_grad_() for t in base]
        sink = []
        ops.set_timing_sink(sink)
        out, _ = ops.fused_aggregate(gpu[0], h, gpu[1], gpu[2], alpha, True, True, False, row_sparse_hint=hint)
        (torch.nn.functional.dropout(out, 0.5, training=True) * w.to(DEV)).sum().backward()
        ops.set_timing_sink(None)
        names.append([x[0] for x in sink])
        grads.append([t.grad.clone() for t in gpu])
        torch.manual_seed(0)
    assert 'agg_gather_src' in names[0] and 'agg_gather_src_rowsparse' not in names[0]
    # the prologue flags the live rows of G while it writes them: no separate pass
    assert 'agg_gather_src_rowsparse' in names[1] and 'row_any_nonzero' not in names[1]
    # the dropout masks differ between the two runs (the RNG moved on): compare what does not depend on them --
    # rows of dH that only gather from dead rows are zero in both, and a deterministic rerun matches bit for bit
    for hint in (False, True):
        res = []
        for _ in range(2):
            torch.manual_seed(11)
            gpu = [t.clone().to(DEV).requires_grad_() for t in base]
            out, _ = ops.fused_aggregate(gpu[0], h, gpu[1], gpu[2], alpha, True, True, False, row_sparse_hint=hint)
            (torch.nn.functional.dropout(out, 0.5, training=True) * w.to(DEV)).sum().backward()
            res.append([t.grad.clone() for t in gpu])
        grads.append(res[0])
        for a, b in zip(*res):
            assert torch.equal(a, b)
    for a, b in zip(grads[2], grads[3]):             # same seed, hint off vs on
        assert torch.equal(a, b)


@pytest.mark.parametrize('trick,se,L', [('Initial', '000', 2), ('Initial', '111', 3), ('NoRes', '000', 3), ('Residual', '010', 2),
                                         ('InitialJumping', '000', 3)])
def test_layers_that_own_their_dropout_match_separate_dropout_nodes(trick, se, L):
    """Training with dropout > 0 (the reference's defaults, base_options.py:190-220): a layer whose output goes straight
    into F.dropout draws that dropout itself and folds its backward into its prologue.  Same torch call in the same
    order => same masks; logits and every gradient are bit-identical to the path with separate F.dropout nodes."""
    C, G, ops = _pkg()
    from gnn_tail_generalization_b200.GNN_model.GNN_normalizations import TeacherGNN
    n, e, d, Cn = 3000, 30000, 64, 16
    ei = O.canonicalize_planetoid(_multigraph(n, e, 55), n).to(DEV)
    kw = dict(type_trick=trick, whetherHasSE=se, num_layers=L, dim_hidden=d, num_feats=32, num_classes=Cn, N_nodes=n,
              dataset='Cora', res_alpha=0.1, dropout=0.5)
    a = O.make_args(**kw)
    a.device = DEV
    torch.manual_seed(1)
    model = TeacherGNN(a, None).to(DEV).train()
    x = torch.randn(n, 32, generator=torch.Generator().manual_seed(2)).to(DEV)
    y = torch.randint(0, Cn, (n,), generator=torch.Generator().manual_seed(3)).to(DEV)
    mask = torch.arange(n, device=DEV) < n // 5
    results = []
    try:
        for fused in (False, True):
            ops.set_dropout_fusion(fused)
            model.zero_grad(set_to_none=True)
            torch.manual_seed(77)
            sink = []
            ops.set_timing_sink(sink)
            res = model.get_3_embs(x, ei, mask)
            loss = F.nll_loss(F.log_softmax(res.emb4classi, 1), y[mask])
            if model.se_reg_all is not None:
                loss = loss + 0.5 * model.se_reg_all
            loss.backward()
            ops.set_timing_sink(None)
            results.append((res.emb4classi_full.detach().clone(), {k: p.grad.clone() for k, p in model.named_parameters()
                                                                   if p.grad is not None}, [s[0] for s in sink]))
    finally:
        ops.set_dropout_fusion(True)
        ops.set_timing_sink(None)
    (lg0, g0, _), (lg1, g1, names) = results
    assert torch.equal(lg0, lg1)
    assert set(g0) == set(g1) and len(g0) > 0
    for k in g0:
        assert torch.equal(g0[k], g1[k]), k
    if trick in ('Initial', 'NoRes'):
        assert 'agg_gather_src_rowsparse' in names        # the layer under the head found its dead rows itself


def test_head_weight_gradient_on_compacted_rows(monkeypatch):
    """Under a loss over the train rows the head's dW = X^T dY has terms from those rows only: the weight-gradient GEMM
    runs on the compacted operands (one host sync per step, large graphs only).  Same gradient up to the reassociation
    of the fp32 sums; every other gradient bit-identical."""
    C, G, ops = _pkg()
    from gnn_tail_generalization_b200.GNN_model.GNN_normalizations import TeacherGNN
    n, e, d, Cn = 6000, 60000, 64, 32
    ei = O.canonicalize_planetoid(_multigraph(n, e, 91), n).to(DEV)
    a = O.make_args(type_trick='Initial', whetherHasSE='000', num_layers=2, dim_hidden=d, num_feats=d, num_classes=Cn,
                    N_nodes=n, dataset='Cora', res_alpha=0.1, dropout=0.0)
    a.device = DEV
    torch.manual_seed(4)
    model = TeacherGNN(a, None).to(DEV).train()
    x = torch.randn(n, d, generator=torch.Generator().manual_seed(5)).to(DEV)
    y = torch.randint(0, Cn, (n,), generator=torch.Generator().manual_seed(6)).to(DEV)
    mask = torch.zeros(n, dtype=torch.bool, device=DEV)
    mask[torch.randperm(n, generator=torch.Generator().manual_seed(7))[:n // 8].to(DEV)] = True     # scattered train rows
    grads = []
    for min_rows in (1 << 30, 1000):
        monkeypatch.setattr(ops, '_COMPACT_DW_MIN_ROWS', min_rows)
        model.zero_grad(set_to_none=True)
        sink = []
        ops.set_timing_sink(sink)
        res = model.get_3_embs(x, ei, mask)
        F.nll_loss(F.log_softmax(res.emb4classi, 1), y[mask]).backward()
        ops.set_timing_sink(None)
        grads.append(({k: p.grad.clone() for k, p in model.named_parameters()}, [s[2] for s in sink if s[0] == 'gemm_tn']))
    (g0, _), (g1, _) = grads
    head = 'model.model.layers_MLP.1.weight'
    assert head in g0
    for k in g0:
        if k == head:
            scale = float(g0[k].abs().max())
            assert 0 < float((g0[k] - g1[k]).abs().max()) <= 2e-6 * scale or torch.equal(g0[k], g1[k])
        else:
            assert torch.equal(g0[k], g1[k]), k


def test_transpose_identity():
    """<A x, y> == <x, A^T y>: the backward gather is the exact transpose of the forward one."""
    C, G, ops = _pkg()
    n, e, d = 20000, 300000, 64
    ei = _multigraph(n, e, 77)
    h = G.GraphHandle(ei.to(DEV), n)
    gen = torch.Generator().manual_seed(1)
    x, y = torch.randn(n, d, generator=gen).to(DEV), torch.randn(n, d, generator=gen).to(DEV)
    lhs = (ops.agg_gather_raw(h, C.CB_BY_DST, x).double() * y.double()).sum()
    rhs = (x.double() * ops.agg_gather_raw(h, C.CB_BY_SRC, y).double()).sum()
    assert abs(float(lhs - rhs)) <= 1e-6 * abs(float(lhs))


def test_row_scale_and_frobenius():
    C, G, ops = _pkg()
    gen = torch.Generator().manual_seed(3)
    for rows, d in ((1, 1), (33, 7), (1000, 256), (4097, 12)):
        x, s = torch.randn(rows, d, generator=gen), torch.rand(rows, generator=gen)
        assert torch.equal(ops.row_scale_raw(x.to(DEV), s.to(DEV)).cpu(), x * s[:, None])
        got = float(ops.frob_norm(x.to(DEV)))
        assert got == pytest.approx(float(torch.norm(x.double())), rel=2e-6)
    e = torch.randn(500, 64, generator=gen)
    a, b = e.clone().to(DEV).requires_grad_(), e.clone().requires_grad_()
    (3 * ops.frob_norm(a)).backward()
    (3 * torch.norm(b)).backward()
    assert torch.allclose(a.grad.cpu(), b.grad, rtol=1e-5, atol=1e-7)
    z = torch.zeros(4, 4, device=DEV, requires_grad=True)
    ops.frob_norm(z).backward()
    assert torch.equal(z.grad, torch.zeros_like(z))


# ------------------------------------------------------------------------------------------------
# whole model against the reference-generated fixtures and the oracle
# ------------------------------------------------------------------------------------------------
def _teacher(a):
    from gnn_tail_generalization_b200.GNN_model.GNN_normalizations import TeacherGNN
    return TeacherGNN(a, None)


def test_kat_toy_through_gcnconv():
    from gnn_tail_generalization_b200.GNN_model.GCN import GCNConv
    _, G, _ = _pkg()
    z = load_golden('kat_toy')
    ei = torch.from_numpy(z['edge_index'])
    g = G.GraphHandle(ei.to(DEV), 3)
    from types import SimpleNamespace
    for se in (False, True):
        conv = GCNConv(3, 3, args=SimpleNamespace(N_nodes=3), whetherHasSE=se).to(DEV)
        with torch.no_grad():
            conv.weight.copy_(torch.eye(3)); conv.bias.zero_()
            if se:
                conv.le.copy_(torch.arange(9.).view(3, 3) / 10)
        rst, reg = conv(g, torch.eye(3, device=DEV))
        assert np.allclose(rst.detach().cpu().numpy(), z[f'rst_se{int(se)}'], rtol=1e-6, atol=1e-7)
        if se:
            assert float(reg) == pytest.approx(float(z['se_reg']), rel=1e-6)
        else:
            assert reg is None


@pytest.mark.parametrize('name', golden_cases())
def test_model_matches_reference_fixture(name):
    z = load_golden(name)
    a = golden_args(z, O.make_args)
    a.device = DEV
    model = _teacher(a)
    load_params(model, z)
    model.to(DEV)
    model.eval() if name.startswith('exact_batchnorm') else model.train()
    x, ei = torch.from_numpy(z['x']).to(DEV), torch.from_numpy(z['edge_index']).to(DEV)
    y, mask = torch.from_numpy(z['y']).to(DEV), torch.from_numpy(z['train_mask']).to(DEV)
    res = model.get_3_embs(x, ei, mask)
    got = res.emb4classi_full.detach().cpu().numpy()
    assert np.abs(got - z['logits']).max() <= LOGIT_TOL
    assert np.allclose(got, z['logits'], rtol=2e-5, atol=2e-6)            # what is actually observed
    if np.isnan(z['se_reg_all']):
        assert model.se_reg_all is None
    else:
        assert float(model.se_reg_all) == pytest.approx(float(z['se_reg_all']), rel=1e-6)
    loss = F.nll_loss(F.log_softmax(res.emb4classi, 1), y[mask])
    if int(z['reg_in_loss']):
        loss = loss + 0.5 * model.se_reg_all
    assert float(loss) == pytest.approx(float(z['loss']), rel=1e-5)
    loss.backward()
    grads = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
    want = {k[5:]: v for k, v in z.items() if k.startswith('grad/')}
    assert set(grads) == set(want)
    for k, v in want.items():
        g = grads[k].cpu().numpy()
        assert np.abs(g - v).max() <= 1e-5 * max(1.0, np.abs(v).max()), k
    xin = x if a.dim_learnable_input == 0 else model.embs
    les = model.model.model.collect_SE(xin, ei)
    assert np.allclose(les.cpu().numpy(), z['les'], rtol=2e-5, atol=2e-6)
    assert model.model.model.get_se_dim(xin, ei) == z['les'].shape[1]


CONFIG_SHAPES = {
    # BASELINE.json configs at their real shapes, shape-matched synthetic data (datasets are not on disk)
    'cfg1_cora_nores_se000': dict(n=2708, und=5278, F=1433, H=64, C=7, trick='NoResNodeNorm', se='000', ds='Cora'),
    'cfg2_pubmed_initial_se111': dict(n=19717, und=44324, F=500, H=256, C=3, trick='InitialBatchNorm', se='111',
                                      ds='Pubmed'),
    'cfg3_arxiv_initial_se100': dict(n=169343, und=1157799, F=128, H=256, C=40, trick='InitialBatchNorm', se='100',
                                     ds='ogbn-arxiv'),
}


@pytest.mark.parametrize('cfg', sorted(CONFIG_SHAPES))
def test_config_shapes_match_oracle(cfg):
    c = CONFIG_SHAPES[cfg]
    torch.manual_seed(3)
    ei = O.powerlaw_graph(c['n'], c['und'], seed=0)
    kw = dict(type_trick=c['trick'], whetherHasSE=c['se'], num_layers=2, dim_hidden=c['H'], num_feats=c['F'],
              num_classes=c['C'], N_nodes=c['n'], dataset=c['ds'], res_alpha=0.1)
    ref = O.OracleTeacherGNN(O.make_args(**kw), None)
    a = O.make_args(**kw)
    a.device = DEV
    model = _teacher(a)
    model.load_state_dict(ref.state_dict(), strict=True)
    model.to(DEV)
    x = torch.randn(c['n'], c['F'], generator=torch.Generator().manual_seed(1))
    y = torch.randint(0, c['C'], (c['n'],), generator=torch.Generator().manual_seed(2))
    mask = torch.zeros(c['n'], dtype=torch.bool)
    mask[: c['n'] // 10] = True
    for m in (ref, model):
        m.train()
    lr = O.teacher_loss(ref, x, ei, y, mask, 0.5)
    lr.backward()
    res = model.get_3_embs(x.to(DEV), ei.to(DEV), mask.to(DEV))
    lg = F.nll_loss(F.log_softmax(res.emb4classi, 1), y.to(DEV)[mask.to(DEV)])
    if model.se_reg_all is not None:
        lg = lg + 0.5 * model.se_reg_all
    lg.backward()
    want = ref.get_3_embs(x, ei, mask).emb4classi_full.detach()
    got = res.emb4classi_full.detach().cpu()
    assert float((got - want).abs().max()) <= LOGIT_TOL
    assert float(lg) == pytest.approx(float(lr), rel=1e-5)
    # Gradients are judged against the SAME oracle run in fp64 (every config here has <= 2e5 rows): the fp32 CPU
    # oracle is itself up to ~1e-3 of the largest entry away from fp64 on the weight gradients (fp32 reductions
    # over up to 1.7e5 rows with heavy cancellation), so it cannot referee a 1e-4 bar.
    ref64 = O.OracleTeacherGNN(O.make_args(**kw), None)
    ref64.load_state_dict(ref.state_dict(), strict=True)
    ref64.double().train()
    l64 = O.teacher_loss(ref64, x.double(), ei, y, mask, 0.5)
    l64.backward()
    assert float(lg) == pytest.approx(float(l64), rel=2e-6)
    rg, rg64 = dict(ref.named_parameters()), dict(ref64.named_parameters())
    worst = {}
    for k, p in model.named_parameters():
        if rg64[k].grad is None:
            assert p.grad is None, k
            continue
        scale = max(1e-6, float(rg64[k].grad.abs().max()))
        ours = float((p.grad.cpu().double() - rg64[k].grad).abs().max()) / scale
        cpu32 = float((rg[k].grad.double() - rg64[k].grad).abs().max()) / scale
        worst[k] = (ours, cpu32)
        # UNPINNED relu gates: a handful of pre-activations within 1e-5 of zero gate differently in this fp32 path and in
        # fp64, and each flip is worth ~1e-3 of the largest entry of a cancelling reduction (see
        # test_config_shapes_gradients_with_pinned_relu_gates, which pins them and holds 1e-4).  Here: 1e-2, as before.
        bar = 1e-2
        assert ours <= max(bar, 2.5 * cpu32) + 1e-7 / scale, (k, ours, cpu32)
    print(cfg, 'max grad error / largest entry (ours vs fp64, fp32 CPU oracle vs fp64):',
          {k.split('model.model.')[-1]: (f'{a:.1e}', f'{b:.1e}') for k, (a, b) in worst.items()})


@pytest.mark.parametrize('cfg', sorted(CONFIG_SHAPES))
def test_config_shapes_gradients_with_pinned_relu_gates(cfg):
    """The gradient ARITHMETIC of the CUDA path at the BASELINE shapes against the fp64 oracle, with the relu gates
    of the fp64 run pinned to the ones the CUDA forward took.

    Why pinned: on ogbn-arxiv shape ~4e6 pre-activations belong to train rows; the 3xTF32 transform is ~1e-5 away
    from fp64, so a few tens of them land on the other side of zero.  One flipped gate moves a bias-gradient entry
    (a sum of ~1.7e4 cancelling terms) by ~1e-3 of the largest entry -- measured unpinned: 1.3e-3 / 1.5e-3 on
    layers_GCN.1.{weight,bias}, 6e-3 on layers_MLP.0.weight, while every GEMM involved is within 1e-6 of fp64
    (scripts/grad_error_probe.py, scripts/gemm_k40_probe.py).  With the gates pinned what is left is the summation
    and GEMM error, and the bar is north_star's 1e-4 of the largest entry."""
    from gnn_tail_generalization_b200 import ops
    c = CONFIG_SHAPES[cfg]
    torch.manual_seed(3)
    ei = O.powerlaw_graph(c['n'], c['und'], seed=0)
    kw = dict(type_trick=c['trick'], whetherHasSE=c['se'], num_layers=2, dim_hidden=c['H'], num_feats=c['F'],
              num_classes=c['C'], N_nodes=c['n'], dataset=c['ds'], res_alpha=0.1)
    ref64 = O.OracleTeacherGNN(O.make_args(**kw), None)
    a = O.make_args(**kw)
    a.device = DEV
    model = _teacher(a)
    model.load_state_dict(ref64.state_dict(), strict=True)
    model.to(DEV).train()
    ref64.double().train()
    x = torch.randn(c['n'], c['F'], generator=torch.Generator().manual_seed(1))
    y = torch.randint(0, c['C'], (c['n'],), generator=torch.Generator().manual_seed(2))
    mask = torch.zeros(c['n'], dtype=torch.bool)
    mask[: c['n'] // 10] = True
    tc = model.model.model
    xg, eg = x.to(DEV), ei.to(DEV)
    # CUDA run with the pre-activations exposed (want_les: relu and the residual mix run as torch ops on the kernels'
    # outputs; transforms, aggregations and their adjoints are the same kernels as on the fused path)
    logits, se_reg, les = tc.forward(xg, eg, want_les=True)
    loss = F.nll_loss(F.log_softmax(logits[mask.to(DEV)], 1), y.to(DEV)[mask.to(DEV)])
    if se_reg is not None:
        loss = loss + 0.5 * se_reg
    loss.backward()
    L, H = 2, c['H']
    gates = []
    if tc.has_residual_MLP:
        with torch.no_grad():
            lin = tc.layers_MLP[0]
            x0, _ = ops.dense(xg, lin.weight, 'nk', bias=lin.bias, relu=True)
        gates.append((x0 > 0).cpu())
    widths = [H] * (L if tc.has_residual_MLP else L - 1)
    off = 0
    for w in widths:                                  # les = cat of every layer's pre-activation (GCN.py:124-125)
        gates.append((les[:, off:off + w] > 0).cpu())
        off += w
    out64, reg64 = ref64.model.model(x.double(), ei, relu_masks=gates)
    l64 = F.nll_loss(F.log_softmax(out64[mask], 1), y[mask])
    if reg64 is not None:
        l64 = l64 + 0.5 * reg64
    l64.backward()
    assert float(loss) == pytest.approx(float(l64), rel=2e-6)
    rg64 = dict(ref64.named_parameters())
    worst = {}
    for k, p in model.named_parameters():
        if rg64[k].grad is None:
            continue
        scale = max(1e-6, float(rg64[k].grad.abs().max()))
        worst[k.split('model.model.')[-1]] = float((p.grad.cpu().double() - rg64[k].grad).abs().max()) / scale
    print(cfg, 'pinned gates, max grad error / largest entry:', {k: f'{v:.1e}' for k, v in worst.items()})
    for k, v in worst.items():
        assert v <= 1e-4, (k, v)


def test_zero_in_degree_raises_dglerror():
    from gnn_tail_generalization_b200 import DGLError
    a = O.make_args(N_nodes=3, num_feats=4, dim_hidden=4, num_classes=2)
    a.device = DEV
    model = _teacher(a).to(DEV)
    ei = torch.tensor([[0, 1], [1, 1]]).to(DEV)
    with pytest.raises(DGLError):
        model(torch.randn(3, 4, device=DEV), ei)
    for conv in model.model.model.layers_GCN:
        conv.set_allow_zero_in_degree(True)
    model.model.model.dglgraph = None
    assert model(torch.randn(3, 4, device=DEV), ei).shape == (3, 2)


def test_training_dropout_path_runs_and_eval_is_deterministic():
    n = 400
    ei = O.powerlaw_graph(n, 1500, seed=4)
    a = O.make_args(N_nodes=n, num_feats=32, dim_hidden=64, num_classes=5, type_trick='InitialBatchNorm',
                    whetherHasSE='111', dropout=0.5)
    a.device = DEV
    model = _teacher(a).to(DEV)
    x = torch.randn(n, 32, device=DEV)
    model.train()
    out = model(x, ei.to(DEV))
    (out.sum() + model.se_reg_all).backward()
    assert all(p.grad is not None for k, p in model.named_parameters() if 'layers_norm' not in k)
    model.eval()
    with torch.no_grad():
        assert torch.equal(model(x, ei.to(DEV)), model(x, ei.to(DEV)))


# ------------------------------------------------------------------------------------------------
# size-independent properties at a size the oracle would not finish quickly
# ------------------------------------------------------------------------------------------------
def test_large_graph_properties():
    C, G, ops = _pkg()
    n, und, d = 1_000_000, 5_000_000, 256
    ei = O.powerlaw_graph(n, und, seed=0).to(DEV)
    h = G.GraphHandle(ei, n)
    assert h.num_edges == 2 * und + n and not h.has_zero_in_degree
    rowptr, col, perm = h.csr(C.CB_BY_DST)
    # sortedness + stability: keys nondecreasing, edge ids increasing inside a row
    dst_sorted = ei[1][perm.long()]
    assert bool((dst_sorted[1:] >= dst_sorted[:-1]).all())
    same = dst_sorted[1:] == dst_sorted[:-1]
    assert bool((perm[1:][same] > perm[:-1][same]).all())
    assert torch.equal(col.long(), ei[0][perm.long()])
    assert torch.equal(rowptr[1:] - rowptr[:-1], torch.bincount(ei[1], minlength=n))
    gen = torch.Generator(device=DEV).manual_seed(0)
    x = torch.randn(n, d, device=DEV, generator=gen)
    y = torch.randn(n, d, device=DEV, generator=gen)
    ax, ay = ops.agg_gather_raw(h, C.CB_BY_DST, x), ops.agg_gather_raw(h, C.CB_BY_DST, y)
    # checksum: column sums of A x equal out-degree-weighted column sums of x
    lhs = ax.double().sum(0)
    rhs = (x.double() * h.out_degrees().double()[:, None]).sum(0)
    assert torch.allclose(lhs, rhs, rtol=1e-5, atol=0.5)      # column sums ~1e5; one missing edge shifts them by ~1
    # linearity
    axy = ops.agg_gather_raw(h, C.CB_BY_DST, x + 2 * y)
    assert torch.allclose(axy, ax + 2 * ay, rtol=1e-4, atol=1e-3)
    # symmetric graph: by-source gather == by-destination gather up to summation order
    assert torch.allclose(ops.agg_gather_raw(h, C.CB_BY_SRC, x), ax, rtol=1e-4, atol=1e-3)
    # against torch's own scatter on the GPU, in fp64 (its fp32 atomics add in a different order every run,
    # which on the hub rows -- up to ~3e4 terms -- is worth more than the tolerance)
    want = torch.zeros(n, d, dtype=torch.float64, device=DEV).index_add_(0, ei[1], x.double()[ei[0]])
    assert torch.allclose(ax.double(), want, rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize('trick,se,layers', [('Initial', '000', 2), ('Initial', '111', 3), ('NoRes', '010', 3),
                                             ('InitialBatchNorm', '100', 2)])
def test_backward_fusion_matches_unfused(trick, se, layers):
    """The backward prologues folded into the dX GEMM epilogues (ops.BwdPlan) change no number except the
    order of the bias-gradient column sums."""
    _, _, ops = _pkg()
    n, F_in, H, Cn = 5000, 96, 128, 12
    ei = O.powerlaw_graph(n, 20000, seed=4)
    kw = dict(type_trick=trick, whetherHasSE=se, num_layers=layers, dim_hidden=H, num_feats=F_in, num_classes=Cn,
              N_nodes=n, dataset='Cora', res_alpha=0.1)
    x = torch.randn(n, F_in, generator=torch.Generator().manual_seed(1)).to(DEV)
    y = torch.randint(0, Cn, (n,), generator=torch.Generator().manual_seed(2)).to(DEV)
    mask = (torch.arange(n) < n // 4).to(DEV)
    grads = []
    launches = []
    for fused in (False, True):
        ops.set_backward_fusion(fused)
        try:
            torch.manual_seed(7)
            a = O.make_args(**kw)
            a.device = DEV
            model = _teacher(a).to(DEV)
            model.train()
            sink = []
            ops.set_timing_sink(sink)
            res = model.get_3_embs(x, ei.to(DEV), mask)
            loss = F.nll_loss(F.log_softmax(res.emb4classi, 1), y[mask])
            if model.se_reg_all is not None:
                loss = loss + 0.5 * model.se_reg_all
            loss.backward()
            torch.cuda.synchronize()
            launches.append([s[0] for s in sink])
            grads.append({k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None})
        finally:
            ops.set_timing_sink(None)
            ops.set_backward_fusion(True)
    assert 'gemm_rows_grad' not in launches[0] and 'backward_prep' in launches[0]
    assert 'gemm_rows_grad' in launches[1]
    assert set(grads[0]) == set(grads[1])
    for k in grads[0]:
        a_, b_ = grads[0][k], grads[1][k]
        if k.endswith('bias'):
            assert float((a_ - b_).abs().max()) <= 1e-5 * float(a_.abs().max()) + 1e-9, k
        else:
            # the weight gradients see bias-free inputs only: bit-identical
            assert torch.equal(a_, b_), k


# ------------------------------------------------------------------------------------------------
# bf16 storage (BASELINE configs[4]): gathered rows widened to fp32, in-order fp32 sums, fp32 epilogue,
# round-to-nearest-even on store -- bit-exact against the same arithmetic on the CPU
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('d', [1, 6, 8, 24, 64, 128, 136, 256, 520])
def test_aggregate_bf16_bit_exact(d):
    C, G, ops = _pkg()
    n, e, hub = 3000, 40000, 48
    ei = _multigraph(n, e, 300 + d)
    xb = torch.randn(n, d, generator=torch.Generator().manual_seed(d)).bfloat16()
    h = G.GraphHandle(ei.to(DEV), n, hub_chunk=hub)
    assert h.num_hub_chunks[0] > 0
    for side, key, val in ((C.CB_BY_DST, ei[1], ei[0]), (C.CB_BY_SRC, ei[0], ei[1])):
        rp, cl, _ = O.build_csr(key.numpy(), val.numpy(), n)
        want = torch.from_numpy(O.aggregate_sum_csr_ordered(xb.float().numpy(), rp, cl, hub_chunk=hub)).bfloat16()
        got = ops.agg_gather_raw(h, side, xb.to(DEV))
        assert got.dtype == torch.bfloat16
        assert torch.equal(got.cpu().view(torch.int16), want.view(torch.int16))


@pytest.mark.parametrize('d', [12, 128, 256])
@pytest.mark.parametrize('relu,mix', [(True, True), (False, False)])
def test_fused_forward_bf16_matches_fp32_chain(d, relu, mix):
    C, G, ops = _pkg()
    n, alpha = 2500, 0.1
    ei = torch.cat([_multigraph(n, 20000, 9), torch.arange(n).repeat(2, 1)], 1)       # every row has an in-edge
    gen = torch.Generator().manual_seed(d)
    hb = torch.randn(n, d, generator=gen).bfloat16()
    x0b = torch.randn(n, d, generator=gen).bfloat16() if mix else None
    bias = torch.randn(d, generator=gen)
    h = G.GraphHandle(ei.to(DEV), n)
    out, out_s, mask = ops.agg_forward_raw(h, hb.to(DEV), bias.to(DEV), x0b.to(DEV) if mix else None, alpha, relu,
                                           want_out=True, want_scaled=True, want_mask=True)
    rp, cl, _ = O.build_csr(ei[1].numpy(), ei[0].numpy(), n)
    acc = torch.from_numpy(O.aggregate_sum_csr_ordered(hb.float().numpy(), rp, cl, hub_chunk=h.hub_chunk))
    z = acc * h.din_inv_sqrt.cpu()[:, None] + bias                                    # separately rounded fp32 ops
    r = z.clamp_min(0) if relu else z
    if mix:
        r = (torch.tensor(1.0 - alpha, dtype=torch.float32) * r) + (torch.tensor(alpha, dtype=torch.float32) * x0b.float())
    assert torch.equal(out.cpu().view(torch.int16), r.bfloat16().view(torch.int16))
    assert torch.equal(out_s.cpu().view(torch.int16), (r * h.dout_inv_sqrt.cpu()[:, None]).bfloat16().view(torch.int16))
    assert torch.equal(mask.cpu().bool(), z > 0)


def test_graphed_train_step_matches_eager():
    """A whole training step captured in a CUDA graph (graphs.GraphedTrainStep) replays to exactly the eager
    result: the kernels never synchronise or allocate behind the caller's back."""
    from gnn_tail_generalization_b200.graphs import GraphedTrainStep
    n, F_in, H, Cn = 2708, 120, 64, 7
    ei = O.powerlaw_graph(n, 5278, seed=0).to(DEV)
    kw = dict(type_trick='Initial', whetherHasSE='111', num_layers=2, dim_hidden=H, num_feats=F_in, num_classes=Cn,
              N_nodes=n, dataset='Cora', res_alpha=0.1)
    x = torch.randn(n, F_in, generator=torch.Generator().manual_seed(1)).to(DEV)
    y = torch.randint(0, Cn, (n,), generator=torch.Generator().manual_seed(2)).to(DEV)
    mask = torch.arange(n // 5).to(DEV)        # index mask: a boolean one makes emb[mask] synchronise (nonzero)

    def loss_fn(res, yy, model):
        return F.nll_loss(F.log_softmax(res.emb4classi, 1), yy[mask]) + 0.5 * model.se_reg_all

    def make():
        torch.manual_seed(11)
        a = O.make_args(**kw)
        a.device = DEV
        m = _teacher(a).to(DEV).train()
        m.model.model.dropout = m.model.model.embedding_dropout = m.model.model.args.dropout = 0.0
        return m, torch.optim.Adam(m.parameters(), lr=1e-2, capturable=True)

    eager, opt_e = make()
    graphed, opt_g = make()
    warm, steps = 2, 5
    step = GraphedTrainStep(graphed, opt_g, loss_fn, x, ei, mask, y, warmup=warm)     # runs the warm-up steps; the capture
    losses_g = [float(step()) for _ in range(steps)]                                   # itself only records
    losses_e = []
    for _ in range(warm + steps):
        opt_e.zero_grad(set_to_none=True)
        le = loss_fn(eager.get_3_embs(x, ei, mask), y, eager)
        le.backward()
        opt_e.step()
        losses_e.append(float(le.detach()))
    assert losses_g == losses_e[warm:]
    for (k, p), (_, q) in zip(eager.named_parameters(), graphed.named_parameters()):
        assert torch.equal(p, q), k


@pytest.mark.parametrize('mode', ['DAD', 'DA', 'AD'])
def test_label_propagation_matches_edge_valued_spmm(mode):
    """the reference's propagation loop (outcome_correlation.py:128-158) with its normalised adjacency
    as an edge-valued sparse matrix on the CPU (fp64) against the unit-weight gather between two row scalings."""
    from gnn_tail_generalization_b200 import label_propagation as LP
    _, G, _ = _pkg()
    n, c, alpha, iters = 4000, 9, 0.8, 50
    ei = O.powerlaw_graph(n, 12000, seed=3)
    ei = ei[:, ei[0] != ei[1]]                                  # to_undirected graph without self loops
    ei = ei[:, (ei[0] != 5) & (ei[1] != 5)]                      # node 5 isolated: the inf -> 0 rule
    labels = torch.randint(0, c, (n, 1), generator=torch.Generator().manual_seed(1))
    idx = torch.randperm(n, generator=torch.Generator().manual_seed(2))[: n // 3]
    deg = torch.bincount(ei[0], minlength=n).double()
    f = {'DAD': (-0.5, -0.5), 'DA': (-1.0, 0.0), 'AD': (0.0, -1.0)}[mode]
    pw = lambda p: torch.where(deg > 0, deg.pow(p), torch.zeros_like(deg)) if p else torch.ones_like(deg)
    vals = pw(f[0])[ei[0]] * pw(f[1])[ei[1]]
    adj = torch.sparse_coo_tensor(ei, vals, (n, n)).coalesce()
    y = torch.zeros(n, c, dtype=torch.float64)
    y[idx] = F.one_hot(labels[idx].reshape(-1), c).double()
    want = y.clone()
    for _ in range(iters):
        want = torch.clamp(alpha * torch.sparse.mm(adj, want) + (1 - alpha) * y, 0, 1)
    g = G.GraphHandle(ei.to(DEV), n)
    from gnn_tail_generalization_b200 import ops
    sink = []
    ops.set_timing_sink(sink)
    got = LP.label_propagation(g, labels.to(DEV), idx.to(DEV), alpha, iters, mode)
    ops.set_timing_sink(None)
    # one kernel per iteration (cb_agg_propagate: source factor on the iterate, destination factor, axpy and clamp in
    # the gather's epilogue), plus at most one row scaling before the loop
    names = [s_[0] for s_ in sink]
    assert names.count('agg_propagate') == iters and len(names) <= iters + 1, names
    assert float((got.cpu().double() - want).abs().max()) <= 2e-5
    # the same loop with separate gather / axpy / clamp kernels (any post_step callable takes that path)
    yg = y.float().to(DEV)
    unfused = LP.general_outcome_correlation(g, yg, alpha, iters, lambda t: torch.clamp(t, 0, 1), True, mode)
    assert float((got - unfused).abs().max()) <= 2e-6
    # residual correlation flavour: coefficient 1 on y, no clamp (outcome_correlation.py:141-144)
    want2 = y.clone()
    for _ in range(5):
        want2 = 0.5 * torch.sparse.mm(adj, want2) + y
    got2 = LP.general_outcome_correlation(g, yg, 0.5, 5, None, False, mode)
    assert float((got2.cpu().double() - want2).abs().max()) <= 2e-5


@pytest.mark.parametrize('mode', ['DAD', 'DA', 'AD'])
def test_label_propagation_matches_the_reference_code(mode):
    """SURVEY 8f-3 against the REFERENCE'S OWN functions (Label_propagation_model/outcome_correlation.py: process_adj,
    gen_normalized_adjs, label_propagation -> general_outcome_correlation), imported from its checkout and run
    unmodified; torch_sparse.SparseTensor and torch_geometric.utils.to_undirected come from shims/ (COO + index_add_).
    Skipped where no checkout exists (/root/reference, or baseline/_ref/reference staged by scripts/stage_reference.sh)."""
    import sys
    from types import SimpleNamespace
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref = next((c for c in ('/root/reference', os.path.join(root, 'baseline', '_ref', 'reference'))
                if os.path.exists(os.path.join(c, 'Label_propagation_model', 'outcome_correlation.py'))), None)
    if ref is None:
        pytest.skip('no reference checkout')
    for p_ in (ref, os.path.join(root, 'shims')):
        if p_ not in sys.path:
            sys.path.append(p_)
    import importlib
    OC = importlib.import_module('Label_propagation_model.outcome_correlation')
    from gnn_tail_generalization_b200 import label_propagation as LP
    _, G, _ = _pkg()
    n, c, alpha, iters = 6000, 12, 0.5, 50                      # trainer_node_classification.py:36-37: alpha 0.5, 50 steps
    ei = O.powerlaw_graph(n, 20000, seed=6)
    ei = ei[:, ei[0] < ei[1]]                                    # one direction only: process_adj symmetrises it
    ei = ei[:, (ei[0] != 9) & (ei[1] != 9)]                      # node 9 isolated: the inf -> 0 rule (:47-48)
    labels = torch.randint(0, c, (n, 1), generator=torch.Generator().manual_seed(1)).to(DEV)
    idx = torch.randperm(n, generator=torch.Generator().manual_seed(2))[: n // 4].to(DEV)
    data = SimpleNamespace(num_nodes=n, edge_index=ei.to(DEV), y=labels)
    adj, d_isqrt = OC.process_adj(data)                          # mutates data.edge_index (to_undirected)
    A = dict(zip(('DAD', 'DA', 'AD'), OC.gen_normalized_adjs(adj, d_isqrt)))[mode]
    want = OC.label_propagation(data, {'train': idx}, A=A, alpha=alpha, num_propagations=iters, idxs=['train'])
    g = G.GraphHandle(data.edge_index, n)
    got = LP.label_propagation(g, labels, idx, alpha, iters, mode)
    assert got.shape == want.shape
    assert float((got - want).abs().max()) <= 2e-5
