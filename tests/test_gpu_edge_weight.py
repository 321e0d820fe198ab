"""Edge-weighted aggregation (GCN.py:199-202: fn.u_mul_e('h', '_edge_weight', 'm') + fn.sum) -- off the TeacherGNN path,
part of GCNConv.forward's contract.  Bare kernels bit-exact against the in-order C oracle (product rounded, then added;
hub chunks included; both CSR sides), the layer and its three gradients (input, weight, edge weights) against the
oracle's autograd."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import coldbrew_oracle as O
from tests.test_gpu_parity import _multigraph

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _pkg():
    from gnn_tail_generalization_b200 import _cabi, graph, ops
    return _cabi, graph, ops


@pytest.mark.parametrize('d', [1, 3, 4, 20, 64, 130, 256, 512])
def test_weighted_gather_bit_exact(d):
    C, G, ops = _pkg()
    n, e, hub = 2000, 30000, 48
    ei = _multigraph(n, e, 300 + d)
    gen = torch.Generator().manual_seed(d)
    x = torch.randn(n, d, generator=gen)
    w = torch.randn(e, generator=gen)
    h = G.GraphHandle(ei.to(DEV), n, hub_chunk=hub)
    assert h.num_hub_chunks[0] > 0 and h.num_hub_chunks[1] > 0
    for side, key, val in ((C.CB_BY_DST, ei[1], ei[0]), (C.CB_BY_SRC, ei[0], ei[1])):
        rp, cl, pm = O.build_csr(key.numpy(), val.numpy(), n)
        want = O.aggregate_mul_sum_csr_ordered(x.numpy(), rp, cl, w.numpy()[pm], hub_chunk=hub)
        ws = ops.sort_edge_values_raw(h, side, w.to(DEV))
        assert np.array_equal(ws.cpu().numpy(), w.numpy()[pm])
        got = ops.agg_gather_weighted_raw(h, side, x.to(DEV), ws).cpu().numpy()
        assert np.array_equal(got, want)
    # unit weights reproduce the unweighted kernel bit for bit
    ones = ops.sort_edge_values_raw(h, C.CB_BY_DST, torch.ones(e, device=DEV))
    assert torch.equal(ops.agg_gather_weighted_raw(h, C.CB_BY_DST, x.to(DEV), ones), ops.agg_gather_raw(h, C.CB_BY_DST, x.to(DEV)))


def test_weighted_gather_bf16_storage():
    C, G, ops = _pkg()
    n, e, d = 3000, 40000, 128
    ei = _multigraph(n, e, 9)
    gen = torch.Generator().manual_seed(1)
    x = torch.randn(n, d, generator=gen).bfloat16()
    w = torch.randn(e, generator=gen)
    h = G.GraphHandle(ei.to(DEV), n, hub_chunk=64)
    rp, cl, pm = O.build_csr(ei[1].numpy(), ei[0].numpy(), n)
    want = O.aggregate_mul_sum_csr_ordered(x.float().numpy(), rp, cl, w.numpy()[pm], hub_chunk=64)   # fp32 sums of bf16 rows
    got = ops.agg_gather_weighted_raw(h, C.CB_BY_DST, x.to(DEV), ops.sort_edge_values_raw(h, C.CB_BY_DST, w.to(DEV)))
    assert torch.equal(got.cpu(), torch.from_numpy(want).bfloat16())               # one rounding on store


@pytest.mark.parametrize('d', [5, 64, 256])
def test_edge_dot_matches_fp64(d):
    C, G, ops = _pkg()
    n, e = 1500, 25000
    ei = _multigraph(n, e, 40 + d)
    gen = torch.Generator().manual_seed(d)
    x, y = torch.randn(n, d, generator=gen), torch.randn(n, d, generator=gen)
    h = G.GraphHandle(ei.to(DEV), n, hub_chunk=32)
    got = ops.edge_dot_raw(h, C.CB_BY_DST, x.to(DEV), y.to(DEV), e).cpu().double()
    want = (x.double()[ei[0]] * y.double()[ei[1]]).sum(1)
    bound = (x.double()[ei[0]].abs() * y.double()[ei[1]].abs()).sum(1)
    assert float(((got - want).abs() / (bound + 1e-30)).max()) <= 2e-6


@pytest.mark.parametrize('norm,se', [('both', False), ('both', True), ('right', False), ('none', False), ('left', False)])
def test_gcnconv_with_edge_weight_matches_oracle(norm, se):
    C, G, ops = _pkg()
    from gnn_tail_generalization_b200.GNN_model.GCN import GCNConv
    n, e, fin, fout = 900, 9000, 32, 64
    ei = O.canonicalize_planetoid(_multigraph(n, e, 5), n)
    E = ei.shape[1]
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(n, fin, generator=gen)
    w_e = torch.rand(E, generator=gen) + 0.25
    torch.manual_seed(0)
    conv = GCNConv(fin, fout, norm=norm, args=SimpleNamespace(N_nodes=n), whetherHasSE=se).to(DEV)
    graph = G.GraphHandle(ei.to(DEV), n)
    xg, wg = x.to(DEV).requires_grad_(), w_e.to(DEV).requires_grad_()
    rst, reg = conv(graph, xg, edge_weight=wg)
    probe = torch.randn(n, fout, generator=gen)
    (rst * probe.to(DEV)).sum().backward()

    W = conv.weight.detach().cpu().clone().requires_grad_()
    b = conv.bias.detach().cpu().clone().requires_grad_()
    le = conv.le.detach().cpu().clone().requires_grad_() if se else None
    xc, wc = x.clone().requires_grad_(), w_e.clone().requires_grad_()
    if norm == 'both':
        want, want_reg = O.gcn_conv(xc, ei, n, W, b, le, edge_weight=wc)
    else:   # the other norms of GCN.py:205-213, 242-250, written out
        dout = torch.bincount(ei[0], minlength=n).float().clamp(min=1)
        din = torch.bincount(ei[1], minlength=n).float().clamp(min=1)
        hh = (xc / dout[:, None] if norm == 'left' else xc) @ W
        want = O.aggregate_mul_sum(hh, ei, wc, n)
        want = (want / din[:, None] if norm == 'right' else want) + b
        want_reg = None
    (want * probe).sum().backward()
    assert float((rst.detach().cpu() - want.detach()).abs().max()) <= 1e-4
    assert (reg is None) == (want_reg is None)
    for got, ref, name in ((xg.grad, xc.grad, 'x'), (wg.grad, wc.grad, 'edge_weight'), (conv.weight.grad, W.grad, 'W'),
                           (conv.bias.grad, b.grad, 'bias')) + (((conv.le.grad, le.grad, 'le'),) if se else ()):
        scale = float(ref.abs().max()) + 1e-12
        assert float((got.cpu() - ref).abs().max()) <= 1e-4 * scale, name
    # [E, 1] weights are accepted like [E]; a wrong length is the reference's AssertionError
    rst2, _ = conv(graph, xg.detach(), edge_weight=wg.detach().view(-1, 1))
    assert torch.equal(rst2, rst.detach())
    with pytest.raises(AssertionError):
        conv(graph, xg.detach(), edge_weight=wg.detach()[:-1])
