"""Source-panel passes (cb_graph_create_panelled, cb_agg_*_pass, cb_peer_push_t.tile_first / tile_step) on ONE GPU.

What the multi-GPU exchange relies on, checked without peers: the grouped neighbour lists are the plain lists stably
partitioned by source panel; S passes through the carry buffer leave bit-identical outputs to the one-pass kernels on
the same graph (hub rows, both sides, fp32 and bf16 storage); the in-order C oracle on the grouped lists gives the same
bits; a producing GEMM launched once per panel on that panel's row tiles writes exactly what one launch writes.
"""
import ctypes

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


def _panel_of(col, S):
    return (col >> 7) % S


@pytest.mark.parametrize('S', [2, 4])
@pytest.mark.parametrize('n,e,lo,hi', [(3000, 40000, 0, 3000), (5000, 60000, 1280, 3840), (700, 0, 0, 700),
                                       (2000, 30000, 512, 512)])
def test_panelled_lists_are_the_plain_lists_grouped_by_panel(S, n, e, lo, hi):
    C, G, _ = _pkg()
    ei = _multigraph(n, e, 7 + n) if e else torch.zeros(2, 0, dtype=torch.int64)
    plain = G.GraphHandle(ei.to(DEV), n, row_begin=lo, row_end=hi, hub_chunk=32)
    pan = G.GraphHandle(ei.to(DEV), n, row_begin=lo, row_end=hi, hub_chunk=32, src_panels=S)
    assert pan.src_panels == S and pan.num_edges == plain.num_edges and pan.num_edges_by_src == plain.num_edges_by_src
    assert pan.has_zero_in_degree == plain.has_zero_in_degree
    assert torch.equal(pan.din_inv_sqrt, plain.din_inv_sqrt) and torch.equal(pan.dout_inv_sqrt, plain.dout_inv_sqrt)
    for side in (C.CB_BY_DST, C.CB_BY_SRC):
        rp, cl, pm = (t.cpu().numpy() for t in plain.csr(side))
        rq, cq, pq = (t.cpu().numpy() for t in pan.csr(side))
        rexp = pan.rowptr_exp(side).cpu().numpy()
        assert np.array_equal(rp, rq)
        assert np.array_equal(rexp[::S], rq)                  # group offsets refine the row offsets
        assert np.all(np.diff(rexp) >= 0)
        # expected: a stable partition of every row by panel of the column id
        rows = np.repeat(np.arange(hi - lo), np.diff(rp))
        order = np.lexsort((np.arange(len(cl)), _panel_of(cl.astype(np.int64), S), rows))
        assert np.array_equal(cq, cl[order])
        assert np.array_equal(pq, pm[order])
        grp = rows * S + _panel_of(cq.astype(np.int64), S)
        assert np.array_equal(np.bincount(grp, minlength=(hi - lo) * S), np.diff(rexp))


@pytest.mark.parametrize('S', [2, 4])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('d', [8, 64, 256])
def test_passes_equal_one_pass_and_the_in_order_oracle(S, dtype, d):
    C, G, ops = _pkg()
    n, e, hub = 4000, 70000, 48
    ei = _multigraph(n, e, 31 + d)
    h = G.GraphHandle(ei.to(DEV), n, hub_chunk=hub, src_panels=S)
    assert h.num_hub_chunks[0] > 0 and h.num_hub_chunks[1] > 0
    g = torch.Generator().manual_seed(d)
    x = torch.randn(n, d, generator=g).to(DEV).to(dtype)
    x0 = torch.randn(n, d, generator=g).to(DEV).to(dtype)
    bias = torch.randn(d, generator=g).to(DEV)
    for side in (C.CB_BY_DST, C.CB_BY_SRC):
        one = ops.agg_gather_raw(h, side, x, row_scale=h.din_inv_sqrt)
        out = torch.full_like(one, float('nan'))
        for p in range(S):
            ops.agg_gather_raw(h, side, x, row_scale=h.din_inv_sqrt, out=out, src_pass=p)
        assert torch.equal(out, one)
        if dtype == torch.float32:      # the grouped order, summed in order on the CPU
            rp, cl, _ = (t.cpu().numpy() for t in h.csr(side))
            want = O.aggregate_sum_csr_ordered(x.cpu().numpy(), rp, cl.astype(np.int64), hub_chunk=hub)
            assert np.array_equal(ops.agg_gather_raw(h, side, x).cpu().numpy(), want)
    # the fused forward: epilogue (bias, relu, Initial mix, scaled copy, mask) only in the last pass
    one = ops.agg_forward_raw(h, x, bias, x0, 0.3, True, want_out=True, want_scaled=True, want_mask=True)
    outs = (torch.full_like(one[0], float('nan')), torch.full_like(one[1], float('nan')), torch.full_like(one[2], 77))
    for p in range(S):
        ops.agg_forward_raw(h, x, bias, x0, 0.3, True, outs=outs, src_pass=p)
    for a, b in zip(outs, one):
        assert torch.equal(a, b)


def test_pass_calls_reject_a_plain_graph():
    C, G, ops = _pkg()
    ei = _multigraph(500, 4000, 3)
    h = G.GraphHandle(ei.to(DEV), 500)
    x = torch.randn(500, 16, device=DEV)
    with pytest.raises(C.ColdBrewError):
        ops.agg_gather_raw(h, C.CB_BY_DST, x, out=torch.empty_like(x), src_pass=0)


class _LocalSlot:
    """A PushSlot without peers: the producer writes its tile subsets into ``local`` only."""

    def __init__(self, C, rows, width, dtype, S, row_begin):
        self.src_passes, self.n_panels, self.width, self.local_rows, self.dtype = S, 1, width, rows, dtype
        self.panel_width, self.pushed_rows = width, 0
        self.local = torch.full((rows, width), float('nan'), dtype=dtype, device=DEV)
        self.panel_local = [self.local]
        self.descs, self.calls = [], []
        for p in range(S):
            desc = C.PeerPush()
            desc.n_peers, desc.max_ctas, desc.need, desc.row0, desc.ld = 0, 0, None, row_begin, width
            desc.tile_first, desc.tile_step = (p - (row_begin >> 7)) % S, S
            self.descs.append(desc)

    @property
    def n_launches(self):
        return self.src_passes

    def pushed(self, p):
        self.calls.append(p)


@pytest.mark.parametrize('S', [2, 4])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('M,row_begin', [(1000, 0), (777, 384), (100, 128), (128 * 5, 256)])
def test_producing_gemms_on_tile_subsets_write_what_one_launch_writes(S, dtype, M, row_begin):
    C, G, ops = _pkg()
    K, N = 128, 128
    g = torch.Generator().manual_seed(M + S)
    A = torch.randn(M, K, generator=g).to(DEV).to(dtype)
    W = torch.randn(K, N, generator=g).to(DEV) * 0.1
    rs = torch.rand(M, generator=g).to(DEV) + 0.5
    add = torch.randn(M, N, generator=g).to(DEV).to(dtype)
    bias = torch.randn(N, generator=g).to(DEV)
    wt = ops.split_weight(W, transpose=True, dtype=dtype)
    one, one2 = ops.gemm_rows_raw(A, wt, rs, bias, add, True, out2_scale=rs, want_out2=True)
    slot = _LocalSlot(C, M, N, dtype, S, row_begin)
    got, got2 = ops.gemm_rows_raw(A, wt, rs, bias, add, True, out2_scale=rs, want_out2=True, push=slot)
    assert slot.calls == list(range(S))
    assert torch.equal(got, one) and torch.equal(got2, one2)
    # the adjoint GEMM with its epilogue: gate, x0 gradient, column sums added over the passes
    gate = (torch.rand(M, N, generator=g) > 0.4).to(torch.uint8).to(DEV)
    ref, col, dx0 = ops.gemm_rows_grad_raw(A, wt, row_scale=rs, gate_u8=gate, mixed=True, alpha=0.25, want_x0=True,
                                           post_scale=rs, want_col_sum=True)
    slot = _LocalSlot(C, M, N, dtype, S, row_begin)
    out, col_p, dx0_p = ops.gemm_rows_grad_raw(A, wt, row_scale=rs, gate_u8=gate, mixed=True, alpha=0.25, want_x0=True,
                                               post_scale=rs, want_col_sum=True, push=slot)
    assert torch.equal(out, ref) and torch.equal(dx0_p, dx0)
    assert torch.allclose(col_p, col, rtol=1e-5, atol=1e-5 * float(col.abs().max()))
