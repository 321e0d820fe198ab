"""GPU parity of the tcgen05 3xTF32 dense transform (cb_gemm_rows) against an fp64 reference.

The reference GEMM is fp32 th.matmul with TF32 off (GNN_model/GCN.py:225); the tolerance below is the
fp32-class bound the kernel is designed to: |err| <= 2e-6 * sum_k |a_k b_k| + tiny, which is far inside the
1e-4 logit tolerance of BASELINE.json's north_star.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops():
    from gnn_tail_generalization_b200 import ops
    return ops


def _ref(A, W_nk, rs=None, bias=None, add=None, relu=False):
    acc = A.double() @ W_nk.double().t()
    if rs is not None:
        acc = acc * rs.double()[:, None]
    if bias is not None:
        acc = acc + bias.double()
    if add is not None:
        acc = acc + add.double()
    if relu:
        acc = acc.clamp_min(0)
    return acc


def _bound(A, W_nk):
    return (A.abs().double() @ W_nk.abs().double().t())


@pytest.mark.parametrize('M,K,N', [(128, 32, 64), (128, 256, 256), (1000, 256, 256), (300, 64, 64), (4097, 128, 128),
                                   (777, 100, 40), (5000, 256, 64), (2000, 64, 256), (3000, 500, 256),
                                   (1500, 256, 512), (129, 36, 8)])
def test_gemm_rows_matches_fp64(M, K, N):
    ops = _ops()
    g = torch.Generator(device='cuda').manual_seed(M * 7 + K * 3 + N)
    A = torch.randn(M, K, device='cuda', generator=g)
    W = torch.randn(N, K, device='cuda', generator=g) / K ** 0.5
    wt = ops.split_weight(W, transpose=False)
    out = ops.gemm_rows_raw(A, wt)
    ref = _ref(A, W)
    err = (out.double() - ref).abs()
    bound = 2e-6 * _bound(A, W) + 1e-7
    assert bool((err <= bound).all()), f'max err {float(err.max()):.3e}, worst ratio {float((err / bound).max()):.2f}'


def test_gemm_rows_transposed_weight_and_epilogue():
    ops = _ops()
    g = torch.Generator(device='cuda').manual_seed(5)
    M, K, N = 3001, 256, 256
    A = torch.randn(M, K, device='cuda', generator=g)
    W_kn = torch.randn(K, N, device='cuda', generator=g) / 16        # GCNConv weight layout [in, out]
    rs = torch.rand(M, device='cuda', generator=g) + 0.1
    rs2 = torch.rand(M, device='cuda', generator=g) + 0.1
    bias = torch.randn(N, device='cuda', generator=g)
    add = torch.randn(M, N, device='cuda', generator=g)
    wt = ops.split_weight(W_kn, transpose=True)
    out, out2 = ops.gemm_rows_raw(A, wt, row_scale=rs, bias=bias, add=add, relu=True, out2_scale=rs2, want_out2=True)
    ref = _ref(A, W_kn.t(), rs, bias, add, True)
    tol = 2e-6 * (_bound(A, W_kn.t()) * rs.double()[:, None] + bias.abs().double() + add.abs().double()) + 1e-7
    assert bool(((out.double() - ref).abs() <= tol).all())
    assert torch.equal(out2, out * rs2[:, None])
    # adjoint use: dX = dH @ W^T  (transpose=0 on the same [in,out] weight)
    dH = torch.randn(M, N, device='cuda', generator=g)
    wb = ops.split_weight(W_kn, transpose=False)
    dX = ops.gemm_rows_raw(dH, wb)
    refx = dH.double() @ W_kn.double().t()
    assert bool(((dX.double() - refx).abs() <= 2e-6 * (dH.abs().double() @ W_kn.abs().double().t()) + 1e-7).all())


def test_gemm_rows_close_to_fp32_cublas_at_scale():
    """1M x 256 x 256: no worse than cuBLAS fp32 against the fp64 answer, row sampled."""
    ops = _ops()
    g = torch.Generator(device='cuda').manual_seed(11)
    M, K, N = 1_000_000, 256, 256
    A = torch.randn(M, K, device='cuda', generator=g)
    W = torch.randn(N, K, device='cuda', generator=g) / 16
    out = ops.gemm_rows_raw(A, ops.split_weight(W, transpose=False))
    idx = torch.randint(0, M, (4096,), device='cuda', generator=g)
    idx[:3] = torch.tensor([0, M - 1, M - 129], device='cuda')
    ref = A[idx].double() @ W.double().t()
    e_ours = float((out[idx].double() - ref).abs().max())
    e_cublas = float(((A[idx] @ W.t()).double() - ref).abs().max())
    assert e_ours <= max(4 * e_cublas, 2e-5), (e_ours, e_cublas)
    # every tile was written (no stale rows): compare a strided sample against fp32 matmul loosely
    sl = slice(0, M, 997)
    assert torch.allclose(out[sl], A[sl] @ W.t(), atol=1e-4, rtol=1e-4)


@pytest.mark.parametrize('M,Ka,Nb', [(16, 32, 32), (100, 32, 64), (4096, 256, 256), (5000, 128, 256), (7777, 256, 64),
                                     (30001, 64, 32), (200000, 256, 256), (3000, 512, 288)])
def test_gemm_tn_matches_fp64(M, Ka, Nb):
    ops = _ops()
    g = torch.Generator(device='cuda').manual_seed(M + Ka + Nb)
    A = torch.randn(M, Ka, device='cuda', generator=g)
    B = torch.randn(M, Nb, device='cuda', generator=g)
    out = ops.gemm_tn_raw(A, B)
    ref = A.double().t() @ B.double()
    bound = 2e-6 * (A.abs().double().t() @ B.abs().double()) + 1e-7
    err = (out.double() - ref).abs()
    assert bool((err <= bound).all()), f'max err {float(err.max()):.3e}, worst ratio {float((err / bound).max()):.2f}'
    assert torch.equal(out, ops.gemm_tn_raw(A, B)), 'split-K reduction must be bit-stable'


def test_gemm_tn_same_sign_data_keeps_fp32_class_relative_error():
    """All-positive operands: every truncation of the tensor core's accumulator has the same sign, which is
    the worst case for a long accumulation chain.  The segmented accumulation bounds it."""
    ops = _ops()
    g = torch.Generator(device='cuda').manual_seed(3)
    M = 600_000
    A = torch.rand(M, 64, device='cuda', generator=g) + 0.5
    B = torch.rand(M, 96, device='cuda', generator=g) + 0.5
    out = ops.gemm_tn_raw(A, B)
    ref = A.double().t() @ B.double()
    rel = float(((out.double() - ref).abs() / ref.abs()).max())
    rel_cublas = float((((A.t() @ B).double() - ref).abs() / ref.abs()).max())
    assert rel <= 4e-5, (rel, rel_cublas)


# ------------------------------------------------------------------------------------------------
# cb_gemm_rows_grad: the dX GEMM with the backward prologue of the layer below in its epilogue must give
# exactly what the two-kernel path gives (cb_gemm_rows, then cb_agg_backward_prep / torch relu backward)
# ------------------------------------------------------------------------------------------------
def _small_graph(n, seed):
    from gnn_tail_generalization_b200 import graph
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(0, n, (4 * n,), generator=g)
    dst = torch.randint(0, n, (4 * n,), generator=g)
    loops = torch.arange(n)
    ei = torch.stack([torch.cat([src, loops]), torch.cat([dst, loops])])
    return graph.GraphHandle(ei.cuda(), n)


@pytest.mark.parametrize('M,K,N', [(300, 64, 64), (1000, 64, 256), (4097, 128, 128), (3000, 256, 256), (777, 256, 512)])
@pytest.mark.parametrize('gate', ['u8', 'f32', 'none'])
@pytest.mark.parametrize('slot', ['out', 'out_scaled'])
def test_gemm_rows_grad_equals_gemm_then_prep(M, K, N, gate, slot):
    ops = _ops()
    gph = _small_graph(M, M + K)
    g = torch.Generator(device='cuda').manual_seed(M + 3 * K + 5 * N)
    dY = torch.randn(M, K, device='cuda', generator=g)
    W = torch.randn(N, K, device='cuda', generator=g) / K ** 0.5
    act = torch.randn(M, N, device='cuda', generator=g)
    mask = (act > 0).to(torch.uint8)
    relu_out = act.clamp_min(0)
    wt = ops.split_weight(W, transpose=False)
    alpha, mixed = 0.1, True
    relu = gate != 'none'
    for accumulate in (False, True):
        parked = torch.randn(M, N, device='cuda', generator=g) if accumulate else None
        # two kernels
        dx = ops.gemm_rows_raw(dY, wt)
        G_ref, db_ref, dx0_ref = ops.backward_prep_raw(
            gph, dx if slot == 'out' else None, dx if slot == 'out_scaled' else None,
            mask if gate == 'u8' else None, relu_out if gate == 'f32' else None, relu, mixed, alpha, True, True,
            d_x0_accum=parked.clone() if accumulate else None)
        # one kernel
        G, db, dx0 = ops.gemm_rows_grad_raw(
            dY, wt, row_scale=gph.dout_inv_sqrt if slot == 'out_scaled' else None,
            gate_u8=mask if gate == 'u8' else None, gate_f32=relu_out if gate == 'f32' else None, mixed=mixed,
            alpha=alpha, d_x0=parked.clone() if accumulate else None, accumulate_x0=accumulate, want_x0=True,
            post_scale=gph.din_inv_sqrt, want_col_sum=True)
        assert torch.equal(G, G_ref)
        assert torch.equal(dx0, dx0_ref)
        scale = float(db_ref.abs().max()) + 1e-6
        assert float((db - db_ref).abs().max()) <= 1e-5 * scale * max(1.0, (M / 1000) ** 0.5)


def test_gemm_rows_grad_relu_bias_mode():
    ops = _ops()
    g = torch.Generator(device='cuda').manual_seed(11)
    M, K, N = 2500, 256, 256
    dY = torch.randn(M, K, device='cuda', generator=g)
    W = torch.randn(N, K, device='cuda', generator=g) / 16
    rs = torch.rand(M, device='cuda', generator=g) + 0.1
    parked = torch.randn(M, N, device='cuda', generator=g)
    y = torch.randn(M, N, device='cuda', generator=g).clamp_min(0)
    wt = ops.split_weight(W, transpose=False)
    dx = ops.gemm_rows_raw(dY, wt, row_scale=rs, add=parked)
    want = torch.where(y > 0, dx, torch.zeros((), device='cuda'))
    got, col, _ = ops.gemm_rows_grad_raw(dY, wt, row_scale=rs, add=parked, gate_f32=y, want_col_sum=True)
    assert torch.equal(got, want)
    ref = want.double().sum(0)
    assert float((col.double() - ref).abs().max()) <= 1e-5 * float(ref.abs().max())


def test_gemm_rows_grad_row_live_flags_and_sparse_gather():
    """row_live marks exactly the non-zero output rows, and a gather that skips the unmarked rows is bit-identical."""
    from gnn_tail_generalization_b200 import _cabi as C
    ops = _ops()
    M, K, N = 3000, 64, 256
    gph = _small_graph(M, 3)
    g = torch.Generator(device='cuda').manual_seed(21)
    dY = torch.randn(M, K, device='cuda', generator=g)
    dY[M // 5:] = 0                                   # loss over the first rows only
    dY[7] = 0
    W = torch.randn(N, K, device='cuda', generator=g) / 8
    mask = (torch.rand(M, N, device='cuda', generator=g) > 0.3).to(torch.uint8)
    mask[11] = 0                                      # a live input row whose output is gated to zero
    wt = ops.split_weight(W, transpose=False)
    live = torch.zeros(M, dtype=torch.uint8, device='cuda')
    G, _, _ = ops.gemm_rows_grad_raw(dY, wt, gate_u8=mask, mixed=True, alpha=0.1, post_scale=gph.din_inv_sqrt,
                                     row_live=live)
    want = (G != 0).any(1)
    assert torch.equal(live.bool(), want)
    assert 0 < int(want.sum()) < M // 5
    dense = ops.agg_gather_raw(gph, C.CB_BY_SRC, G)
    sparse = ops.agg_gather_raw(gph, C.CB_BY_SRC, G, live=live)                      # compacted live columns
    walked = ops.agg_gather_raw(gph, C.CB_BY_SRC, G, live=live, flag_walk=True)      # full walk, flag per column
    assert torch.equal(dense, sparse) and torch.equal(dense, walked)
    # flags that wrongly kill a live row change the result (the kernel really skips)
    wrong = live.clone()
    wrong[int(want.nonzero()[0])] = 0
    assert not torch.equal(ops.agg_gather_raw(gph, C.CB_BY_SRC, G, live=wrong), dense)
    assert not torch.equal(ops.agg_gather_raw(gph, C.CB_BY_SRC, G, live=wrong, flag_walk=True), dense)


@pytest.mark.parametrize('d,frac,dtype', [(256, 0.1, torch.float32), (64, 0.5, torch.float32), (12, 0.02, torch.float32),
                                          (128, 0.1, torch.bfloat16), (256, 0.0, torch.float32),
                                          (256, 1.0, torch.float32), (128, 0.05, torch.float32),
                                          (512, 0.3, torch.float32), (256, 0.1, torch.bfloat16),
                                          (640, 0.02, torch.float32)])
def test_compacted_gather_bit_identical_incl_hub_rows(d, frac, dtype):
    """cb_graph_compact_live + cb_agg_gather_compacted against the dense gather of a row-sparse matrix, on a graph
    with hub rows (chunked sums keep their association), both CSR sides, and the compacted lists themselves
    against a torch restatement (bit-exact indexing)."""
    from gnn_tail_generalization_b200 import _cabi as C, graph as Gm
    from oracle import coldbrew_oracle as O
    ops = _ops()
    n = 20011
    ei = O.powerlaw_graph(n, 150000, seed=5).to('cuda')
    gph = Gm.GraphHandle(ei, n, hub_chunk=64)
    assert min(gph.num_hub_chunks) > 0
    g = torch.Generator(device='cuda').manual_seed(5)
    live = (torch.rand(n, device='cuda', generator=g) < frac).to(torch.uint8)
    X = (torch.randn(n, d, device='cuda', generator=g) * live[:, None]).to(dtype)
    for side in (C.CB_BY_SRC, C.CB_BY_DST):
        dense = ops.agg_gather_raw(gph, side, X, row_scale=gph.din_inv_sqrt)
        assert torch.equal(dense, ops.agg_gather_raw(gph, side, X, row_scale=gph.din_inv_sqrt, live=live))
        # the compacted structure
        ws = ops.compact_live_raw(gph, side, live)
        rowptr, col, _ = gph.csr(side)
        keep = live[col.long()].bool()
        want_col = col[keep]
        csum = torch.cat([torch.zeros(1, dtype=torch.int64, device='cuda'), keep.long().cumsum(0)])
        want_rowptr = csum[rowptr]
        got_rowptr = ws[:(n + 1) * 8].view(torch.int64)
        assert torch.equal(got_rowptr, want_rowptr)
        al = lambda b: (b + 255) // 256 * 256
        nch = gph.num_hub_chunks[0 if side == C.CB_BY_DST else 1]
        off = al((n + 1) * 8) + 2 * al((nch + 1) * 8)
        got_col = ws[off:off + 4 * int(want_col.numel())].view(torch.int32)
        assert torch.equal(got_col, want_col)


@pytest.mark.parametrize('M,K,N,layout', [(2708, 1433, 64, 'nk'), (2708, 64, 7, 'kn'), (1900, 500, 256, 'nk'),
                                          (1900, 256, 3, 'nk'), (5000, 256, 40, 'nk'), (700, 9, 10, 'kn')])
def test_ragged_widths_stay_on_the_tensor_core_kernels(M, K, N, layout):
    """BASELINE configs[0..2] widths (Cora 1433 -> 64 -> 7, Pubmed 500 / 3, ogbn-arxiv 40): ops.dense pads to whole
    128-byte feature groups instead of handing the GEMM to cuBLAS; forward and all three gradients against fp64."""
    ops = _ops()
    g = torch.Generator(device='cuda').manual_seed(M + K + N)
    x = torch.randn(M, K, device='cuda', generator=g, requires_grad=True)
    W = (torch.randn((K, N) if layout == 'kn' else (N, K), device='cuda', generator=g) / K ** 0.5).requires_grad_(True)
    b = torch.randn(N, device='cuda', generator=g, requires_grad=True)
    rs = torch.rand(M, device='cuda', generator=g) + 0.5
    sink = []
    ops.set_timing_sink(sink)
    out, _ = ops.dense(x, W, layout, bias=b, relu=True, row_scale=rs)
    dy = torch.randn(M, N, device='cuda', generator=g)
    out.backward(dy)
    ops.set_timing_sink(None)
    names = [s[0] for s in sink]
    assert names.count('gemm_rows') == 2 and names.count('gemm_tn') == 1, names      # fwd, dX, dW: all ours
    xd, Wd, bd = x.detach().double().requires_grad_(True), W.detach().double().requires_grad_(True), \
        b.detach().double().requires_grad_(True)
    pre = rs.double()[:, None] * (xd @ (Wd if layout == 'kn' else Wd.t())) + bd
    ref = torch.relu(pre)
    # gradients with the relu gates of the CUDA forward (a pre-activation within rounding distance of 0 may gate
    # differently in fp64; that is not what this test measures)
    (pre * (out.detach() > 0).double()).backward(dy.double())
    # the bound of the tensor-core kernels everywhere in this file: 2e-6 of sum |a||b| (K = 1433 makes a long chain)
    Wkn = (W if layout == 'kn' else W.t()).detach().double().abs()
    bound = 3e-6 * (rs.double()[:, None] * (x.detach().double().abs() @ Wkn) + b.detach().double().abs()) + 1e-6
    assert out.shape == (M, N) and bool(((out.double() - ref).abs() <= bound).all()), float((out.double() - ref).abs().max())
    gate = (out.detach() > 0).double() * dy.double().abs() * rs.double()[:, None]
    bx = 3e-6 * (gate @ Wkn.t()) + 1e-6
    bw = 3e-6 * (x.detach().double().abs().t() @ gate) + 1e-6
    if layout == 'nk':
        bw = bw.t()
    assert x.grad.shape == xd.grad.shape and bool(((x.grad.double() - xd.grad).abs() <= bx).all())
    assert W.grad.shape == Wd.grad.shape and bool(((W.grad.double() - Wd.grad).abs() <= bw).all())
    assert float((b.grad.double() - bd.grad).abs().max()) <= 2e-5 * max(1.0, float(bd.grad.abs().max()))


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_gemm_rows_grad_skips_rows_flagged_dead_and_sink_stays_consistent(dtype):
    """a_live: rows whose A row is all-zero are neither computed on nor stored (out / d_x0 keep what the buffers held);
    x0_valid: the next, accumulating call reads such d_x0 rows as zero.  Everything that IS stored is bit-identical to
    the dense calls."""
    ops = _ops()
    M, K, N, alpha = 3000, 64, 256, 0.1
    g = torch.Generator(device='cuda').manual_seed(5)
    dY = torch.randn(M, K, device='cuda', generator=g).to(dtype)
    dead = torch.rand(M, device='cuda', generator=g) < 0.8
    dY[dead] = 0
    W = torch.randn(N, K, device='cuda', generator=g) / 8
    mask = (torch.rand(M, N, device='cuda', generator=g) > 0.3).to(torch.uint8)
    ps = torch.rand(M, device='cuda', generator=g) + 0.5
    wt = ops.split_weight(W, transpose=False, dtype=dtype)
    flags = ops.row_any_nonzero_raw(dY)
    assert torch.equal(flags.bool(), ~dead)
    # dense reference
    G_d, col_d, x0_d = ops.gemm_rows_grad_raw(dY, wt, gate_u8=mask, mixed=True, alpha=alpha, post_scale=ps, want_x0=True,
                                              want_col_sum=True)
    assert float(G_d[dead].abs().max()) == 0.0 and float(x0_d[dead].abs().max()) == 0.0
    # sparse first writer into a sentinel-filled d_x0
    x0_s = torch.full((M, N), 777.0, device='cuda', dtype=dtype)
    G_s, col_s, _ = ops.gemm_rows_grad_raw(dY, wt, gate_u8=mask, mixed=True, alpha=alpha, post_scale=ps, d_x0=x0_s,
                                           accumulate_x0=False, want_col_sum=True, a_live=flags)
    assert torch.equal(G_s[~dead], G_d[~dead]) and torch.equal(x0_s[~dead], x0_d[~dead])
    assert bool((x0_s[dead] == 777.0).all())                      # never written
    assert torch.equal(col_s, col_d)
    # second, dense, accumulating writer: through x0_valid the unwritten rows count as zero
    dY2 = torch.randn(M, K, device='cuda', generator=g).to(dtype)
    x0_ref = x0_d.clone()
    G2_d, _, _ = ops.gemm_rows_grad_raw(dY2, wt, gate_u8=mask, mixed=True, alpha=alpha, post_scale=ps, d_x0=x0_ref,
                                        accumulate_x0=True)
    G2_s, _, _ = ops.gemm_rows_grad_raw(dY2, wt, gate_u8=mask, mixed=True, alpha=alpha, post_scale=ps, d_x0=x0_s,
                                        accumulate_x0=True, x0_valid=flags)
    assert torch.equal(G2_s, G2_d) and torch.equal(x0_s, x0_ref)
    # the sink materialises unwritten rows for every other consumer
    sink = ops.GradSink()
    sink.buf, sink.valid = torch.full((M, N), float('nan'), device='cuda', dtype=dtype), flags
    sink.buf[~dead] = 1.0
    taken = sink.take()
    assert sink.buf is None and sink.valid is None
    assert bool((taken[dead] == 0).all()) and bool((taken[~dead] == 1).all())


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('d', [1, 3, 16, 40, 64, 128, 300])
def test_row_any_nonzero(d, dtype):
    ops = _ops()
    g = torch.Generator(device='cuda').manual_seed(d)
    M = 5003
    x = torch.zeros(M, d, device='cuda')
    rows = torch.randperm(M, device='cuda', generator=g)[: M // 3]
    cols = torch.randint(0, d, (rows.numel(),), device='cuda', generator=g)
    x[rows, cols] = torch.randn(rows.numel(), device='cuda', generator=g) + 3.0
    x[rows[0], :] = 0
    x[rows[0], d - 1] = -0.0                      # a negative zero is a zero
    x = x.to(dtype)
    flags = ops.row_any_nonzero_raw(x)
    assert torch.equal(flags.bool(), (x != 0).any(1))
