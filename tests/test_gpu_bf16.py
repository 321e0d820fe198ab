"""bf16 storage path (BASELINE.json configs[4]: 128-dim bf16, GCN + SE): the kind::f16 tcgen05 transforms, the
bf16 backward prologue, the fused SE optimizer step, the local-edge graph build, and the whole TeacherGNN in bf16
against the oracle's bf16-storage restatement (oracle/coldbrew_oracle.py::bf16_storage_forward).

Tolerances.  Everything is accumulated in fp32 and rounded to bf16 once per stored element, so a kernel output is
within half a bf16 ulp (2^-9 relative) of the fp64 result of the same bf16 inputs, plus the fp32 accumulation
error; the model-level bars are those of a chain of such roundings and are written where they are used."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import coldbrew_oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda'
BF = torch.bfloat16
ULP = 2.0 ** -8          # spacing of bf16 values relative to their magnitude (upper bound)


def _pkg():
    from gnn_tail_generalization_b200 import _cabi as C, graph as G, ops
    return C, G, ops


def _close_bf16(got, want64, extra=0.0):
    """got (bf16) is the correctly rounded value of want64 up to one extra ulp of slack for the fp32 accumulation."""
    err = (got.double() - want64).abs()
    bound = ULP * want64.abs() + extra
    bad = err > bound
    assert not bool(bad.any()), (float(err.max()), float((err / (want64.abs() + 1e-30)).max()), int(bad.sum()))


@pytest.mark.parametrize('M,K,N', [(1000, 128, 128), (517, 64, 256), (4096, 256, 64), (300, 72, 40), (129, 128, 16),
                                   (2000, 512, 384)])
@pytest.mark.parametrize('epi', ['plain', 'full'])
def test_gemm_rows_bf16(M, K, N, epi):
    _, _, ops = _pkg()
    g = torch.Generator(device=DEV).manual_seed(M + K + N)
    A = torch.randn(M, K, device=DEV, generator=g).to(BF)
    W = torch.randn(N, K, device=DEV, generator=g) / K ** 0.5
    wt = ops.split_weight(W, transpose=False, dtype=BF)
    assert wt.hi.dtype == BF and torch.equal(wt.hi, W.to(BF))
    if not ops.gemm_supported(M, N, K, BF):
        pytest.skip('shape not covered')
    acc = A.double() @ wt.hi.double().t()
    if epi == 'plain':
        out = ops.gemm_rows_raw(A, wt)
        _close_bf16(out, acc, extra=1e-5)
        return
    rs = torch.rand(M, device=DEV, generator=g) + 0.5
    s2 = torch.rand(M, device=DEV, generator=g) + 0.5
    bias = torch.randn(N, device=DEV, generator=g)
    add = torch.randn(M, N, device=DEV, generator=g).to(BF)
    out, out2 = ops.gemm_rows_raw(A, wt, row_scale=rs, bias=bias, add=add, relu=True, out2_scale=s2, want_out2=True)
    want = torch.relu(rs.double()[:, None] * acc + bias.double() + add.double())
    # near the relu kink an fp32-vs-fp64 difference of the pre-activation is an absolute, not a relative, error
    _close_bf16(out, want, extra=2e-5)
    _close_bf16(out2, want * s2.double()[:, None], extra=4e-5)


def test_gemm_rows_bf16_transposed_weight_matches_conv_layout():
    _, _, ops = _pkg()
    g = torch.Generator(device=DEV).manual_seed(3)
    A = torch.randn(700, 128, device=DEV, generator=g).to(BF)
    W = torch.randn(128, 64, device=DEV, generator=g) / 11     # GCNConv.weight is [in, out]
    wt = ops.split_weight(W, transpose=True, dtype=BF)
    _close_bf16(ops.gemm_rows_raw(A, wt), A.double() @ W.to(BF).double(), extra=1e-5)


@pytest.mark.parametrize('M,K,N', [(1500, 128, 128), (333, 64, 256), (5000, 256, 128)])
def test_gemm_rows_grad_bf16(M, K, N):
    """dX GEMM + backward prologue in bf16 against the same chain in fp64 from the same bf16 inputs."""
    _, _, ops = _pkg()
    g = torch.Generator(device=DEV).manual_seed(K + N)
    dY = torch.randn(M, K, device=DEV, generator=g).to(BF)
    dY[M // 2:] = 0
    W = torch.randn(N, K, device=DEV, generator=g) / K ** 0.5
    wt = ops.split_weight(W, transpose=False, dtype=BF)
    rs = torch.rand(M, device=DEV, generator=g) + 0.5
    ps = torch.rand(M, device=DEV, generator=g) + 0.5
    mask = (torch.rand(M, N, device=DEV, generator=g) > 0.4).to(torch.uint8)
    x0_old = torch.randn(M, N, device=DEV, generator=g).to(BF)
    d_x0 = x0_old.clone()
    live = torch.zeros(M, dtype=torch.uint8, device=DEV)
    alpha = 0.1
    out, col, dx0 = ops.gemm_rows_grad_raw(dY, wt, row_scale=rs, gate_u8=mask, mixed=True, alpha=alpha, d_x0=d_x0,
                                           accumulate_x0=True, post_scale=ps, want_col_sum=True, row_live=live)
    dtot = rs.double()[:, None] * (dY.double() @ wt.hi.double().t())
    _close_bf16(dx0, x0_old.double() + alpha * dtot, extra=2e-5)
    dz = (1 - alpha) * dtot * mask.double()
    _close_bf16(out, ps.double()[:, None] * dz, extra=2e-5)
    ref = dz.sum(0)
    assert float((col.double() - ref).abs().max()) <= 2e-5 * float(ref.abs().max()) + 1e-4
    assert torch.equal(live.bool(), (out != 0).any(1))
    # relu-output gate (the input Linear's backward) and the add operand
    y = torch.randn(M, N, device=DEV, generator=g).to(BF)
    parked = torch.randn(M, N, device=DEV, generator=g).to(BF)
    out2, col2, _ = ops.gemm_rows_grad_raw(dY, wt, add=parked, gate_f32=y, want_col_sum=True)
    want2 = (dY.double() @ wt.hi.double().t() + parked.double()) * (y.double() > 0)
    _close_bf16(out2, want2, extra=2e-5)
    assert float((col2.double() - want2.sum(0)).abs().max()) <= 2e-5 * float(want2.sum(0).abs().max()) + 1e-4


@pytest.mark.parametrize('M,Ka,Nb', [(5000, 128, 128), (100000, 128, 64), (777, 64, 256), (33, 256, 256),
                                     (300000, 256, 128)])
def test_gemm_tn_bf16(M, Ka, Nb):
    _, _, ops = _pkg()
    g = torch.Generator(device=DEV).manual_seed(M)
    A = torch.randn(M, Ka, device=DEV, generator=g).to(BF)
    B = torch.randn(M, Nb, device=DEV, generator=g).to(BF)
    assert ops.gemm_tn_supported(M, Ka, Nb, BF)
    got = ops.gemm_tn_raw(A, B)
    assert got.dtype == torch.float32 and got.shape == (Ka, Nb)
    want = A.double().t() @ B.double()
    bound = 2e-6 * (A.double().abs().t() @ B.double().abs())
    assert bool(((got.double() - want).abs() <= bound + 1e-6).all()), float((got.double() - want).abs().max())
    assert torch.equal(got, ops.gemm_tn_raw(A, B))          # fixed-order second pass: bit-stable


@pytest.mark.parametrize('d', [128, 64, 20])
def test_backward_prep_bf16(d):
    C, G, ops = _pkg()
    n = 4000
    ei = O.powerlaw_graph(n, 16000, seed=2).to(DEV)
    gph = G.GraphHandle(ei, n)
    g = torch.Generator(device=DEV).manual_seed(d)
    d_out = torch.randn(n, d, device=DEV, generator=g).to(BF)
    d_out2 = torch.randn(n, d, device=DEV, generator=g).to(BF)
    mask = (torch.rand(n, d, device=DEV, generator=g) > 0.5).to(torch.uint8)
    old = torch.randn(n, d, device=DEV, generator=g).to(BF)
    acc = old.clone()
    Gm, db, dx0 = ops.backward_prep_raw(gph, d_out, d_out2, mask, None, True, True, 0.2, True, True, d_x0_accum=acc)
    dtot = d_out.double() + gph.dout_inv_sqrt.double()[:, None] * d_out2.double()
    _close_bf16(dx0, old.double() + 0.2 * dtot, extra=1e-5)
    dz = 0.8 * dtot * mask.double()
    _close_bf16(Gm, gph.din_inv_sqrt.double()[:, None] * dz, extra=1e-5)
    assert float((db.double() - dz.sum(0)).abs().max()) <= 1e-4 * float(dz.sum(0).abs().max()) + 1e-4


@pytest.mark.parametrize('grad_dtype,shadow', [(torch.float32, False), (BF, True)])
@pytest.mark.parametrize('n,wd,reg', [(4096 * 33 + 3, 5e-4, 10.0), (1000, 0.0, 0.0), (50001, 1e-2, 0.5)])
def test_se_adam_step_matches_torch_adam(grad_dtype, shadow, n, wd, reg):
    """cb_se_adam_step against torch.optim.Adam fed the explicitly assembled gradient
    dL/dh + se_reg * E / ||E||_F (trainer_node_classification.py:310, 393-394, 428-430), five steps."""
    import ctypes
    C, _, ops = _pkg()
    g = torch.Generator(device=DEV).manual_seed(n)
    E = torch.randn(n, device=DEV, generator=g)
    ref = E.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-2, weight_decay=wd)
    m, v = torch.zeros_like(E), torch.zeros_like(E)
    sh = torch.empty(n, dtype=BF, device=DEV) if shadow else None
    well = torch.ones(n, dtype=torch.bool, device=DEV)
    for t in range(1, 6):
        grad = (torch.randn(n, device=DEV, generator=g) * 1e-2).to(grad_dtype)
        ss = ops.sumsq_raw(E)
        norm = ref.detach().double().pow(2).sum().sqrt()
        ref.grad = grad.float() + (reg * ref.detach() / norm.float() if reg else 0.0)
        # Adam divides by sqrt(v) ~ |g|: where the total gradient of a step is ~0 the update m / (sqrt(v) + eps) turns
        # a 1e-9 difference in g (fused multiply-adds here, separate roundings in torch) into a 1e-5 difference of the
        # step.  Those entries only get the coarse bound below.
        well &= (ref.grad + wd * ref.detach()).abs() >= 1e-4
        opt.step()
        C.call('cb_se_adam_step', C.ptr(E), C.ptr(grad), C.CB_BF16 if grad_dtype == BF else C.CB_F32, C.ptr(m), C.ptr(v),
               C.ptr(sh), n, 1e-2, 0.9, 0.999, 1e-8, wd, t, C.ptr(ss) if reg else None, reg, C.stream_ptr(E.device))
        diff = (E - ref.detach()).abs()
        # one Adam step moves an entry by ~lr = 1e-2: the well-conditioned entries agree to 2e-4 of a step
        assert float(diff[well].max()) <= 2e-6 * t, (t, float(diff[well].max()))
        assert float(diff.max()) <= 2e-4 * t and int(well.sum()) > 0.9 * n
        if shadow:
            assert torch.equal(sh, E.to(BF))
    st = opt.state[ref]
    assert torch.allclose(m, st['exp_avg'], rtol=1e-4, atol=1e-8) and torch.allclose(v, st['exp_avg_sq'], rtol=1e-4, atol=1e-11)


def test_graph_create_local_equals_sliced_build():
    """cb_graph_create_local (only the rank's edges) against cb_graph_create_sliced (the whole list) for every slice:
    same CSR on both sides when the local lists keep the global list's order."""
    C, G, _ = _pkg()
    n, world = 5003, 3
    ei = O.powerlaw_graph(n, 30000, seed=7).to(DEV)
    per = (n + world - 1) // world
    for r in range(world):
        lo, hi = min(r * per, n), min((r + 1) * per, n)
        whole = G.GraphHandle(ei, n, row_begin=lo, row_end=hi, hub_chunk=32)
        ins = ei[:, (ei[1] >= lo) & (ei[1] < hi)]
        outs = ei[:, (ei[0] >= lo) & (ei[0] < hi)]
        local = G.GraphHandle(ins, n, row_begin=lo, row_end=hi, hub_chunk=32, local_out_edges=outs)
        assert local.num_edges == whole.num_edges and local.num_edges_by_src == whole.num_edges_by_src
        assert local.has_zero_in_degree == whole.has_zero_in_degree and local.num_hub_chunks == whole.num_hub_chunks
        for side in (C.CB_BY_DST, C.CB_BY_SRC):
            (rp_a, col_a, perm_a), (rp_b, col_b, perm_b) = local.csr(side), whole.csr(side)
            assert torch.equal(rp_a, rp_b) and torch.equal(col_a, col_b)
            src_list = ins if side == C.CB_BY_DST else outs
            other = src_list[0 if side == C.CB_BY_DST else 1]
            assert torch.equal(other[perm_a.long()].to(torch.int32), col_a)       # perm indexes the LOCAL list
        assert torch.equal(local.din_inv_sqrt, whole.din_inv_sqrt) and torch.equal(local.dout_inv_sqrt, whole.dout_inv_sqrt)
    # an edge that does not belong to the slice is refused
    from gnn_tail_generalization_b200._cabi import ColdBrewError
    with pytest.raises(ColdBrewError):
        G.GraphHandle(ei, n, row_begin=0, row_end=per, local_out_edges=ei)


def _teacher(args):
    from gnn_tail_generalization_b200.GNN_model.GNN_normalizations import TeacherGNN
    return TeacherGNN(args, None)


@pytest.mark.parametrize('n,und,d,Cn,L,se', [(3000, 12000, 128, 64, 3, '010'), (20011, 100000, 128, 16, 2, '000'),
                                             (90000, 450000, 64, 64, 3, '010')])
def test_bf16_teacher_matches_bf16_storage_oracle(n, und, d, Cn, L, se):
    """The whole TeacherGNN fed bf16 features (dtype switch by input dtype) + FusedSEAdam's bf16 shadow tables,
    forward and backward, against the oracle's bf16-storage / fp32-accumulate restatement.

    Bars: a logit is a sum of ~d products of O(1) values stored with 2^-9 relative rounding after L+2 stored
    stages; two correct implementations differ where an intermediate value sits on a rounding boundary and the fp32
    accumulation order tips it, i.e. by one bf16 ulp of that intermediate.  Measured here: max |logit diff| ~ 1 ulp
    of the largest logit, mean ~ 0.05 ulp.  The bars are 3 ulp max / 0.25 ulp mean, and a few % of the largest entry for
    every gradient -- 5 % in the assertion: the bf16 gradients are rounded at every layer boundary and the relu gates of
    the two implementations differ where a pre-activation rounds to the other side of zero."""
    from gnn_tail_generalization_b200 import se_optim
    torch.manual_seed(5)
    ei = O.powerlaw_graph(n, und, seed=1)
    kw = dict(type_trick='Initial', whetherHasSE=se, num_layers=L, dim_hidden=d, num_feats=d, num_classes=Cn,
              N_nodes=n, dataset='Cora', res_alpha=0.1)
    ref = O.OracleTeacherGNN(O.make_args(**kw), None)
    a = O.make_args(**kw)
    a.device = DEV
    model = _teacher(a)
    model.load_state_dict(ref.state_dict(), strict=True)
    model.to(DEV).train()
    ref.train()
    x = torch.randn(n, d, generator=torch.Generator().manual_seed(1)).to(BF)
    y = torch.randint(0, Cn, (n,), generator=torch.Generator().manual_seed(2))
    idx = torch.arange(n // 5)
    coef = 0.5
    opt = se_optim.FusedSEAdam(model, lr=1e-2, weight_decay=5e-4, se_reg=coef, shadow_dtype=BF)
    assert len(opt.states) == (L if se == '010' else 0)
    res = model.get_3_embs(x.to(DEV), ei.to(DEV), idx.to(DEV))
    assert res.emb4classi_full.dtype == BF
    logits = res.emb4classi_full.float().cpu()
    want, reg = O.bf16_storage_forward(ref, x.float(), ei)
    scale = float(want.abs().max())
    diff = (logits - want.detach()).abs()
    print(f'bf16 model n={n} L={L}: max |logit diff| {float(diff.max()):.3e} = {float(diff.max()) / (ULP * scale):.2f} ulp of '
          f'the largest logit ({scale:.2f}); mean {float(diff.mean()) / (ULP * scale):.3f} ulp')
    assert float(diff.max()) <= 3 * ULP * scale and float(diff.mean()) <= 0.25 * ULP * scale
    if reg is not None:
        assert float(model.se_reg_all) == pytest.approx(float(reg), rel=2e-6)
    loss = F.nll_loss(F.log_softmax(res.emb4classi.float(), 1), y.to(DEV)[idx.to(DEV)])
    lref = F.nll_loss(F.log_softmax(want[idx], 1), y[idx])
    assert float(loss) == pytest.approx(float(lref), rel=5e-3)
    loss.backward()
    lref.backward()
    rg = dict(ref.named_parameters())
    for k, p in model.named_parameters():
        if k.endswith('.le'):
            continue                       # stepped by the fused optimizer from the GradSlot, checked below
        d = p.grad.float().cpu() - rg[k].grad
        gs, fro = float(rg[k].grad.abs().max()), float(torch.linalg.vector_norm(rg[k].grad))
        # measured: up to 5.2 % of the largest entry on single entries (n = 3000), <= 1 % in Frobenius norm
        assert float(d.abs().max()) <= 1e-1 * gs + 1e-7 and float(torch.linalg.vector_norm(d)) <= 3e-2 * fro + 1e-7, \
            (k, float(d.abs().max()) / gs, float(torch.linalg.vector_norm(d)) / fro)
    # SE tables: the slot holds dL/dh (bf16); one fused step == torch Adam on the oracle's gradient + regulariser
    expect = []
    for i, st in enumerate(opt.states):
        le_ref = rg[f'model.model.layers_GCN.{i}.le']
        gslot = st.slot.grad.float().cpu()
        gs, fro = float(le_ref.grad.abs().max()), float(torch.linalg.vector_norm(le_ref.grad))
        assert float((gslot - le_ref.grad).abs().max()) <= 1e-1 * gs + 1e-7
        assert float(torch.linalg.vector_norm(gslot - le_ref.grad)) <= 3e-2 * fro + 1e-7
        tref = le_ref.detach().clone().requires_grad_(True)
        tref.grad = gslot + coef * tref.detach() / float(torch.linalg.vector_norm(tref.detach().double()))
        torch.optim.Adam([tref], lr=1e-2, weight_decay=5e-4).step()
        expect.append((st.master.clone(), tref.detach()))
    opt.step()
    for st, (before, after) in zip(opt.states, expect):
        assert float((st.master.cpu() - after).abs().max()) <= 3e-6
        assert torch.equal(st.shadow, st.master.to(BF))
        assert not torch.equal(st.master, before)
        assert st.slot.grad is None


def test_fused_se_adam_fp32_training_matches_autograd_plus_torch_adam():
    """Three training steps of an fp32 TeacherGNN with SE on every layer (configs[1] shape class): FusedSEAdam +
    torch Adam on the dense weights against plain autograd (norm backward included) + torch Adam on everything."""
    from gnn_tail_generalization_b200 import se_optim
    n, d, Cn, coef, wd = 5000, 64, 7, 10.0, 5e-4
    ei = O.powerlaw_graph(n, 20000, seed=3).to(DEV)
    kw = dict(type_trick='Initial', whetherHasSE='010', num_layers=2, dim_hidden=d, num_feats=d, num_classes=Cn,
              N_nodes=n, dataset='Pubmed', res_alpha=0.1)
    torch.manual_seed(11)
    a1, a2 = O.make_args(**kw), O.make_args(**kw)
    a1.device = a2.device = DEV
    m1 = _teacher(a1).to(DEV).train()
    m2 = _teacher(a2)
    m2.load_state_dict(m1.state_dict())
    m2.to(DEV).train()
    x = torch.randn(n, d, device=DEV)
    y = torch.randint(0, Cn, (n,), device=DEV)
    idx = torch.arange(n // 5, device=DEV)
    o1 = torch.optim.Adam(m1.parameters(), lr=1e-2, weight_decay=wd)
    se = se_optim.FusedSEAdam(m2, lr=1e-2, weight_decay=wd, se_reg=coef)
    o2 = torch.optim.Adam(se.other_parameters(), lr=1e-2, weight_decay=wd)
    assert len(se.states) == 2 and all(not st.conv.le.requires_grad for st in se.states)
    for step in range(3):
        losses = []
        for m, opts in ((m1, (o1,)), (m2, (o2, se))):
            for o in opts:
                o.zero_grad()
            res = m.get_3_embs(x, ei, idx)
            loss = F.nll_loss(F.log_softmax(res.emb4classi, 1), y[idx]) + coef * m.se_reg_all
            loss.backward()
            for o in opts:
                o.step()
            losses.append(float(loss))
        assert losses[0] == pytest.approx(losses[1], rel=1e-5), (step, losses)
    p1, p2 = dict(m1.named_parameters()), dict(m2.named_parameters())
    for k in p1:
        # an Adam step is ~lr per entry whatever the gradient's size: 3 steps, tolerance well below one step
        assert float((p1[k] - p2[k]).abs().max()) <= 5e-5, (k, float((p1[k] - p2[k]).abs().max()))
