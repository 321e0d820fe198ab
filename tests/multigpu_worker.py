"""Worker of tests/test_gpu_multi.py (run under torch.distributed.run, one rank per GPU; any world size 2..8).

Checks, on a node-sliced graph, that the exchange fused into the producing GEMM epilogue (peer pushes)
gives bit-identical results to the NCCL all-gather exchange, and that both match the single-GPU run of
the same model on the whole graph (logits bit-identical: the per-row arithmetic is the same; dense
weight gradients within the all-reduce's reassociation).

The train rows change from repetition to repetition, so that a row-liveness flag array left over from the
previous repetition (or read before its all-gather has landed) shows up as a wrong gradient; the last case is
a graph with fewer nodes than ranks, where the last rank owns no rows and must still enter every barrier.
"""
import os
import sys

import torch
import torch.distributed as dist
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gnn_tail_generalization_b200 import dist as cbdist, graph as G, ops, synth  # noqa: E402
from gnn_tail_generalization_b200.GNN_model.GNN_normalizations import TeacherGNN  # noqa: E402
from oracle import coldbrew_oracle as O  # noqa: E402  (test infrastructure: args helper only)


def make_model(rows, d, Cn, L, dev, trick='Initial'):
    torch.manual_seed(3)
    a = O.make_args(type_trick=trick, whetherHasSE='000', num_layers=L, dim_hidden=d, num_feats=d,
                    num_classes=Cn, N_nodes=rows, dataset='Cora', res_alpha=0.1)
    a.device = str(dev)
    return TeacherGNN(a, None).to(dev).train()


def run(model, graph, x, y, idx, n_train, world_):
    cbdist.attach_graph(model, graph)
    model.zero_grad(set_to_none=True)
    sink = []
    ops.set_timing_sink(sink)
    res = model.get_3_embs(x, None, idx)
    loss = F.nll_loss(F.log_softmax(res.emb4classi, 1), y[idx], reduction='sum') / n_train
    loss.backward()
    ops.set_timing_sink(None)
    cbdist.allreduce_dense_grads(model, world_)
    torch.cuda.synchronize()
    return (res.emb4classi_full.detach().clone(), {k: p.grad.clone() for k, p in model.named_parameters()},
            [s[0] for s in sink])


def local_idx(train, lo, hi):
    return train[(train >= lo) & (train < hi)] - lo


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    n, und, d, Cn, L = 30011, 120000, 128, 16, 3
    ei = synth.powerlaw_graph(n, und, seed=0, device=dev)
    x_all = synth.features(n, d, 1, dev)
    y_all = synth.labels(n, Cn, 2, dev)
    # two different train sets (they straddle rank boundaries at every world size)
    trains = [torch.arange(n // 5, device=dev), torch.arange(n // 3, n // 3 + n // 7, device=dev)]

    sg = cbdist.SlicedGraph(ei, n, rank, world)
    lo, hi = sg.row_begin, sg.row_end
    model = make_model(hi - lo, d, Cn, L, dev)
    x_loc, y_loc = x_all[lo:hi].clone(), y_all[lo:hi]
    base = []
    for t in trains:
        lg, gr, names = run(model, sg, x_loc, y_loc, local_idx(t, lo, hi), t.numel(), world)
        assert not any(nm.endswith('_push') for nm in names)
        base.append((lg, gr))
    sg.enable_push(d)
    ctas = int(os.environ.get('CB_TEST_CTAS', '64' if world >= 8 else '40'))
    # (panels, train set): the two exchange buffers are reused from rep to rep and the flags change every time
    plan = [(1, 0), (1, 1), (4, 0), (4, 1), (4, 0), (1, 1)]
    for rep, (panels, tv) in enumerate(plan):
        sg.peer.panels, sg.peer.push_ctas = panels, (0 if panels == 1 else ctas)
        t = trains[tv]
        logits_nccl, grads_nccl = base[tv]
        logits_push, grads_push, names_push = run(model, sg, x_loc, y_loc, local_idx(t, lo, hi), t.numel(), world)
        assert 'gemm_rows_push' in names_push and 'gemm_rows_grad_push' in names_push, names_push
        assert names_push.count('agg_forward') == L * panels, names_push
        assert torch.equal(logits_push, logits_nccl), f'rank {rank}: push logits differ from the all-gather path'
        for k in grads_nccl:
            if panels > 1 and k.endswith('bias'):   # column sums over a different number of CTAs
                scale = float(grads_nccl[k].abs().max()) + 1e-12
                assert float((grads_push[k] - grads_nccl[k]).abs().max()) <= 1e-5 * scale, k
            else:
                diff = float((grads_push[k] - grads_nccl[k]).abs().max())
                assert diff == 0.0, (f'rank {rank}: grad {k} differs (rep {rep}, panels {panels}, train set {tv}): '
                                     f'max |diff| {diff:.3e} of {float(grads_nccl[k].abs().max()):.3e}; all: ' + ', '.join(
                    f'{kk.split("model.model.")[-1]}={float((grads_push[kk] - grads_nccl[kk]).abs().max()):.2e}'
                    for kk in grads_nccl))
    # against the whole graph on one GPU (last repetition used train set 1)
    whole = G.GraphHandle(ei, n)
    ref = make_model(n, d, Cn, L, dev)
    logits_one, grads_one, _ = run(ref, whole, x_all, y_all, trains[1], trains[1].numel(), 1)
    assert torch.equal(logits_push, logits_one[lo:hi]), f'rank {rank}: sliced logits differ from the single-GPU run'
    for k in grads_one:
        scale = float(grads_one[k].abs().max()) + 1e-12
        assert float((grads_push[k] - grads_one[k]).abs().max()) <= 2e-5 * scale, k
    pushed = sg.exchanged_bytes
    del whole, ref
    dist.barrier()

    # ---- source-panel passes: neighbour lists grouped by source panel, 128-row aligned slices, the producing GEMMs
    # launched per panel and the aggregation run as S full-width passes through the carry buffer.  Bit-identical to
    # the all-gather exchange on the same slices and to ONE GPU on the same (grouped) graph.
    for S in (2, 4):
        sg3 = cbdist.SlicedGraph(ei, n, rank, world, src_panels=S)
        lo3, hi3 = sg3.row_begin, sg3.row_end
        assert lo3 % 128 == 0 and sg3.per % 128 == 0
        m3 = make_model(hi3 - lo3, d, Cn, L, dev)
        x3, y3 = x_all[lo3:hi3].clone(), y_all[lo3:hi3]
        base3 = [run(m3, sg3, x3, y3, local_idx(t, lo3, hi3), t.numel(), world)[:2] for t in trains]
        sg3.enable_push(d, push_ctas=ctas)
        for rep, tv in enumerate((0, 1, 0, 1)):
            t = trains[tv]
            lg_p, gr_p, names_p = run(m3, sg3, x3, y3, local_idx(t, lo3, hi3), t.numel(), world)
            assert names_p.count('agg_forward_pass') == L * S, names_p
            assert names_p.count('gemm_rows_push') == L * S, names_p
            assert 'agg_gather_src_pass' in names_p, names_p
            assert torch.equal(lg_p, base3[tv][0]), f'rank {rank}: S={S} pass logits differ from the all-gather path'
            for k in base3[tv][1]:
                if k.endswith('bias'):              # column sums added per pass
                    scale = float(base3[tv][1][k].abs().max()) + 1e-12
                    assert float((gr_p[k] - base3[tv][1][k]).abs().max()) <= 1e-5 * scale, k
                else:
                    assert torch.equal(gr_p[k], base3[tv][1][k]), f'rank {rank}: S={S} rep {rep} grad {k} differs'
        whole3 = G.GraphHandle(ei, n, src_panels=S)
        ref3 = make_model(n, d, Cn, L, dev)
        lg_one, gr_one, _ = run(ref3, whole3, x_all, y_all, trains[1], trains[1].numel(), 1)
        assert torch.equal(lg_p, lg_one[lo3:hi3]), f'rank {rank}: S={S} sliced logits differ from the single-GPU run'
        for k in gr_one:
            scale = float(gr_one[k].abs().max()) + 1e-12
            assert float((gr_p[k] - gr_one[k]).abs().max()) <= 2e-5 * scale, k
        del whole3, ref3
        dist.barrier()

    # ---- a rank without rows: N = world - 1 nodes => ceil(N/P) = 1 row per rank, the last rank owns none ----
    n2 = world - 1
    ring = torch.arange(n2, device=dev)
    ei2 = torch.cat([torch.stack([ring, (ring + 1) % n2]), torch.stack([(ring + 1) % n2, ring]),
                     torch.stack([ring, ring])], 1)
    x2, y2 = synth.features(n2, d, 5, dev), synth.labels(n2, Cn, 6, dev)
    sg2 = cbdist.SlicedGraph(ei2, n2, rank, world)
    lo2, hi2 = sg2.row_begin, sg2.row_end
    assert (hi2 - lo2 == 0) == (rank == world - 1)
    t2 = torch.arange(n2, device=dev)
    m2 = make_model(hi2 - lo2, d, Cn, 2, dev)
    lg_a, gr_a, _ = run(m2, sg2, x2[lo2:hi2].clone(), y2[lo2:hi2], local_idx(t2, lo2, hi2), n2, world)
    sg2.enable_push(d)
    lg_b, gr_b, names_b = run(m2, sg2, x2[lo2:hi2].clone(), y2[lo2:hi2], local_idx(t2, lo2, hi2), n2, world)
    assert torch.equal(lg_a, lg_b)
    for k in gr_a:
        assert torch.equal(gr_a[k], gr_b[k]), k
    one2 = make_model(n2, d, Cn, 2, dev)
    lg_c, gr_c, _ = run(one2, G.GraphHandle(ei2, n2), x2, y2, t2, n2, 1)
    assert torch.equal(lg_b, lg_c[lo2:hi2])
    for k in gr_c:
        scale = float(gr_c[k].abs().max()) + 1e-12
        assert float((gr_b[k] - gr_c[k]).abs().max()) <= 2e-5 * scale, k
    dist.barrier()
    if rank == 0:
        print(f'multigpu parity ok: world {world}, panels 1 and 4 (push grid {ctas}), source-panel passes 2 and 4, '
              f'changing train rows, zero-row rank; pushed {pushed} bytes on rank 0', flush=True)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
