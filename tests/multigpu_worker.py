"""Worker of tests/test_gpu_multi.py (run under torch.distributed.run, one rank per GPU).

Checks, on a node-sliced graph, that the exchange fused into the producing GEMM epilogue (peer pushes)
gives bit-identical results to the NCCL all-gather exchange, and that both match the single-GPU run of
the same model on the whole graph (logits bit-identical: the per-row arithmetic is the same; dense
weight gradients within the all-reduce's reassociation).
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


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    n, und, d, Cn, L = 30011, 120000, 128, 16, 3
    ei = synth.powerlaw_graph(n, und, seed=0, device=dev)
    x_all = synth.features(n, d, 1, dev)
    y_all = synth.labels(n, Cn, 2, dev)
    n_train = n // 5

    def make(rows):
        torch.manual_seed(3)
        a = O.make_args(type_trick='Initial', whetherHasSE='000', num_layers=L, dim_hidden=d, num_feats=d,
                        num_classes=Cn, N_nodes=rows, dataset='Cora', res_alpha=0.1)
        a.device = str(dev)
        return TeacherGNN(a, None).to(dev).train()

    def run(model, graph, x, y, idx, world_):
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

    sg = cbdist.SlicedGraph(ei, n, rank, world)
    lo, hi = sg.row_begin, sg.row_end
    idx = torch.arange(max(0, min(n_train, hi) - lo), device=dev)
    model = make(hi - lo)
    logits_nccl, grads_nccl, names_nccl = run(model, sg, x_all[lo:hi].clone(), y_all[lo:hi], idx, world)
    assert not any(nm.endswith('_push') for nm in names_nccl)
    sg.enable_push(d)
    for rep in range(4):   # several rounds: the two exchange buffers are reused
        # reps 2, 3: the exchange pipelined in 4 column panels against the aggregation, pushing GEMMs on 40 CTAs
        sg.peer.panels, sg.peer.push_ctas = (1, 0) if rep < 2 else (4, int(os.environ.get('CB_TEST_CTAS', '40')))
        logits_push, grads_push, names_push = run(model, sg, x_all[lo:hi].clone(), y_all[lo:hi], idx, world)
        assert 'gemm_rows_push' in names_push and 'gemm_rows_grad_push' in names_push, names_push
        assert names_push.count('agg_forward') == L * (1 if rep < 2 else 4), names_push
        assert torch.equal(logits_push, logits_nccl), f'rank {rank}: push logits differ from the all-gather path'
        for k in grads_nccl:
            if rep >= 2 and k.endswith('bias'):   # column sums over a different number of CTAs
                scale = float(grads_nccl[k].abs().max()) + 1e-12
                assert float((grads_push[k] - grads_nccl[k]).abs().max()) <= 1e-5 * scale, k
            else:
                diff = float((grads_push[k] - grads_nccl[k]).abs().max())
                assert diff == 0.0, (f'rank {rank}: grad {k} differs (rep {rep}): max |diff| {diff:.3e} of '
                                     f'{float(grads_nccl[k].abs().max()):.3e}; all: ' + ', '.join(
                    f'{kk.split("model.model.")[-1]}={float((grads_push[kk] - grads_nccl[kk]).abs().max()):.2e}'
                    for kk in grads_nccl))
    # against the whole graph on one GPU
    whole = G.GraphHandle(ei, n)
    ref = make(n)
    logits_one, grads_one, _ = run(ref, whole, x_all, y_all, torch.arange(n_train, device=dev), 1)
    assert torch.equal(logits_push, logits_one[lo:hi]), f'rank {rank}: sliced logits differ from the single-GPU run'
    for k in grads_one:
        scale = float(grads_one[k].abs().max()) + 1e-12
        assert float((grads_push[k] - grads_one[k]).abs().max()) <= 2e-5 * scale, k
    pushed = sg.exchanged_bytes
    dist.barrier()
    if rank == 0:
        print(f'multigpu parity ok: world {world}, pushed {pushed} bytes on rank 0', flush=True)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
