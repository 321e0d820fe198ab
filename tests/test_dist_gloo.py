"""world_size=2 coverage of the node-sliced path's host logic on CPU (gloo): slice bounds, the per-aggregation
row exchange, the dense-gradient all-reduce and the sharded Frobenius norm, against the single-process oracle."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import coldbrew_oracle as O
from gnn_tail_generalization_b200 import dist as cbdist


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, d, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        ei = O.powerlaw_graph(n, 4 * n, seed=0)
        H = torch.randn(n, d, generator=torch.Generator().manual_seed(1))
        lo, hi = cbdist.slice_bounds(n, world, rank)
        # forward: exchange the row blocks, aggregate the owned destination rows only
        full = cbdist.exchange_rows(H[lo:hi].clone(), n, world)
        assert torch.equal(full, H)
        src, dst = ei
        keep = (dst >= lo) & (dst < hi)
        mine = torch.zeros(hi - lo, d).index_add_(0, dst[keep] - lo, full[src[keep]])
        want = O.aggregate_sum(H, ei, n)[lo:hi]
        assert torch.allclose(mine, want, rtol=1e-6, atol=1e-6)
        # backward: owned source rows gather from the exchanged gradient rows
        keep_s = (src >= lo) & (src < hi)
        back = torch.zeros(hi - lo, d).index_add_(0, src[keep_s] - lo, full[dst[keep_s]])
        want_b = torch.zeros(n, d).index_add_(0, src, H[dst])[lo:hi]
        assert torch.allclose(back, want_b, rtol=1e-6, atol=1e-6)
        # dense gradients are summed over ranks, row-sharded parameters are left alone
        lin = torch.nn.Linear(4, 3)
        lin.le = torch.nn.Parameter(torch.zeros(2, 2))
        holder = torch.nn.Module()
        holder.layer = lin                                  # parameter names: layer.weight, layer.bias, layer.le
        for p in lin.parameters():
            p.grad = torch.full_like(p, float(rank + 1))
        n_red = cbdist.allreduce_dense_grads(holder, world)
        assert n_red == 4 * 3 + 3
        assert torch.equal(lin.weight.grad, torch.full_like(lin.weight, 3.0))
        assert torch.equal(lin.le.grad, torch.full_like(lin.le, float(rank + 1)))
        # sharded ||E||_F: sqrt of the all-reduced sum of squares
        ss = cbdist.allreduce_scalar_sum((H[lo:hi] ** 2).sum().reshape(1), world)
        assert float(ss.sqrt()) == pytest.approx(float(torch.linalg.vector_norm(H)), rel=1e-5)
        open(os.path.join(out_dir, f'ok{rank}'), 'w').close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n', [1001, 64])
def test_world2_exchange_and_reductions(tmp_path, n):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n, 8, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f'ok{r}') for r in range(world))


def test_slice_bounds_cover_everything():
    for n in (0, 1, 7, 8, 9, 1000):
        for world in (1, 2, 3, 8):
            spans = [cbdist.slice_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert cbdist.is_row_sharded('model.model.layers_GCN.0.le') and not cbdist.is_row_sharded('x.weight')


def _need_worker(rank, world, port, n, out_dir, align=1):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        ei = O.powerlaw_graph(n, 3 * n, seed=1)
        # drop some edges so that not every row is needed everywhere
        ei = ei[:, torch.randperm(ei.shape[1], generator=torch.Generator().manual_seed(0))[: 2 * n]]
        src, dst = ei
        per = cbdist.rows_per_rank(n, world, align)
        lo, hi = cbdist.slice_bounds(n, world, rank, align)
        # the row exchange with the same (possibly block-aligned) slicing: trailing ranks may own fewer or no rows
        H = torch.randn(n, 4, generator=torch.Generator().manual_seed(2))
        assert torch.equal(cbdist.exchange_rows(H[lo:hi].clone(), n, world, per=per), H)
        # forward: this rank gathers the sources of the in-edges of its rows
        needed = torch.zeros(per * world, dtype=torch.uint8)
        needed[src[(dst >= lo) & (dst < hi)]] = 1
        mask, peers = cbdist.need_masks(needed, n, world, rank, per=per)
        assert peers == [r for r in range(world) if r != rank] and mask.shape == (per,)
        # brute force: bit j of mask[m] <=> some edge (lo+m -> v) has v owned by peers[j]
        owner = torch.div(dst, per, rounding_mode='floor')
        for j, r in enumerate(peers):
            want = torch.zeros(per, dtype=torch.bool)
            sel = (owner == r) & (src >= lo) & (src < hi)
            want[src[sel] - lo] = True
            assert torch.equal(((mask >> j) & 1).bool(), want), (rank, r)
        open(os.path.join(out_dir, f'need{rank}'), 'w').close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,n,align', [(2, 501, 1), (3, 1000, 1), (3, 700, 128), (2, 100, 128)])
def test_need_masks_match_brute_force(tmp_path, world, n, align):
    """Host logic of the fused exchange: which local rows each peer gathers (dist.need_masks), also with the
    block-aligned slicing of source-panelled graphs (a trailing rank then owns few or no rows)."""
    mp.spawn(_need_worker, args=(world, _free_port(), n, str(tmp_path), align), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f'need{r}') for r in range(world))


def _shard_worker(rank, world, port, n, und, out_dir):
    from gnn_tail_generalization_b200 import synth
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        ei = synth.powerlaw_graph_sharded(n, und, rank, world, seed=5, device='cpu')
        lo, hi = cbdist.slice_bounds(n, world, rank)
        src, dst = ei
        assert bool(((dst >= lo) & (dst < hi)).all())                      # every in-edge sits at the owner of its dst
        key = dst * n + src
        assert bool((key[1:] > key[:-1]).all())                            # sorted by (dst, src), no duplicates
        torch.save(ei, os.path.join(out_dir, f'shard{rank}.pt'))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,n,und', [(2, 3001, 12000), (3, 500, 4000)])
def test_sharded_generator_builds_one_canonical_graph(tmp_path, world, n, und):
    """synth.powerlaw_graph_sharded (configs[4]'s per-shard generation): the union of the shards is symmetric, free of
    duplicates, has exactly one self loop per node and the requested size."""
    mp.spawn(_shard_worker, args=(world, _free_port(), n, und, str(tmp_path)), nprocs=world, join=True)
    ei = torch.cat([torch.load(tmp_path / f'shard{r}.pt') for r in range(world)], 1)
    src, dst = ei
    assert int((src == dst).sum()) == n and torch.equal(torch.sort(src[src == dst]).values, torch.arange(n))
    fwd, bwd = dst * n + src, src * n + dst
    assert torch.unique(fwd).numel() == fwd.numel()
    assert torch.equal(torch.sort(fwd).values, torch.sort(bwd).values)      # symmetric
    assert abs(ei.shape[1] - (2 * und + n)) <= 2 * world
    deg = torch.bincount(dst, minlength=n)
    assert int(deg.max()) > 20 * int(deg.median())                           # power law: hubs exist


def test_slice_bounds_with_block_alignment():
    """Slices of whole 128-row blocks (source-panelled graphs: a row tile of the producing GEMM must lie in one panel):
    every node owned exactly once, every boundary but the last on a block edge, trailing ranks may be empty."""
    from gnn_tail_generalization_b200 import dist as cbdist
    for n, world in ((30011, 2), (30011, 8), (1000, 8), (127, 4), (10_000_000, 8), (0, 3)):
        per = cbdist.rows_per_rank(n, world, cbdist.PANEL_ROWS)
        assert per % cbdist.PANEL_ROWS == 0 and per * world >= n
        spans = [cbdist.slice_bounds(n, world, r, cbdist.PANEL_ROWS) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        for (a, b), (c, d) in zip(spans, spans[1:]):
            assert b == c and a <= b
        for a, b in spans:
            assert a % cbdist.PANEL_ROWS == 0 or a == n
        # the panel of local tile t on a rank that starts at `lo` is ((lo >> 7) + t) % S: the producer launch of panel p
        # starts at tile (p - (lo >> 7)) % S and steps by S (dist.PeerExchange.slot)
        for S in (2, 4):
            for lo, hi in spans:
                tiles = -(-(hi - lo) // cbdist.PANEL_ROWS)
                covered = sorted(t for p in range(S) for t in range((p - (lo >> 7)) % S, tiles, S))
                assert covered == list(range(tiles))
                for p in range(S):
                    for t in range((p - (lo >> 7)) % S, tiles, S):
                        assert ((lo >> 7) + t) % S == p
    # alignment 1 is the plain ceil(N / P) slicing
    assert cbdist.rows_per_rank(10, 4) == 3 and cbdist.slice_bounds(10, 4, 3) == (9, 10)
