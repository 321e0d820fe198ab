"""The bulk-copy staged aggregation (k_agg_bulk, CB_AGG_BULK=R: rows staged through shared memory by cp.async.bulk) is
an experiment behind an environment switch read at first use, so it runs in a child process: in-order sums bit-exact
against the C oracle, fused epilogue identical to the default kernel's (whose outputs the parent computes)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import sys, numpy as np, torch
sys.path.insert(0, sys.argv[1])
from gnn_tail_generalization_b200 import _cabi as C, graph as G, ops
from oracle import coldbrew_oracle as O
from tests.test_gpu_parity import _multigraph
out = {}
for d, dt in ((256, torch.float32), (192, torch.float32), (132, torch.float32), (512, torch.bfloat16)):
    n, e, hub = 6000, 90000, 48
    ei = _multigraph(n, e, 900 + d)
    g = torch.Generator().manual_seed(d)
    x = torch.randn(n, d, generator=g).cuda().to(dt)
    x0 = torch.randn(n, d, generator=g).cuda().to(dt)
    bias = torch.randn(d, generator=g).cuda()
    h = G.GraphHandle(ei.cuda(), n, hub_chunk=hub)
    assert h.num_hub_chunks[0] > 0
    before = C.launch_count()
    got = ops.agg_gather_raw(h, C.CB_BY_DST, x)
    if dt == torch.float32:
        rp, cl, _ = O.build_csr(ei[1].numpy(), ei[0].numpy(), n)
        want = O.aggregate_sum_csr_ordered(x.cpu().numpy(), rp, cl, hub_chunk=hub)
        assert np.array_equal(got.cpu().numpy(), want), d
    fo, fs, fm = ops.agg_forward_raw(h, x, bias, x0, 0.3, True, want_out=True, want_scaled=True, want_mask=True)
    key = f'{d}_{str(dt)[6:]}'
    out[key + '_g'] = got.float().cpu().numpy(); out[key + '_o'] = fo.float().cpu().numpy()
    out[key + '_s'] = fs.float().cpu().numpy(); out[key + '_m'] = fm.cpu().numpy()
np.savez(sys.argv[2], **out)
print('child ok')
'''


@pytest.mark.parametrize('ring', [4, 8])
def test_bulk_staged_aggregation_is_bit_identical(ring, tmp_path):
    files = {}
    for tag, env in (('default', {}), ('bulk', {'CB_AGG_BULK': str(ring)})):
        files[tag] = str(tmp_path / f'{tag}.npz')
        r = subprocess.run([sys.executable, '-c', CHILD, ROOT, files[tag]], cwd=ROOT, capture_output=True, text=True,
                           env={**os.environ, **env}, timeout=600)
        assert r.returncode == 0 and 'child ok' in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    a, b = np.load(files['default']), np.load(files['bulk'])
    assert set(a.files) == set(b.files) and len(a.files) == 16
    for k in a.files:
        assert np.array_equal(a[k], b[k]), k
