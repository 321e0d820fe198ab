"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): the node-sliced path with the
exchange fused into the GEMM epilogue (NVLink peer stores) against the NCCL all-gather exchange and
against the single-GPU run.  The checks live in tests/multigpu_worker.py (one process per GPU)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_push_exchange_matches_allgather_and_single_gpu():
    n = min(8, torch.cuda.device_count())     # every GPU of the box: world 2, 4 or 8
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={n}',
           '--master-addr', '127.0.0.1', '--master-port', '29533', os.path.join(ROOT, 'tests', 'multigpu_worker.py')]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert 'multigpu parity ok' in r.stdout
