"""Functional stand-in for the DGL surface GNN_model/GCN.py touches (SURVEY 8b "lower seam"), CPU only: lets the
REFERENCE's own GNN_model run as the comparison arm of tests/dropin_harness.py.  ``update_all(copy_src, sum)`` is a
sequential ``index_add_`` over the COO edge list (multigraph semantics), like tests/golden/make_golden.py.
This repo's GNN_model never imports dgl."""
import contextlib

import torch

from . import base, function, utils  # noqa: F401


class _Graph:
    def __init__(self, pair):
        self.src = torch.as_tensor(pair[0], dtype=torch.long)
        self.dst = torch.as_tensor(pair[1], dtype=torch.long)
        self.n = int(max(self.src.max(), self.dst.max())) + 1
        self.srcdata, self.edata = {}, {}
        self.dstdata = self.srcdata

    def to(self, device):
        self.src, self.dst = self.src.to(device), self.dst.to(device)
        return self

    @contextlib.contextmanager
    def local_scope(self):
        saved = dict(self.srcdata), dict(self.edata)
        try:
            yield
        finally:
            self.srcdata.clear(); self.srcdata.update(saved[0])
            self.edata.clear(); self.edata.update(saved[1])

    def in_degrees(self):
        return torch.bincount(self.dst, minlength=self.n)

    def out_degrees(self):
        return torch.bincount(self.src, minlength=self.n)

    def number_of_edges(self):
        return self.src.numel()

    def update_all(self, msg, red):
        kind, field = msg[0], msg[1]
        m = self.srcdata[field][self.src]
        if kind == 'u_mul_e':
            m = m * self.edata[msg[2]].reshape(-1, *([1] * (m.dim() - 1)))
        out = torch.zeros((self.n,) + tuple(m.shape[1:]), dtype=m.dtype, device=m.device)
        self.dstdata[red[2]] = out.index_add_(0, self.dst, m)


def graph(pair):
    return _Graph(pair)


def _off_path(*a, **k):
    raise RuntimeError('dgl is a shim: only dgl.graph / update_all(copy_src|u_mul_e, sum) exist')


heterograph = to_homogeneous = _off_path
