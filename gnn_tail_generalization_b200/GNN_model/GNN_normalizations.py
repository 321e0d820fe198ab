"""TeacherGNN and its wrapper (mirror of the reference's GNN_model/GNN_normalizations.py:9-73).

The upper seam the unchanged trainer calls (trainer_node_classification.py:4,242-247,309,386-398,486):
``TeacherGNN(args, proj2class)``, ``.get_3_embs(x, edge_index, mask, want_heads)``, ``.se_reg_all``,
``.model.model.collect_SE / get_se_dim``, ``.graph2commonEmb``, ``state_dict`` keys ``embs`` and
``model.model.*``.
"""
import torch
from torch import nn

from .GCN import TricksComb
from .norm_tricks import *  # noqa: F401,F403  (the reference re-exports these names from here)


class D:
    """Bare attribute bag; the reference imports it from its utils.py (utils.py:857)."""


class GNN_norm(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.model = TricksComb(args)

    def forward(self, x, edge_index):
        return self.model.forward(x, edge_index)


class TeacherGNN(nn.Module):
    """Teacher GCN with structural embeddings (Cold Brew)."""

    def __init__(self, args, proj2class=None):
        super().__init__()
        proj2class = proj2class or nn.Identity()
        # the trainer keeps reading these rewritten fields afterwards (GNN_normalizations.py:13-22)
        args.num_classes_bkup = args.num_classes
        args.num_classes = args.dim_commonEmb
        self.args = args
        if self.args.dim_learnable_input > 0:
            self.embs = nn.Parameter(torch.randn(args.N_nodes, args.dim_learnable_input) * 0.001, requires_grad=True)
            self.args.num_feats_bkup = self.args.num_feats
            self.args.num_feats = self.args.dim_learnable_input
        self.model = GNN_norm(args)
        self.proj2linkp = nn.Identity()
        self.proj2class = proj2class
        self.dglgraph = None
        self.se_reg_all = None

    def forward(self, x, edge_index):
        if self.args.TeacherGNN.change_to_featureless:
            x = x * 0
        if self.args.dim_learnable_input > 0:
            x = self.embs
        commonEmb, self.se_reg_all = self.model(x, edge_index)
        self.out = commonEmb
        return commonEmb

    def get_3_embs(self, x, edge_index, mask=None, want_heads=True):
        commonEmb = self.forward(x, edge_index)
        emb4classi_full = self.proj2class(commonEmb)
        emb4linkp = emb4classi = None
        if want_heads:
            emb4classi = emb4classi_full[mask] if mask is not None else emb4classi_full
            emb4linkp = self.proj2linkp(commonEmb)
        res = D()
        res.commonEmb, res.emb4classi, res.emb4classi_full, res.emb4linkp = \
            commonEmb, emb4classi, emb4classi_full, emb4linkp
        return res

    def get_emb4linkp(self, x, edge_index, mask=None):
        # the reference unpacks the namespace as a tuple here (GNN_normalizations.py:57-60), which raises;
        # return the field it was after
        return self.get_3_embs(x, edge_index, want_heads=True).emb4linkp

    def graph2commonEmb(self, x, edge_index, train_mask):
        commonEmb = self.forward(x, edge_index)
        return commonEmb[train_mask], commonEmb
