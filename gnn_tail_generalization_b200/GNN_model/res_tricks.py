"""Residual tricks of the layer stack (mirror of the reference's GNN_model/res_tricks.py:7-55).

The Initial mix is normally fused into the aggregation epilogue (see GCN.TricksComb); these modules
are the general path and keep the reference's parameter names (``layer_transform``, ``layer_att``)
so that checkpoints load with ``strict=True``.
"""
import torch
from torch import nn


def _blend(last, other, alpha):
    return (1 - alpha) * last + alpha * other


class ResidualConnection(nn.Module):
    """(1-alpha) * X[-1] + alpha * X[-2]  (res_tricks.py:12-14)."""

    def __init__(self, alpha=0.5):
        super().__init__()
        self.alpha = alpha

    def forward(self, Xs: list):
        assert len(Xs) >= 1
        if len(Xs) == 1:
            return Xs[-1]
        return _blend(Xs[-1], Xs[-2], self.alpha)


class InitialConnection(nn.Module):
    """(1-alpha) * X[-1] + alpha * X[0]  (res_tricks.py:21-23)."""

    def __init__(self, alpha=0.5):
        super().__init__()
        self.alpha = alpha

    def forward(self, Xs: list):
        assert len(Xs) >= 1
        if len(Xs) == 1:
            return Xs[-1]
        return _blend(Xs[-1], Xs[0], self.alpha)


class DenseConnection(nn.Module):
    """Concat+Linear, element-wise max, or sigmoid-attention over all earlier layers (res_tricks.py:26-55)."""

    def __init__(self, in_dim, out_dim, aggregation='concat'):
        super().__init__()
        self.in_dim, self.out_dim, self.aggregation = in_dim, out_dim, aggregation
        if aggregation == 'concat':
            self.layer_transform = nn.Linear(in_dim, out_dim, bias=True)
        elif aggregation == 'attention':
            self.layer_att = nn.Linear(in_dim, 1, bias=True)

    def forward(self, Xs: list):
        assert len(Xs) >= 1
        kind = self.aggregation
        if kind == 'concat':
            return self.layer_transform(torch.cat(Xs, dim=-1))
        if kind == 'maxpool':
            return torch.stack(Xs, dim=-1).max(dim=-1, keepdim=False).values
        if kind == 'attention':
            stacked = torch.stack(Xs, dim=1)                                    # [n, k+1, c]
            gate = torch.sigmoid(self.layer_att(stacked).squeeze()).unsqueeze(1)  # [n, 1, k+1]
            return torch.matmul(gate, stacked).squeeze()
        raise Exception("Unknown aggregation")
