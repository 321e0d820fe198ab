"""Normalisation layers of the layer stack (mirror of the reference's GNN_model/norm_tricks.py).

Out of the hot path: under the shipped configurations these layers are constructed (their parameters
sit in the state_dict) but never executed, because ``appendNormLayer`` matches ``type_trick`` by
substring while ``run_norm_if_any`` requires an exact name (norm_tricks.py:130-150).  They
are plain PyTorch; class and attribute names follow the reference for checkpoint compatibility.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

_RUNNABLE = ('BatchNorm', 'PairNorm', 'NodeNorm', 'MeanNorm', 'GroupNorm', 'CombNorm')


def AcontainsB(A, listB):
    """True if any string of listB occurs inside A."""
    return any(s in A for s in listB)


class pair_norm(nn.Module):
    def forward(self, x):
        centred = x - x.mean(dim=0)
        return centred / (1e-6 + centred.pow(2).sum(dim=1).mean()).sqrt()


class mean_norm(nn.Module):
    def forward(self, x):
        return x - x.mean(dim=0)


class node_norm(nn.Module):
    def __init__(self, node_norm_type="n", unbiased=False, eps=1e-5, power_root=2, **kwargs):
        super().__init__()
        self.node_norm_type, self.unbiased, self.eps, self.power = node_norm_type, unbiased, eps, 1 / power_root

    def _std(self, x):
        return (torch.var(x, unbiased=self.unbiased, dim=1, keepdim=True) + self.eps).sqrt()

    def forward(self, x):
        kind = self.node_norm_type
        if kind == "n":
            return (x - torch.mean(x, dim=1, keepdim=True)) / self._std(x)
        if kind == "v":
            return x / self._std(x)
        if kind == "m":
            return x - torch.mean(x, dim=1, keepdim=True)
        if kind == "srv":
            return x / torch.sqrt(self._std(x))
        if kind == "pr":
            return x / torch.pow(self._std(x), self.power)
        return x

    def extra_repr(self):
        return f"node_norm_type={self.node_norm_type}"


class group_norm(nn.Module):
    def __init__(self, dim_to_norm=None, dim_hidden=16, num_groups=None, skip_weight=None, **w):
        super().__init__()
        self.num_groups, self.skip_weight = num_groups, skip_weight
        self.dim_hidden = dim_hidden if dim_to_norm is None else dim_to_norm
        self.bn = nn.BatchNorm1d(self.dim_hidden * num_groups, momentum=0.3)
        self.group_func = nn.Linear(self.dim_hidden, num_groups, bias=True)

    def forward(self, x):
        if self.num_groups == 1:
            t = self.bn(x)
        else:
            score = F.softmax(self.group_func(x), dim=1)
            t = torch.cat([score[:, g].unsqueeze(dim=1) * x for g in range(self.num_groups)], dim=1)
            t = self.bn(t).view(-1, self.num_groups, self.dim_hidden).sum(dim=1)
        return x + t * self.skip_weight


class comb_norm(nn.Module):
    def __init__(self, norm_list):
        super().__init__()
        self.norm_list = nn.ModuleList(norm_list)

    def forward(self, x):
        for mod in self.norm_list:
            x = mod(x)
        return x


# skip_weight by (dataset family, model family, depth), norm_tricks.py:153-206
def _skip_weight(dataset, type_model, L):
    attn_or_gcn = type_model in ('GAT', 'GCN')
    if dataset == 'Citeseer' or 'CV' in dataset or dataset == 'ogbn-arxiv':
        return (0.001 if L < 6 else 0.005) if attn_or_gcn else (0.0005 if L < 60 else 0.002)
    if dataset == 'Pubmed':
        return {'GCN': 0.001 if L < 6 else 0.01, 'GAT': 0.005 if L < 6 else 0.01}.get(type_model, 0.05)
    if dataset == 'Cora':
        return {'GCN': 0.001 if L < 6 else 0.03, 'GAT': 0.001 if L < 6 else 0.01}.get(
            type_model, 0.01 if L < 60 else 0.005)
    if dataset == 'CoauthorCS':
        return (0.001 if L < 6 else 0.03) if attn_or_gcn else (0.001 if L < 10 else .5)
    if dataset in ('CoauthorPhysics', 'AmazonComputers', 'AmazonPhoto', 'TEXAS', 'WISCONSIN', 'CORNELL'):
        return 0.005
    raise NotImplementedError


def reset_weight_GroupNorm(args):
    """Fills args.skip_weight / args.num_groups in place when the user did not set num_groups."""
    if args.num_groups is not None:
        return args
    args.miss_rate = 0.
    if args.dataset == 'CoauthorCS' and args.type_model not in ('GAT', 'GCN'):
        args.epochs = 500
    args.skip_weight = _skip_weight(args.dataset, args.type_model, args.num_layers)
    args.num_groups = 5 if args.dataset == 'Pubmed' else 10
    return args


def appendNormLayer(net, args, dim_to_norm=None):
    """Appends (at most) one norm layer to net.layers_norm; the trick name is matched by substring."""
    trick = args.type_trick
    dim = net.dim_hidden if dim_to_norm is None else dim_to_norm
    if 'BatchNorm' in trick:
        net.layers_norm.append(nn.BatchNorm1d(dim))
    elif 'PairNorm' in trick:
        net.layers_norm.append(pair_norm())
    elif 'NodeNorm' in trick:
        net.layers_norm.append(node_norm(**vars(net.args)))
    elif 'MeanNorm' in trick:
        net.layers_norm.append(mean_norm())
    elif 'GroupNorm' in trick:
        net.layers_norm.append(group_norm(dim_to_norm, **vars(reset_weight_GroupNorm(args))))
    elif 'CombNorm' in trick:
        net.layers_norm.append(comb_norm([group_norm(dim_to_norm, **vars(reset_weight_GroupNorm(args))),
                                          node_norm(**vars(net.args))]))


def norm_is_executed(type_trick):
    """run_norm_if_any's rule: the layer runs only when type_trick IS one of the norm names."""
    return type_trick in _RUNNABLE


def run_norm_if_any(net, x, ilayer):
    return net.layers_norm[ilayer](x) if norm_is_executed(net.args.type_trick) else x
