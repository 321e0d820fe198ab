"""CPU oracle for Cold Brew's TeacherGNN hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this file.  The product package
(``gnn_tail_generalization_b200``) never imports anything under ``oracle/``.

What it restates (all paths relative to /root/reference):

* ``GNN_model/GCN.py:184-258``   GCNConv.forward   -> :func:`gcn_conv`
* ``GNN_model/GCN.py:19-89``     TricksComb.__init__ -> :class:`OracleTricksComb.__init__`
* ``GNN_model/GCN.py:91-142``    TricksComb.forward  -> :meth:`OracleTricksComb.forward`
* ``GNN_model/res_tricks.py:7-55`` residual / initial / dense connections -> :func:`mix_*`
* ``GNN_model/norm_tricks.py:130-150`` construct-vs-run rule for norm layers
* ``GNN_model/GNN_normalizations.py:9-73`` TeacherGNN / GNN_norm wrappers
* ``trainer_node_classification.py:655-658`` + ``utils.py:667-674`` graph canonicalisation
* ``trainer_node_classification.py:390-394`` loss assembly

The aggregation primitive itself lives in an un-vendored dependency, **dgl==0.7.0**
(requirements.txt:19): ``graph.update_all(fn.copy_src('h','m'), fn.sum('m','h'))`` is DGL's
gSpMM('copy_lhs','sum'); on CPU it is ``SpMMSumCsr`` (dgl/src/array/cpu/spmm.h): for every
destination row, for every stored in-edge in CSR order, ``out[row,:] += X[src,:]``.  DGL
semantics restated here: multigraph (duplicate edges count twice), ``out_degrees`` counts
edges by source, ``in_degrees`` by destination, a node with no in-edge aggregates to 0.

Pinning status: the reference ships no tests for this path.  The oracle is pinned by
(1) the hand-derived known-answer vector on the reference's own toy graph
(``utils.py:1096``) and (2) fixtures under ``tests/golden/`` produced
by executing the reference's *own* ``GNN_model`` Python code in this container against a
~40-line stand-in for the DGL primitive (``tests/golden/make_golden.py``).  The DGL C++
kernel itself could not be executed here (wheel absent, no network), so the summation
order inside a row is the published one, not an observed one.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

_HERE = os.path.dirname(os.path.abspath(__file__))

# --------------------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------------------

def contains_any(name: str, needles) -> bool:
    """Substring test the reference uses to decode ``type_trick`` (norm_tricks.py:124-128)."""
    return any(n in name for n in needles)


_EXACT_NORM_NAMES = ('BatchNorm', 'PairNorm', 'NodeNorm', 'MeanNorm', 'GroupNorm', 'CombNorm')


# --------------------------------------------------------------------------------------
# graph canonicalisation and structure
# --------------------------------------------------------------------------------------

def symmetrize(edge_index: torch.Tensor) -> torch.Tensor:
    """utils.py:667-674 ``ensure_symmetric``: A + A^T, coalesced (sorted by (row, col), unique)."""
    n = int(edge_index.max()) + 1
    both = torch.cat([edge_index, edge_index.flip(0)], dim=1)
    key = torch.unique(both[0] * n + both[1])          # sorted, deduplicated
    return torch.stack([key // n, key % n])


def canonicalize_planetoid(edge_index: torch.Tensor, num_nodes: int) -> torch.Tensor:
    """trainer_node_classification.py:655-658: symmetrise, drop self loops, append one per node."""
    ei = symmetrize(edge_index)
    ei = ei[:, ei[0] != ei[1]]
    loops = torch.arange(num_nodes, dtype=ei.dtype)
    return torch.cat([ei, torch.stack([loops, loops])], dim=1)


def degree_inv_sqrt(edge_index: torch.Tensor, num_nodes: int):
    """GCN.py:205-209 / 242-246.  Returns (dout^-1/2, din^-1/2) as float32, degrees clamped to >=1."""
    src, dst = edge_index[0], edge_index[1]
    dout = torch.bincount(src, minlength=num_nodes).float().clamp(min=1)
    din = torch.bincount(dst, minlength=num_nodes).float().clamp(min=1)
    return torch.pow(dout, -0.5), torch.pow(din, -0.5)


def has_zero_in_degree(edge_index: torch.Tensor, num_nodes: int) -> bool:
    """GCN.py:187-188 guard."""
    return bool((torch.bincount(edge_index[1], minlength=num_nodes) == 0).any())


def build_csr(keys: np.ndarray, vals: np.ndarray, num_rows: int):
    """Stable counting sort of the edge list by ``keys``.

    Returns (rowptr[int64 num_rows+1], cols = vals[perm], perm) where ``perm[j]`` is the position in
    the original edge list of the j-th stored entry.  Stability fixes the in-row order to the original
    edge order, which is what a sequential ``out[dst] += h[src]`` over the COO list produces.
    """
    keys = np.asarray(keys)
    perm = np.argsort(keys, kind='stable')
    counts = np.bincount(keys, minlength=num_rows)
    rowptr = np.zeros(num_rows + 1, dtype=np.int64)
    np.cumsum(counts, out=rowptr[1:])
    return rowptr, np.asarray(vals)[perm], perm


# --------------------------------------------------------------------------------------
# aggregation primitive (DGL update_all(copy_src, sum))
# --------------------------------------------------------------------------------------

def aggregate_sum(h: torch.Tensor, edge_index: torch.Tensor, num_dst: int) -> torch.Tensor:
    """rst[v] = sum over edges (u->v) of h[u].  Differentiable; sequential edge order on CPU."""
    out = torch.zeros((num_dst,) + tuple(h.shape[1:]), dtype=h.dtype, device=h.device)
    return out.index_add_(0, edge_index[1], h[edge_index[0]])


_CLIB = None


def _c_oracle():
    """Load (building if needed) the C restatement of DGL's CPU SpMMSumCsr."""
    global _CLIB
    if _CLIB is None:
        so = os.path.join(_HERE, 'libcb_oracle.so')
        src = os.path.join(_HERE, 'spmm_sum_csr.c')
        if (not os.path.exists(so)) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(['make', '-s', '-C', _HERE, 'libcb_oracle.so'])
        lib = ctypes.CDLL(so)
        lib.cb_oracle_spmm_sum_csr.restype = ctypes.c_int
        lib.cb_oracle_spmm_sum_csr.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
            ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int]
        lib.cb_oracle_num_threads.restype = ctypes.c_int
        lib.cb_oracle_spmm_mul_sum_csr.restype = ctypes.c_int
        lib.cb_oracle_spmm_mul_sum_csr.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
            ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int]
        _CLIB = lib
    return _CLIB


def aggregate_sum_csr_ordered(h: np.ndarray, rowptr: np.ndarray, cols: np.ndarray,
                              hub_chunk: int = 0, threads: int = 0) -> np.ndarray:
    """In-order fp32 CSR row sums through the C oracle (bit-stable, any thread count).

    ``hub_chunk > 0`` reproduces the association the CUDA path uses for rows longer than
    ``hub_chunk``: consecutive chunks of ``hub_chunk`` entries are summed in order, then the chunk
    partials are summed in order.
    """
    lib = _c_oracle()
    h = np.ascontiguousarray(h, dtype=np.float32)
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
    cols = np.ascontiguousarray(cols, dtype=np.int64)
    nrows = rowptr.shape[0] - 1
    out = np.zeros((nrows, h.shape[1]), dtype=np.float32)
    rc = lib.cb_oracle_spmm_sum_csr(rowptr.ctypes.data, cols.ctypes.data, h.ctypes.data,
                                    h.shape[0], h.shape[1], out.ctypes.data, nrows,
                                    int(hub_chunk), int(threads))
    if rc != 0:
        raise RuntimeError(f'cb_oracle_spmm_sum_csr failed rc={rc}')
    return out


def aggregate_mul_sum_csr_ordered(h: np.ndarray, rowptr: np.ndarray, cols: np.ndarray, vals: np.ndarray,
                                  hub_chunk: int = 0, threads: int = 0) -> np.ndarray:
    """Edge-weighted in-order CSR row sums (GCN.py:199-202 ``u_mul_e`` + ``sum``): ``vals[j]`` is the weight of the
    stored edge j; every product is rounded to fp32 before it is added, chunks as in aggregate_sum_csr_ordered."""
    lib = _c_oracle()
    h = np.ascontiguousarray(h, dtype=np.float32)
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
    cols = np.ascontiguousarray(cols, dtype=np.int64)
    vals = np.ascontiguousarray(vals, dtype=np.float32)
    nrows = rowptr.shape[0] - 1
    out = np.zeros((nrows, h.shape[1]), dtype=np.float32)
    rc = lib.cb_oracle_spmm_mul_sum_csr(rowptr.ctypes.data, cols.ctypes.data, vals.ctypes.data, h.ctypes.data,
                                        h.shape[0], h.shape[1], out.ctypes.data, nrows, int(hub_chunk), int(threads))
    if rc != 0:
        raise RuntimeError(f'cb_oracle_spmm_mul_sum_csr failed rc={rc}')
    return out


def aggregate_mul_sum(h: torch.Tensor, edge_index: torch.Tensor, edge_weight: torch.Tensor, num_dst: int) -> torch.Tensor:
    """rst[v] = sum over edges e = (u->v) of h[u] * w[e] (GCN.py:199-202).  Differentiable in h and w."""
    w = edge_weight.reshape(-1, *([1] * (h.dim() - 1)))
    out = torch.zeros((num_dst,) + tuple(h.shape[1:]), dtype=h.dtype, device=h.device)
    return out.index_add_(0, edge_index[1], h[edge_index[0]] * w)


def c_oracle_threads() -> int:
    return int(_c_oracle().cb_oracle_num_threads())


class CsrPlan:
    """Both CSR views of a graph for the multi-threaded C aggregation (the timed CPU baseline).

    DGL builds the same two structures lazily (CSC for the forward gSpMM, CSR for its autograd
    transpose, dgl/python/dgl/backend/pytorch/sparse.py GSpMM.backward -> gspmm on the reversed graph).
    """

    def __init__(self, edge_index: torch.Tensor, num_nodes: int):
        src, dst = edge_index[0].numpy(), edge_index[1].numpy()
        self.n = num_nodes
        self.by_dst = build_csr(dst, src, num_nodes)[:2]
        self.by_src = build_csr(src, dst, num_nodes)[:2]


class _CsrAggregate(torch.autograd.Function):
    """update_all(copy_src, sum) through the OpenMP C restatement; backward = the transposed walk."""

    @staticmethod
    def forward(ctx, h, plan):
        ctx.plan = plan
        return torch.from_numpy(aggregate_sum_csr_ordered(h.detach().numpy(), *plan.by_dst))

    @staticmethod
    def backward(ctx, g):
        return torch.from_numpy(aggregate_sum_csr_ordered(g.contiguous().numpy(), *ctx.plan.by_src)), None


def aggregate_sum_planned(h: torch.Tensor, plan: CsrPlan) -> torch.Tensor:
    return _CsrAggregate.apply(h, plan)


# --------------------------------------------------------------------------------------
# one GCNConv layer (GCN.py:184-258)
# --------------------------------------------------------------------------------------

def gcn_conv(feat, edge_index, num_nodes, weight, bias=None, le=None, allow_zero_in_degree=False, plan=None,
             edge_weight=None):
    """Returns (rst, se_reg).  Order of operations is the reference's:

    scale source rows by dout^-1/2  ->  @ W  ->  + E (unscaled)  ->  sum over in-edges  ->
    scale by din^-1/2  ->  + bias.   se_reg = ||E||_F or None.
    """
    if not allow_zero_in_degree and has_zero_in_degree(edge_index, num_nodes):
        raise RuntimeError('There are 0-in-degree nodes in the graph')         # DGLError in the reference
    dout_is, din_is = degree_inv_sqrt(edge_index, num_nodes)
    dout_is, din_is = dout_is.to(feat.dtype), din_is.to(feat.dtype)
    h = feat * dout_is.reshape(-1, 1)                                          # GCN.py:205-213
    if weight is not None:
        h = torch.matmul(h, weight)                                            # GCN.py:225
    se_reg = None
    if le is not None:
        h = h + le                                                             # GCN.py:231
        # GCN.py:232 th.norm(self.le).  Accumulated in fp64 here: torch's CPU fp32 norm is off by 1.3e-4 at
        # Pubmed size (19 717 x 256), which is the accumulation error of that one kernel, not reference
        # semantics; the small golden fixtures still agree with the reference's fp32 value to 1e-6.
        se_reg = torch.linalg.vector_norm(le, dtype=torch.float64).to(le.dtype)
    if edge_weight is not None:                                                # GCN.py:199-202 (degrees stay counts)
        assert edge_weight.shape[0] == edge_index.shape[1]
        rst = aggregate_mul_sum(h, edge_index, edge_weight, num_nodes)
    else:
        rst = aggregate_sum(h, edge_index, num_nodes) if plan is None else aggregate_sum_planned(h, plan)  # GCN.py:238
    rst = rst * din_is.reshape(-1, 1)                                          # GCN.py:242-250
    if bias is not None:
        rst = rst + bias                                                       # GCN.py:252-253
    return rst, se_reg


# --------------------------------------------------------------------------------------
# residual tricks (res_tricks.py)
# --------------------------------------------------------------------------------------

def mix_residual(xs, alpha):      # res_tricks.py:12-14
    return xs[-1] if len(xs) == 1 else (1 - alpha) * xs[-1] + alpha * xs[-2]


def mix_initial(xs, alpha):       # res_tricks.py:21-23
    return xs[-1] if len(xs) == 1 else (1 - alpha) * xs[-1] + alpha * xs[0]


class OracleDense(nn.Module):
    """res_tricks.py:26-55; parameter names kept (``layer_transform`` / ``layer_att``)."""

    def __init__(self, in_dim, out_dim, aggregation):
        super().__init__()
        self.aggregation = aggregation
        if aggregation == 'concat':
            self.layer_transform = nn.Linear(in_dim, out_dim, bias=True)
        elif aggregation == 'attention':
            self.layer_att = nn.Linear(in_dim, 1, bias=True)

    def forward(self, xs):
        if self.aggregation == 'concat':
            return self.layer_transform(torch.cat(xs, dim=-1))
        if self.aggregation == 'maxpool':
            return torch.stack(xs, dim=-1).max(dim=-1).values
        if self.aggregation == 'attention':
            pps = torch.stack(xs, dim=1)                       # [N, k+1, c]
            score = torch.sigmoid(self.layer_att(pps).squeeze()).unsqueeze(1)
            return torch.matmul(score, pps).squeeze()
        raise Exception('Unknown aggregation')


class _Mix(nn.Module):
    def __init__(self, kind, alpha):
        super().__init__()
        self.kind, self.alpha = kind, alpha

    def forward(self, xs):
        return (mix_residual if self.kind == 'residual' else mix_initial)(xs, self.alpha)


# --------------------------------------------------------------------------------------
# norm layers (norm_tricks.py) -- constructed by substring, executed only on exact names (F4)
# --------------------------------------------------------------------------------------

class OraclePairNorm(nn.Module):            # norm_tricks.py:20-31
    def forward(self, x):
        x = x - x.mean(dim=0)
        return x / (1e-6 + x.pow(2).sum(dim=1).mean()).sqrt()


class OracleMeanNorm(nn.Module):            # norm_tricks.py:34-42
    def forward(self, x):
        return x - x.mean(dim=0)


class OracleNodeNorm(nn.Module):            # norm_tricks.py:45-85 (type "n" default)
    def __init__(self, node_norm_type='n', unbiased=False, eps=1e-5, power_root=2, **_):
        super().__init__()
        self.t, self.unbiased, self.eps, self.power = node_norm_type, unbiased, eps, 1 / power_root

    def forward(self, x):
        std = (torch.var(x, unbiased=self.unbiased, dim=1, keepdim=True) + self.eps).sqrt()
        if self.t == 'n':
            return (x - x.mean(dim=1, keepdim=True)) / std
        if self.t == 'v':
            return x / std
        if self.t == 'm':
            return x - x.mean(dim=1, keepdim=True)
        if self.t == 'srv':
            return x / torch.sqrt(std)
        if self.t == 'pr':
            return x / torch.pow(std, self.power)
        return x


class OracleGroupNorm(nn.Module):           # norm_tricks.py:97-121
    def __init__(self, dim, num_groups, skip_weight):
        super().__init__()
        self.num_groups, self.skip_weight, self.dim_hidden = num_groups, skip_weight, dim
        self.bn = nn.BatchNorm1d(dim * num_groups, momentum=0.3)
        self.group_func = nn.Linear(dim, num_groups, bias=True)

    def forward(self, x):
        if self.num_groups == 1:
            t = self.bn(x)
        else:
            s = F.softmax(self.group_func(x), dim=1)
            t = torch.cat([s[:, g].unsqueeze(1) * x for g in range(self.num_groups)], dim=1)
            t = self.bn(t).view(-1, self.num_groups, self.dim_hidden).sum(dim=1)
        return x + t * self.skip_weight


class OracleCombNorm(nn.Module):            # norm_tricks.py:9-17
    def __init__(self, mods):
        super().__init__()
        self.norm_list = nn.ModuleList(mods)

    def forward(self, x):
        for m in self.norm_list:
            x = m(x)
        return x


def groupnorm_hparams(args):
    """norm_tricks.py:153-206 ``reset_weight_GroupNorm`` as a table: (num_groups, skip_weight)."""
    if getattr(args, 'num_groups', None) is not None:
        return args.num_groups, args.skip_weight
    ds, tm, L = args.dataset, args.type_model, args.num_layers
    gat_gcn = tm in ('GAT', 'GCN')
    if ds == 'Citeseer' or 'CV' in ds or ds == 'ogbn-arxiv':
        sw = (0.001 if L < 6 else 0.005) if gat_gcn else (0.0005 if L < 60 else 0.002)
    elif ds == 'Pubmed':
        sw = (0.001 if L < 6 else 0.01) if tm == 'GCN' else ((0.005 if L < 6 else 0.01) if tm == 'GAT' else 0.05)
    elif ds == 'Cora':
        sw = (0.001 if L < 6 else 0.03) if tm == 'GCN' else (
            (0.001 if L < 6 else 0.01) if tm == 'GAT' else (0.01 if L < 60 else 0.005))
    elif ds == 'CoauthorCS':
        sw = (0.001 if L < 6 else 0.03) if gat_gcn else (0.001 if L < 10 else .5)
    elif ds in ('CoauthorPhysics', 'AmazonComputers', 'AmazonPhoto', 'TEXAS', 'WISCONSIN', 'CORNELL'):
        sw = 0.005
    else:
        raise NotImplementedError
    return (5 if ds == 'Pubmed' else 10), sw


def make_norm_layer(args, dim):
    """norm_tricks.py:130-143 ``appendNormLayer``; returns None when nothing is appended."""
    t = args.type_trick
    if 'BatchNorm' in t:
        return nn.BatchNorm1d(dim)
    if 'PairNorm' in t:
        return OraclePairNorm()
    if 'NodeNorm' in t:
        return OracleNodeNorm(**{k: v for k, v in vars(args).items()
                                 if k in ('node_norm_type', 'unbiased', 'eps', 'power_root')})
    if 'MeanNorm' in t:
        return OracleMeanNorm()
    if 'GroupNorm' in t:
        g, sw = groupnorm_hparams(args)
        return OracleGroupNorm(dim, g, sw)
    if 'CombNorm' in t:
        g, sw = groupnorm_hparams(args)
        return OracleCombNorm([OracleGroupNorm(dim, g, sw), OracleNodeNorm()])
    return None


# --------------------------------------------------------------------------------------
# the layer stack (GCN.py:18-150) and wrappers (GNN_normalizations.py)
# --------------------------------------------------------------------------------------

class OracleGCNConv(nn.Module):
    """Parameter container for one layer; init order weight -> bias -> le (GCN.py:170-182,260-264)."""

    def __init__(self, in_feats, out_feats, n_nodes, has_se):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(in_feats, out_feats))
        self.bias = nn.Parameter(torch.empty(out_feats))
        nn.init.xavier_uniform_(self.weight)
        nn.init.zeros_(self.bias)
        self.has_se = bool(has_se)
        if self.has_se:
            self.le = nn.Parameter(torch.randn(n_nodes, out_feats))


class OracleTricksComb(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.args = args
        self.plan = None          # set to a CsrPlan to aggregate through the OpenMP C path (CPU baseline timing)
        t = args.type_trick
        L, H, Fin, C = args.num_layers, args.dim_hidden, args.num_feats, args.num_classes
        se = args.TeacherGNN.whetherHasSE
        self.has_residual_MLP = contains_any(t, ('Jumping', 'Initial', 'Residual', 'Dense'))
        self.layers_GCN, self.layers_res = nn.ModuleList(), nn.ModuleList()
        self.layers_norm, self.layers_MLP = nn.ModuleList(), nn.ModuleList()
        self.layers_MLP.append(nn.Linear(Fin, H))                                         # GCN.py:43
        if not self.has_residual_MLP:
            self.layers_GCN.append(OracleGCNConv(Fin, H, args.N_nodes, se[0]))           # GCN.py:45
        for i in range(L):
            if self.has_residual_MLP or 0 < i < L - 1:                                    # GCN.py:48-52 (F5: flag [1])
                self.layers_GCN.append(OracleGCNConv(H, H, args.N_nodes, se[1]))
            nl = make_norm_layer(args, H if i < L - 1 else C)                             # GCN.py:54
            if nl is not None:
                self.layers_norm.append(nl)
            if 'Residual' in t:                                                           # GCN.py:57-67
                self.layers_res.append(_Mix('residual', args.res_alpha))
            elif 'Initial' in t:
                self.layers_res.append(_Mix('initial', args.res_alpha))
            elif 'Dense' in t:
                if args.layer_agg in ('concat', 'maxpool'):
                    self.layers_res.append(OracleDense((i + 2) * H, H, args.layer_agg))
                elif args.layer_agg == 'attention':
                    self.layers_res.append(OracleDense(H, H, args.layer_agg))
        if not self.has_residual_MLP:
            self.layers_GCN.append(OracleGCNConv(H, C, args.N_nodes, se[2]))              # GCN.py:71
        if 'Jumping' in t:                                                                # GCN.py:73-81
            if args.layer_agg in ('concat', 'maxpool'):
                self.layers_res.append(OracleDense((L + 1) * H, C, args.layer_agg))
            elif args.layer_agg == 'attention':
                self.layers_res.append(OracleDense(H, C, args.layer_agg))
        else:
            self.layers_MLP.append(nn.Linear(H, C))

    def forward(self, x, edge_index, want_les=False, relu_masks=None):
        """relu_masks (tests only): a list of 0/1 tensors, consumed in order, that replace every ``F.relu(v)`` by
        ``v * mask`` -- the gates of ANOTHER run of the same model.  A pre-activation within rounding distance of 0
        gates differently in fp32 and fp64; with the gates pinned, a gradient comparison measures the arithmetic and
        not the handful of flipped gates."""
        a, t = self.args, self.args.type_trick
        n = x.shape[0]
        xs, les, se_reg_all = [], [], None
        masks = list(relu_masks) if relu_masks is not None else None

        def relu(v):
            return F.relu(v) if masks is None else v * masks.pop(0).to(v.dtype)
        if self.has_residual_MLP:                                                         # GCN.py:103-107
            x = F.dropout(x, p=a.dropout, training=self.training)
            x = relu(self.layers_MLP[0](x))
            xs.append(x)
        for i in range(a.num_layers):                                                     # GCN.py:109-131
            x = F.dropout(x, p=a.dropout, training=self.training)
            lyr = self.layers_GCN[i]
            x, reg = gcn_conv(x, edge_index, n, lyr.weight, lyr.bias, lyr.le if lyr.has_se else None, plan=self.plan)
            if reg is not None:
                se_reg_all = reg if se_reg_all is None else se_reg_all + reg
            if t in _EXACT_NORM_NAMES:                                                    # norm_tricks.py:146-150
                x = self.layers_norm[i](x)
            if want_les:
                les.append(x.clone().detach())
            if self.has_residual_MLP or i < a.num_layers - 1:
                x = relu(x)
            xs.append(x)
            if contains_any(t, ('Initial', 'Dense', 'Residual')):
                x = self.layers_res[i](xs)
        x = F.dropout(x, p=a.dropout, training=self.training)                             # GCN.py:133
        if self.has_residual_MLP:
            x = self.layers_res[0](xs) if 'Jumping' in t else self.layers_MLP[-1](x)      # GCN.py:134-138
        if want_les:
            return x, se_reg_all, torch.cat(les, dim=-1)
        return x, se_reg_all


class OracleGNNNorm(nn.Module):            # GNN_normalizations.py:67-73
    def __init__(self, args):
        super().__init__()
        self.model = OracleTricksComb(args)

    def forward(self, x, edge_index):
        return self.model(x, edge_index)


class OracleTeacherGNN(nn.Module):         # GNN_normalizations.py:9-65
    def __init__(self, args, proj2class=None):
        super().__init__()
        args.num_classes_bkup = args.num_classes
        args.num_classes = args.dim_commonEmb
        self.args = args
        if args.dim_learnable_input > 0:
            self.embs = nn.Parameter(torch.randn(args.N_nodes, args.dim_learnable_input) * 0.001)
            args.num_feats_bkup = args.num_feats
            args.num_feats = args.dim_learnable_input
        self.model = OracleGNNNorm(args)
        self.proj2linkp = nn.Identity()
        self.proj2class = proj2class or nn.Identity()
        self.se_reg_all = None

    def forward(self, x, edge_index):
        if self.args.TeacherGNN.change_to_featureless:
            x = x * 0
        if self.args.dim_learnable_input > 0:
            x = self.embs
        out, self.se_reg_all = self.model(x, edge_index)
        return out

    def get_3_embs(self, x, edge_index, mask=None, want_heads=True):
        common = self.forward(x, edge_index)
        full = self.proj2class(common)
        res = SimpleNamespace(commonEmb=common, emb4classi_full=full, emb4classi=None, emb4linkp=None)
        if want_heads:
            res.emb4classi = full[mask] if mask is not None else full
            res.emb4linkp = self.proj2linkp(common)
        return res


def teacher_loss(model: OracleTeacherGNN, x, edge_index, y, train_mask, se_reg_coef, lossa_semantic=1.0):
    """trainer_node_classification.py:386-394: nll(log_softmax(logits[mask])) + se_reg * sum_l ||E_l||_F."""
    res = model.get_3_embs(x, edge_index, train_mask)
    loss = F.nll_loss(F.log_softmax(res.emb4classi, 1), y[train_mask]) * lossa_semantic
    if model.se_reg_all is not None:
        loss = loss + se_reg_coef * model.se_reg_all
    return loss


# --------------------------------------------------------------------------------------
# bf16 STORAGE restatement (BASELINE.json configs[4]: "128-dim bf16"): the reference arithmetic of the Initial
# topology (GCN.py:91-142 with type_trick 'Initial', no norm layer, no dropout) with every matrix that the CUDA
# path keeps in HBM rounded to bf16 where it is stored -- features, the SE tables as read by the forward pass, the
# dense weights as fed to the tensor cores -- and every sum, scale, bias, relu and mix in fp32.  Gradients crossing
# the same storage points are rounded too (the CUDA path keeps dlogits, G, dH and dX in bf16).
# --------------------------------------------------------------------------------------
class _RoundBf16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


def rbf16(x):
    """Value as stored in bf16 (round to nearest even), carried on in fp32; the gradient is rounded the same way."""
    return _RoundBf16.apply(x)


def bf16_storage_forward(model: 'OracleTeacherGNN', x: torch.Tensor, edge_index: torch.Tensor):
    """logits [N, C] (fp32 values representable in bf16) and se_reg_all for an 'Initial' TeacherGNN."""
    tc = model.model.model
    a = tc.args
    assert 'Initial' in a.type_trick and not contains_any(a.type_trick, _EXACT_NORM_NAMES + ('Jumping', 'Dense', 'Residual'))
    n = x.shape[0]
    dout_is, din_is = degree_inv_sqrt(edge_index, n)
    lin0, head = tc.layers_MLP[0], tc.layers_MLP[-1]
    x = rbf16(x)
    x0 = rbf16(F.relu(x @ rbf16(lin0.weight).t() + lin0.bias))                 # GCN.py:104-107
    cur_scaled, se_reg_all, out = None, None, None
    for i in range(a.num_layers):
        lyr = tc.layers_GCN[i]
        w = rbf16(lyr.weight)
        if cur_scaled is None:
            h = dout_is.reshape(-1, 1) * (x0 @ w)                               # GCN.py:205-225 ((D X) W = D (X W))
        else:
            h = cur_scaled @ w
        if lyr.has_se:
            h = h + rbf16(lyr.le)                                               # GCN.py:231 (bf16 shadow of the table)
            reg = torch.linalg.vector_norm(lyr.le, dtype=torch.float64).to(lyr.le.dtype)   # fp32 master
            se_reg_all = reg if se_reg_all is None else se_reg_all + reg
        h = rbf16(h)
        z = aggregate_sum(h, edge_index, n) * din_is.reshape(-1, 1) + lyr.bias  # GCN.py:238-253
        o = (1 - a.res_alpha) * F.relu(z) + a.res_alpha * x0                    # GCN.py:127-131, res_tricks.py:23
        out = rbf16(o)
        cur_scaled = rbf16(dout_is.reshape(-1, 1) * o)                          # next layer's (D X), stored
    logits = rbf16(out @ rbf16(head.weight).t() + head.bias)                    # GCN.py:137-138
    return logits, se_reg_all


def make_args(**kw):
    """Namespace with the fields TricksComb/TeacherGNN read, reference defaults."""
    d = dict(type_trick='NoRes', type_model='GCN', num_layers=2, dim_hidden=64, num_feats=16, num_classes=7,
             dropout=0.0, res_alpha=0.1, layer_agg='concat', transductive=True, N_nodes=0, device='cpu',
             dataset='Cora', dim_learnable_input=0, lamda=0.5, num_groups=None, skip_weight=None,
             graph_dropout=0.0, layerwise_dropout=False, whetherHasSE=(0, 0, 0), change_to_featureless=False,
             dim_commonEmb=None)
    d.update(kw)
    se = d.pop('whetherHasSE')
    if isinstance(se, str):
        se = [int(c) for c in se]
    featureless = d.pop('change_to_featureless')
    a = SimpleNamespace(**d)
    a.TeacherGNN = SimpleNamespace(whetherHasSE=list(se), change_to_featureless=featureless)
    if a.dim_commonEmb is None:
        a.dim_commonEmb = a.num_classes
    return a


# --------------------------------------------------------------------------------------
# synthetic graphs shared by tests and the bench
# --------------------------------------------------------------------------------------

def powerlaw_graph(num_nodes: int, num_undirected: int, seed: int = 0, gamma: float = 2.5,
                   device='cpu') -> torch.Tensor:
    """Chung-Lu style power-law graph, canonicalised like the trainer does (symmetric, no
    duplicates, exactly one self loop per node).  Endpoint i is drawn with probability
    proportional to (i+1)^(-1/(gamma-1)) via the inverse CDF, then ids are permuted."""
    g = torch.Generator(device='cpu').manual_seed(seed)
    expo = 1.0 / (1.0 - 1.0 / (gamma - 1.0))          # inverse-CDF exponent: i = N * U^expo
    draw = int(num_undirected * 1.25) + 16
    u = torch.rand(2, draw, generator=g, dtype=torch.float64)
    ends = (num_nodes * u.pow(expo)).long().clamp_(max=num_nodes - 1)
    perm = torch.randperm(num_nodes, generator=g)
    ends = perm[ends]
    lo, hi = torch.minimum(ends[0], ends[1]), torch.maximum(ends[0], ends[1])
    keep = lo != hi
    key = torch.unique(lo[keep] * num_nodes + hi[keep])
    if key.numel() > num_undirected:
        sel = torch.randperm(key.numel(), generator=g)[:num_undirected].sort().values
        key = key[sel]
    lo, hi = key // num_nodes, key % num_nodes
    loops = torch.arange(num_nodes)
    src = torch.cat([lo, hi, loops])
    dst = torch.cat([hi, lo, loops])
    return torch.stack([src, dst]).to(device)
