"""Cold Brew's GCN layer and layer stack on the B200 kernels.

Mirror of the reference's GNN_model/GCN.py (``GCNConv`` :152-277, ``TricksComb`` :18-150): same class
names, constructor signatures, attribute names and ``state_dict`` keys
(``layers_GCN.{i}.{weight,bias,le}``, ``layers_MLP``, ``layers_norm``, ``layers_res``), same order of
parameter creation (so ``set_seed`` reproduces the reference's initial weights), same return values.

What changed underneath:
  * the DGL graph (GCN.py:92-94) is a device-built ``GraphHandle`` (CSR both ways, cached degree^-1/2);
  * ``update_all(copy_src, sum)`` + in-degree scale + bias (+ relu + Initial mix + the next layer's
    out-degree scale) is ONE kernel, ``cb_agg_forward``; its backward is ``cb_agg_backward_prep`` +
    the transposed gather ``cb_agg_gather``;
  * the zero-in-degree check (GCN.py:187-197) reads a flag computed once at graph build instead of
    synchronising the device in every layer of every forward.
The dense ``X W`` is a true GEMM and runs on the tensor cores: ``cb_gemm_rows`` (tcgen05, 3xTF32 split for
fp32-class accuracy) with the out-degree scale, the SE add, the Linear bias and relu in its epilogue;
weight gradients use ``cb_gemm_tn``.  Shapes those kernels do not cover fall back to cuBLAS fp32.
"""
import math

import torch as th
import torch.nn.functional as F
from torch import nn
from torch.nn import init

from ._backend import DGLError, graph as _graph, ops as _ops
from .drop_tricks import DropoutTrick
from .norm_tricks import AcontainsB, appendNormLayer, norm_is_executed, run_norm_if_any
from .res_tricks import DenseConnection, InitialConnection, ResidualConnection

_ZERO_IN_DEGREE_MSG = ('There are 0-in-degree nodes in the graph, output for those nodes will be invalid. '
                       'This is harmful for some applications, causing silent performance regression. '
                       'Adding self-loop on the input graph will resolve the issue. Setting '
                       '``allow_zero_in_degree`` to be `True` when constructing this module will suppress '
                       'the check and let the code run.')


class GCNConv(nn.Module):
    """out = D_in^-1/2 . A^T-sum( (D_out^-1/2 X) W + E ) + b ;  returns (out, ||E||_F or None)."""

    def __init__(self, in_feats, out_feats, norm='both', weight=True, bias=True, activation=None,
                 allow_zero_in_degree=False, cached=None, args=None, whetherHasSE=False):
        super().__init__()
        if norm not in ('none', 'both', 'right', 'left'):
            raise DGLError(f'Invalid norm value. Must be either "none", "both", "right" or "left". But got "{norm}".')
        self.args = args
        self._in_feats, self._out_feats, self._norm = in_feats, out_feats, norm
        self._allow_zero_in_degree = allow_zero_in_degree
        # creation order weight -> bias -> le is part of the RNG contract (GCN.py:170-182)
        if weight:
            self.weight = nn.Parameter(th.Tensor(in_feats, out_feats))
        else:
            self.register_parameter('weight', None)
        if bias:
            self.bias = nn.Parameter(th.Tensor(out_feats))
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()
        self._activation = activation
        self.whetherHasSE = whetherHasSE
        self.se_fused = None       # set by se_optim.FusedSEAdam: the table is then stepped by cb_se_adam_step
        if whetherHasSE:
            self.le = nn.Parameter(th.randn(args.N_nodes, self._out_feats), requires_grad=True)

    def reset_parameters(self):
        if self.weight is not None:
            init.xavier_uniform_(self.weight)
        if self.bias is not None:
            init.zeros_(self.bias)

    def set_allow_zero_in_degree(self, set_value):
        self._allow_zero_in_degree = set_value

    def extra_repr(self):
        s = f'in={self._in_feats}, out={self._out_feats}, normalization={self._norm}'
        if '_activation' in self.__dict__:
            s += f', activation={self._activation}'
        return s

    # -- pieces ---------------------------------------------------------------------------------
    def _check(self, graph, weight):
        if not self._allow_zero_in_degree and graph.has_zero_in_degree:
            raise DGLError(_ZERO_IN_DEGREE_MSG)
        if weight is not None and self.weight is not None:
            raise DGLError('External weight is provided while at the same time the module has defined its own '
                           'weight parameter. Please create the module with flag weight=False.')
        return weight if weight is not None else self.weight

    def _transform(self, graph, feat, weight, row_scale=None, dx_sink=None, dx_plan=None):
        """row_scale * (X W) + E  ( = (D_out^-1/2 X) W + E, GCN.py:205-231) and the SE regulariser
        (GCN.py:232-236).  One tcgen05 GEMM with the scale and the SE add in its epilogue."""
        le = self.le if self.whetherHasSE else None
        fused = getattr(self, 'se_fused', None) if self.whetherHasSE else None   # se_optim.FusedSEAdam owns the table
        add_sink = None
        if fused is not None:
            le, add_sink = fused.operand(feat.dtype), fused.slot
        elif le is not None and le.dtype != feat.dtype:
            le = le.to(feat.dtype)      # bf16 forward without the fused optimizer: an autograd cast per call
        if weight is not None:
            h, _ = _ops.dense(feat, weight, 'kn', add=le, row_scale=row_scale, dx_sink=dx_sink, dx_plan=dx_plan,
                              push_graph=graph, add_sink=add_sink)
        else:
            if fused is not None:
                raise DGLError('the fused SE optimizer needs the layer to own its weight')
            h = feat if row_scale is None else _ops.row_scale(feat, row_scale)
            if le is not None:
                h = h + le
        if not self.whetherHasSE:
            return h, None
        return h, (fused.norm(graph) if fused is not None else _ops.frob_norm(self.le, graph))

    def fused(self, graph, feat, prescaled=False, relu=False, x0=None, alpha=0.0, want_out=True,
              want_scaled=False, weight=None, x0_sink=None, dx_sink=None, my_plan=None, dx_plan=None,
              row_sparse_hint=False, dropout_p=0.0):
        """The whole layer plus what follows it in TricksComb, norm='both' only.

        feat        layer input; if ``prescaled`` it already carries the D_out^-1/2 factor
        relu/x0     epilogue: out = (1-alpha) * relu(z) + alpha * x0
        want_scaled also return D_out^-1/2 * out (the next layer's pre-scaled input)
        my_plan     BwdPlan for this layer's backward prologue (run by the consumer of its output)
        dx_plan     BwdPlan of the op that produced ``feat`` (run by this layer's dX GEMM)
        dropout_p   > 0: ``out`` comes back already dropped (the F.dropout in front of its consumer, same mask)
        Returns (out, out_scaled, se_reg).
        """
        assert self._norm == 'both'
        weight = self._check(graph, weight)
        h, se_reg = self._transform(graph, feat, weight, None if prescaled else graph.dout_inv_sqrt, dx_sink,
                                    dx_plan)
        out, out_scaled = _ops.fused_aggregate(h, graph, self.bias, x0, alpha, relu, want_out, want_scaled,
                                               x0_sink, my_plan, row_sparse_hint, dropout_p)
        return out, out_scaled, se_reg

    def forward(self, graph, feat, weight=None, edge_weight=None):
        """Same contract as the reference's GCNConv.forward (GCN.py:184-258)."""
        if edge_weight is not None:
            # GCN.py:199-202: the messages become h[u] * w[e] (fn.u_mul_e); the degree norms stay edge COUNTS
            # (GCN.py:205-213, 242-250).  Never reached from TricksComb (it does not pass edge weights): the plain
            # composition, every piece on this library's kernels, no fused epilogue.
            assert edge_weight.shape[0] == graph.number_of_edges()
            weight = self._check(graph, weight)
            scale_src = None
            if self._norm == 'both':
                scale_src = graph.dout_inv_sqrt
            elif self._norm == 'left':
                scale_src = 1.0 / graph.out_degrees().float().clamp(min=1)
            h, se_reg = self._transform(graph, feat, weight, scale_src)
            rst = _ops.weighted_sum(h, edge_weight, graph)
            if self._norm == 'both':
                rst = _ops.row_scale(rst, graph.din_inv_sqrt)
            elif self._norm == 'right':
                rst = rst * (1.0 / graph.in_degrees().float().clamp(min=1)).unsqueeze(-1)
            if self.bias is not None:
                rst = rst + self.bias
            if self._activation is not None:
                rst = self._activation(rst)
            return rst, se_reg
        if self._norm == 'both':
            rst, _, se_reg = self.fused(graph, feat, weight=weight)
        else:
            weight = self._check(graph, weight)
            x = feat
            if self._norm == 'left':
                x = x * (1.0 / graph.out_degrees().float().clamp(min=1)).unsqueeze(-1)
            h, se_reg = self._transform(graph, x, weight)
            rst = _ops.copy_sum(h, graph)
            if self._norm == 'right':
                rst = rst * (1.0 / graph.in_degrees().float().clamp(min=1)).unsqueeze(-1)
            if self.bias is not None:
                rst = rst + self.bias
        if self._activation is not None:
            rst = self._activation(rst)
        return rst, se_reg


class TricksComb(nn.Module):
    """The layer stack with the residual / norm / dropout tricks (GCN.py:18-150)."""

    def __init__(self, args):
        super().__init__()
        self.args = args
        self.dglgraph = None
        self.alpha = args.res_alpha
        self.embedding_dropout = args.dropout
        for k, v in vars(args).items():
            setattr(self, k, v)
        self.cached = self.transductive = args.transductive
        if AcontainsB(self.type_trick, ['DropEdge', 'DropNode', 'FastGCN', 'LADIES']):
            self.cached = False
        self.has_residual_MLP = AcontainsB(self.type_trick, ['Jumping', 'Initial', 'Residual', 'Dense'])
        se = self.args.TeacherGNN.whetherHasSE

        self.layers_GCN = nn.ModuleList([])
        self.layers_res = nn.ModuleList([])
        self.layers_norm = nn.ModuleList([])
        self.layers_MLP = nn.ModuleList([])

        def conv(i, o, flag):
            return GCNConv(i, o, cached=self.cached, args=self.args, whetherHasSE=flag)

        self.layers_MLP.append(nn.Linear(self.num_feats, self.dim_hidden))
        if not self.has_residual_MLP:
            self.layers_GCN.append(conv(self.num_feats, self.dim_hidden, se[0]))
        for i in range(self.num_layers):
            # every hidden->hidden layer takes the MIDDLE SE flag, also under the residual topologies
            if self.has_residual_MLP or 0 < i < self.num_layers - 1:
                self.layers_GCN.append(conv(self.dim_hidden, self.dim_hidden, se[1]))
            appendNormLayer(self, args, self.dim_hidden if i < self.num_layers - 1 else self.num_classes)
            if AcontainsB(self.type_trick, ['Residual']):
                self.layers_res.append(ResidualConnection(alpha=self.alpha))
            elif AcontainsB(self.type_trick, ['Initial']):
                self.layers_res.append(InitialConnection(alpha=self.alpha))
            elif AcontainsB(self.type_trick, ['Dense']):
                if self.layer_agg in ['concat', 'maxpool']:
                    self.layers_res.append(DenseConnection((i + 2) * self.dim_hidden, self.dim_hidden, self.layer_agg))
                elif self.layer_agg == 'attention':
                    self.layers_res.append(DenseConnection(self.dim_hidden, self.dim_hidden, self.layer_agg))
        self.graph_dropout = DropoutTrick(args)
        if not self.has_residual_MLP:
            self.layers_GCN.append(conv(self.dim_hidden, self.num_classes, se[2]))
        if AcontainsB(self.type_trick, ['Jumping']):
            if self.layer_agg in ['concat', 'maxpool']:
                self.layers_res.append(
                    DenseConnection((self.num_layers + 1) * self.dim_hidden, self.num_classes, self.layer_agg))
            elif self.layer_agg == 'attention':
                self.layers_res.append(DenseConnection(self.dim_hidden, self.num_classes, self.layer_agg))
        else:
            self.layers_MLP.append(nn.Linear(self.dim_hidden, self.num_classes))

        if AcontainsB(self.type_trick, ['IdentityMapping']):
            self.lamda = args.lamda
        elif self.type_model == 'SGC':
            self.lamda = 0.
        elif self.type_model == 'GCN':
            self.lamda = 1.

    # -- graph ------------------------------------------------------------------------------------
    def build_graph(self, x, edge_index):
        """edge_index [2,E] int64 (device) -> GraphHandle, built on the device (replaces GCN.py:92-94)."""
        return _graph.GraphHandle(edge_index, x.shape[0])

    def forward(self, x, edge_index, want_les=False):
        if self.dglgraph is None:
            self.dglgraph = self.build_graph(x, edge_index)
        graph = self.dglgraph
        trick, L = self.type_trick, self.num_layers
        x_list, le_collection, se_reg_all, x0_sink = [], [], None, None
        self.graph_dropout(edge_index)  # result unused by the layers, exactly like GCN.py:101-115

        # plan of the op whose output is the current x, handed to x's single consumer (see ops.BwdPlan)
        prev_plan = None
        norm_runs = norm_is_executed(trick)
        initial = AcontainsB(trick, ['Initial'])
        mixes = AcontainsB(trick, ['Initial', 'Dense', 'Residual'])
        keeps_history = AcontainsB(trick, ['Residual', 'Dense', 'Jumping'])   # x_list entries are re-read later
        # the Initial mix rides on the aggregation epilogue only if nothing else reads the pre-mix activations:
        # no norm layer in between, no want_les, and no history consumer (x_list holds the PRE-mix relu outputs,
        # GCN.py:127-131, which a "Jumping"/"Dense"/"Residual" name re-reads)
        mix_will_fuse = initial and not norm_runs and not want_les and not keeps_history
        if self.has_residual_MLP:
            x = F.dropout(x, p=self.embedding_dropout, training=self.training)
            lin = self.layers_MLP[0]
            lin_plan = _ops.new_plan()
            x, _ = _ops.dense(x, lin.weight, 'nk', bias=lin.bias, relu=True, my_plan=lin_plan)
            if mix_will_fuse and x.requires_grad:
                # x0 feeds every layer's residual mix, and the mix is fused into the aggregation epilogue: collect
                # those gradients inside the backward kernels.  (With an un-fused mix -- norm layer in between, or
                # want_les -- x0 also receives plain autograd gradients from layers_res[i](x_list); the Linear then
                # runs its own relu/bias backward on their sum.)
                x0_sink = _ops.GradSink()
                x = _ops.sink_hub(x, x0_sink)
                # every gradient of x0 ends in layer 0's dX GEMM (the residual shares are parked in the
                # sink and added in its epilogue), so that GEMM can also run the Linear's relu/bias backward
                prev_plan = lin_plan
            x_list.append(x)

        no_drop = (not self.training) or self.dropout == 0
        xs_next = None   # D_out^-1/2-scaled copy of x produced by the previous layer's epilogue
        dropped = False  # x already went through the dropout that stands in front of its consumer (the layer owned it)

        for i in range(L):
            x_in = None
            if xs_next is None:
                x_in = x if (no_drop or dropped) else F.dropout(x, p=self.dropout, training=self.training)
            layer = self.layers_GCN[i]
            want_relu = self.has_residual_MLP or i < L - 1
            # the epilogue (relu, Initial mix) can ride on the aggregation kernel unless a norm layer
            # really runs between them or the caller wants the pre-activation values
            fuse_tail = not norm_runs and not want_les
            relu_fused = fuse_tail and want_relu
            mix_fused = mix_will_fuse and len(x_list) >= 1
            last = i == L - 1
            feeds_conv = (not last) and no_drop and (mix_fused or not mixes) and fuse_tail
            need_plain = last or keeps_history or not feeds_conv
            # this layer's output has one consumer whose backward is a GEMM: the next conv (pre-scaled
            # copy only), or the final Linear (plain output, nothing in between)
            single_consumer = fuse_tail and not keeps_history and (
                (feeds_conv and not need_plain) or
                (last and self.has_residual_MLP and not AcontainsB(trick, ['Jumping']) and
                 (mix_fused or not mixes) and ((not self.training) or self.args.dropout == 0)))
            my_plan = _ops.new_plan() if single_consumer else None
            if my_plan is not None and last:
                my_plan.row_sparse_hint = True     # the layer under the output head (see ops.BwdPlan)
            # Training with dropout (the reference's defaults): where the aggregation's output goes straight into the
            # dropout in front of its consumer, the layer draws that dropout itself (same torch call, same mask) and
            # runs its backward inside its own backward prologue.
            p_next = self.args.dropout if last else self.dropout
            owns_dropout = (self.training and 0 < p_next < 1 and fuse_tail and not keeps_history and
                            (mix_fused or not mixes) and _ops.dropout_fusion())
            dx_plan = prev_plan if (xs_next is not None or (x_in is not None and x_in is x)) else None
            out, out_scaled, se_reg = layer.fused(
                graph, xs_next if xs_next is not None else x_in, prescaled=xs_next is not None,
                relu=relu_fused, x0=x_list[0] if mix_fused else None, alpha=self.alpha,
                want_out=need_plain, want_scaled=feeds_conv, x0_sink=x0_sink if mix_fused else None,
                # layer 0 reads x0 itself: its dX GEMM adds the parked residual gradients in its epilogue
                dx_sink=x0_sink if (x0_sink is not None and x_in is not None and x_in is x_list[0]) else None,
                my_plan=my_plan, dx_plan=dx_plan,
                # without the hand-off plan (dropout in front of the head: the reference's training defaults) the
                # last layer still looks for the all-zero rows of its gradient itself
                row_sparse_hint=last and my_plan is None and self.training,
                dropout_p=p_next if owns_dropout else 0.0)
            dropped = owns_dropout
            prev_plan = my_plan
            if se_reg is not None:
                se_reg_all = se_reg if se_reg_all is None else se_reg_all + se_reg
            x = out
            if not fuse_tail:
                x = run_norm_if_any(self, x, i)
                if want_les:
                    le_collection.append(x.clone().detach())
                if want_relu:
                    x = F.relu(x)
            x_list.append(x)
            if mixes and not mix_fused:
                x = self.layers_res[i](x_list)
            xs_next = out_scaled if feeds_conv else None

        if not dropped:
            x = F.dropout(x, p=self.args.dropout, training=self.training)
        if self.has_residual_MLP:
            if AcontainsB(trick, ['Jumping']):
                x = self.layers_res[0](x_list)
            else:
                lin = self.layers_MLP[-1]
                x, _ = _ops.dense(x, lin.weight, 'nk', bias=lin.bias, dx_plan=prev_plan)
        if want_les:
            return x, se_reg_all, th.cat(le_collection, dim=-1)
        return x, se_reg_all

    def get_se_dim(self, x, edge_index):
        _, _, les = self.forward(x, edge_index, want_les=1)
        return les.shape[-1]

    def collect_SE(self, x, edge_index):
        _, _, les = self.forward(x, edge_index, want_les=1)
        return les


def tonp(arr):
    import numpy as np
    if type(arr) is th.Tensor:
        return arr.detach().cpu().data.numpy()
    return np.asarray(arr)
