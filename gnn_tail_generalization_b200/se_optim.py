"""Fused optimizer step of the Structural-Embedding tables.

The reference trains ``GCNConv.le`` (GNN_model/GCN.py:181-182) like every other parameter: autograd adds the
gradient of the additive term (GCN.py:231) and of the regulariser ``se_reg * ||E||_F``
(GCN.py:232, trainer_node_classification.py:393-394) into ``le.grad`` and ``torch.optim.Adam(lr, weight_decay)``
(trainer_node_classification.py:310, 428-430) walks the table several more times.  At BASELINE.json configs[1]
(SE = 111) and configs[4] the tables and their Adam moments are the largest thing in HBM, so here one kernel
(``cb_se_adam_step``) does all of it in a single pass:

    g = dL/dh + se_reg * E / ||E||_F + weight_decay * E ;  Adam(m, v) ;  E -= step ;  shadow = bf16(E)

* ``dL/dh`` is the gradient arriving at the layer's transform output -- the tensor the transposed aggregation
  produced -- handed over by reference (``ops.GradSlot``), never copied or accumulated by autograd;
* ``||E||_F`` is this step's forward value (``cb_sumsq`` + all-reduce over the ranks that shard the table);
* for a bf16 forward the table is kept as an fp32 master (the ``le`` parameter itself, so ``state_dict`` is unchanged)
  plus a bf16 shadow that the transform's epilogue adds.

Row-sharded tables (node-sliced graphs) need nothing else: every rank steps its own rows.
"""
import torch

from . import _cabi as C
from . import ops


class _SEState:
    def __init__(self, conv, shadow_dtype):
        self.conv = conv
        self.master = conv.le.data
        if self.master.dtype != torch.float32 or not self.master.is_cuda:
            raise ValueError('FusedSEAdam needs fp32 SE tables on a CUDA device (move the model first)')
        self.master = self.master.contiguous()
        conv.le.data = self.master
        conv.le.requires_grad_(False)          # stepped here, not by autograd + torch.optim
        self.m = torch.zeros_like(self.master)
        self.v = torch.zeros_like(self.master)
        self.shadow = ops.to_bf16_raw(self.master) if shadow_dtype == torch.bfloat16 else None
        self.slot = ops.GradSlot()
        self.sumsq = None

    def operand(self, dtype):
        """What the transform's epilogue adds: the fp32 table, or its bf16 shadow."""
        if dtype == torch.float32:
            return self.master
        if self.shadow is None or self.shadow.dtype != dtype:
            raise RuntimeError(f'this SE table has no {dtype} shadow (create FusedSEAdam with shadow_dtype={dtype})')
        return self.shadow

    def norm(self, graph):
        """||E||_F over every rank's rows; keeps sum(E^2) on the device for the step kernel."""
        ss = ops.sumsq_raw(self.master)
        if graph is not None:
            ss = graph.allreduce_sum(ss)
        self.sumsq = ss
        return ss.sqrt().reshape(())


class FusedSEAdam:
    """Adam(lr, betas, eps, weight_decay) + the ``se_reg * ||E||_F`` gradient for every SE table of a TeacherGNN.

    Usage (what bench.py does):
        se_opt = FusedSEAdam(teacher, lr=..., weight_decay=..., se_reg=args.se_reg, shadow_dtype=torch.bfloat16)
        opt = torch.optim.Adam(se_opt.other_parameters(), lr=..., weight_decay=...)
        loss = nll + args.se_reg * teacher.se_reg_all      # the value enters the loss, its gradient is applied here
        loss.backward(); opt.step(); se_opt.step()
    """

    def __init__(self, teacher, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, se_reg=0.0, shadow_dtype=None):
        self.teacher = teacher
        self.lr, self.betas, self.eps, self.weight_decay, self.se_reg = float(lr), betas, float(eps), \
            float(weight_decay), float(se_reg)
        self.t = 0
        self.states = []
        for conv in teacher.model.model.layers_GCN:
            if getattr(conv, 'whetherHasSE', False):
                st = _SEState(conv, shadow_dtype)
                conv.se_fused = st
                self.states.append(st)

    def other_parameters(self):
        """Every parameter this optimizer does not step (hand these to torch.optim)."""
        mine = {id(st.conv.le) for st in self.states}
        return [p for p in self.teacher.parameters() if id(p) not in mine]

    def zero_grad(self):
        for st in self.states:
            st.slot.grad = None

    def step(self):
        self.t += 1
        for st in self.states:
            g = st.slot.grad
            if g is not None:
                if g.dim() == 3:
                    g = g.reshape(g.shape[0], -1)
                if g.shape != st.master.shape or not g.is_contiguous():
                    raise RuntimeError('FusedSEAdam: gradient layout does not match the table')
            dev = st.master.device
            n = st.master.numel()
            alg = n * (24 + (2 if st.shadow is not None else 0) + (0 if g is None else g.element_size()))
            with torch.cuda.device(dev), ops._Timed('se_adam_step', alg, dev):
                C.call('cb_se_adam_step', C.ptr(st.master), C.ptr(g),
                       C.CB_BF16 if (g is not None and g.dtype == torch.bfloat16) else C.CB_F32, C.ptr(st.m), C.ptr(st.v),
                       C.ptr(st.shadow), st.master.numel(), self.lr, self.betas[0], self.betas[1], self.eps,
                       self.weight_decay, self.t, C.ptr(st.sumsq) if self.se_reg != 0.0 else None, self.se_reg,
                       C.stream_ptr(dev))
            st.slot.grad = None

    @property
    def bytes_per_step(self):
        """Algorithmic HBM bytes of one step() (E, m, v read + written, gradient read, shadow written)."""
        total = 0
        for st in self.states:
            n = st.master.numel()
            total += n * (24 + (2 if st.shadow is not None else 0) + (2 if st.shadow is not None else 4))
        return total
