"""Whole-step CUDA graph capture for the launch-bound configurations.

At Cora / Pubmed size (BASELINE configs[0..1]) a TeacherGNN training step is 60-90 kernel launches of a few
microseconds each: the step time is the time Python needs to issue them.  The reference pays the same price
(plus DGL's dispatch, GNN_model/GCN.py:184-258 once per layer per pass).  Every kernel of this library launches
on the current stream with caller-owned buffers, never synchronises and never allocates behind the caller's back,
so a whole step -- forward, loss, backward, optimizer -- can be captured once and replayed as ONE graph launch.

    step = GraphedTrainStep(model, optimizer, loss_fn, x, edge_index, mask, y)
    for epoch in range(n):
        loss = step()              # or step(x_new, y_new): copied into the captured input buffers

Constraints (those of CUDA graphs): static shapes, the optimizer must be capturable (``torch.optim.Adam(...,
capturable=True)``), ``loss_fn(result, y, model)`` must not synchronise, and ``mask`` must be an index tensor (a
boolean mask makes ``emb[mask]`` call ``nonzero``, which synchronises).  The graph (``edge_index``) and the mask are fixed
at capture time, which is how the reference's trainer uses them (trainer_node_classification.py:386-394: same
graph and train mask every epoch).
"""
import torch


class GraphedTrainStep:
    def __init__(self, model, optimizer, loss_fn, x, edge_index, mask, y, warmup=3):
        if not x.is_cuda:
            raise ValueError('GraphedTrainStep needs CUDA tensors')
        self.model, self.optimizer, self.loss_fn = model, optimizer, loss_fn
        self.x, self.y = x.clone(), y.clone()
        self.edge_index, self.mask = edge_index, mask
        self.loss = None
        dev = x.device
        # warm-up on a side stream: builds the graph handle, configures the kernels (cudaFuncSetAttribute),
        # fills the allocator pools and initialises the optimizer state -- none of which may happen in a capture
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._step()

    def _step(self):
        self.optimizer.zero_grad(set_to_none=True)
        res = self.model.get_3_embs(self.x, self.edge_index, self.mask)
        loss = self.loss_fn(res, self.y, self.model)
        loss.backward()
        self.optimizer.step()
        return loss.detach()

    def __call__(self, x=None, y=None):
        if x is not None:
            self.x.copy_(x, non_blocking=True)
        if y is not None:
            self.y.copy_(y, non_blocking=True)
        self.graph.replay()
        return self.loss
