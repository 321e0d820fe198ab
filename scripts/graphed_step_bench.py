"""Eager vs CUDA-graph-replayed training step at BASELINE configs[0..2] shapes (graphs.GraphedTrainStep)."""
import os, sys, time
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_tail_generalization_b200.GNN_model.GNN_normalizations import TeacherGNN
from gnn_tail_generalization_b200.graphs import GraphedTrainStep
from oracle import coldbrew_oracle as O   # args helper + graph generator only

CFGS = {
    'cfg1 Cora NoRes SE=000': dict(n=2708, und=5278, F=1433, H=64, C=7, trick='NoResNodeNorm', se='000', ds='Cora'),
    'cfg2 Pubmed Initial SE=111': dict(n=19717, und=44324, F=500, H=256, C=3, trick='InitialBatchNorm', se='111', ds='Pubmed'),
    'cfg3 ogbn-arxiv Initial SE=100': dict(n=169343, und=1157799, F=128, H=256, C=40, trick='InitialBatchNorm', se='100', ds='ogbn-arxiv'),
}
dev = 'cuda:0'
for name, c in CFGS.items():
    ei = O.powerlaw_graph(c['n'], c['und'], seed=0).to(dev)
    kw = dict(type_trick=c['trick'], whetherHasSE=c['se'], num_layers=2, dim_hidden=c['H'], num_feats=c['F'],
              num_classes=c['C'], N_nodes=c['n'], dataset=c['ds'], res_alpha=0.1)
    x = torch.randn(c['n'], c['F'], device=dev); y = torch.randint(0, c['C'], (c['n'],), device=dev)
    mask = torch.arange(c['n'] // 10, device=dev)      # index mask (a boolean mask would synchronise in emb[mask])

    def loss_fn(res, yy, model):
        loss = F.nll_loss(F.log_softmax(res.emb4classi, 1), yy[mask])
        return loss + 0.5 * model.se_reg_all if model.se_reg_all is not None else loss

    def make():
        torch.manual_seed(3)
        a = O.make_args(**kw); a.device = dev
        m = TeacherGNN(a, None).to(dev).train()
        return m, torch.optim.Adam(m.parameters(), lr=1e-3, capturable=True)

    m, opt = make()
    def eager():
        opt.zero_grad(set_to_none=True)
        loss_fn(m.get_3_embs(x, ei, mask), y, m).backward(); opt.step()
    for _ in range(5): eager()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(50): eager()
    torch.cuda.synchronize(); te = (time.perf_counter() - t0) / 50
    m2, opt2 = make()
    step = GraphedTrainStep(m2, opt2, loss_fn, x, ei, mask, y)
    for _ in range(5): step()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(50): step()
    torch.cuda.synchronize(); tg = (time.perf_counter() - t0) / 50
    E = ei.shape[1]
    print(f'{name}: eager {te * 1e3:.3f} ms/step, graph replay {tg * 1e3:.3f} ms/step ({te / tg:.1f}x, '
          f'{4 * E / tg / 1e6:.0f} M edges/s)', flush=True)
