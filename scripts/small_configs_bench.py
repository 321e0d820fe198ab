"""Step time of BASELINE.json configs[0..2] at their real shapes (shape-matched synthetic data: the datasets are
not on disk), this repo's CUDA path against the CPU oracle on the host cores.  Not the headline metric
(bench.py measures configs[3]); recorded in profiles/ for completeness."""
import os, sys, time
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_tail_generalization_b200.GNN_model.GNN_normalizations import TeacherGNN
from oracle import coldbrew_oracle as O

CFGS = {
    'cfg1 Cora NoRes SE=000': dict(n=2708, und=5278, F=1433, H=64, C=7, trick='NoResNodeNorm', se='000', ds='Cora'),
    'cfg2 Pubmed Initial SE=111': dict(n=19717, und=44324, F=500, H=256, C=3, trick='InitialBatchNorm', se='111', ds='Pubmed'),
    'cfg3 ogbn-arxiv Initial SE=100': dict(n=169343, und=1157799, F=128, H=256, C=40, trick='InitialBatchNorm', se='100', ds='ogbn-arxiv'),
}
dev = 'cuda:0'
for name, c in CFGS.items():
    torch.manual_seed(3)
    ei = O.powerlaw_graph(c['n'], c['und'], seed=0)
    kw = dict(type_trick=c['trick'], whetherHasSE=c['se'], num_layers=2, dim_hidden=c['H'], num_feats=c['F'],
              num_classes=c['C'], N_nodes=c['n'], dataset=c['ds'], res_alpha=0.1)
    ref = O.OracleTeacherGNN(O.make_args(**kw), None).train()
    a = O.make_args(**kw); a.device = dev
    model = TeacherGNN(a, None); model.load_state_dict(ref.state_dict()); model.to(dev).train()
    x = torch.randn(c['n'], c['F']); y = torch.randint(0, c['C'], (c['n'],)); mask = torch.arange(c['n']) < c['n'] // 10
    xg, eg, yg, mg = x.to(dev), ei.to(dev), y.to(dev), mask.to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3); opt_r = torch.optim.Adam(ref.parameters(), lr=1e-3)

    def step_gpu():
        opt.zero_grad(set_to_none=True)
        res = model.get_3_embs(xg, eg, mg)
        loss = F.nll_loss(F.log_softmax(res.emb4classi, 1), yg[mg])
        if model.se_reg_all is not None:
            loss = loss + 0.5 * model.se_reg_all
        loss.backward(); opt.step()

    def step_cpu():
        opt_r.zero_grad(set_to_none=True)
        O.teacher_loss(ref, x, ei, y, mask, 0.5).backward(); opt_r.step()

    for _ in range(5): step_gpu()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): step_gpu()
    torch.cuda.synchronize(); tg = (time.perf_counter() - t0) / 20
    step_cpu(); t0 = time.perf_counter()
    for _ in range(3): step_cpu()
    tc = (time.perf_counter() - t0) / 3
    E = ei.shape[1]
    print(f'{name}: N={c["n"]} E={E}: B200 {tg * 1e3:.2f} ms/step ({4 * E / tg / 1e6:.1f} M edges/s), '
          f'CPU oracle ({os.cpu_count()} cores) {tc * 1e3:.1f} ms/step ({4 * E / tc / 1e6:.1f} M edges/s)', flush=True)
