import sys, torch, torch.nn.functional as F
sys.path.insert(0, '/root/repo')
from oracle import coldbrew_oracle as O
from gnn_tail_generalization_b200 import ops
from gnn_tail_generalization_b200.GNN_model.GNN_normalizations import TeacherGNN
DEV = 'cuda:0'
c = dict(n=19717, und=44324, F=500, H=256, C=3, trick='InitialBatchNorm', se='111', ds='Pubmed')
for backend in ('cublas', 'tcgen05'):
    ops.set_dense_backend(backend)
    torch.manual_seed(3)
    ei = O.powerlaw_graph(c['n'], c['und'], seed=0)
    kw = dict(type_trick=c['trick'], whetherHasSE=c['se'], num_layers=2, dim_hidden=c['H'], num_feats=c['F'],
              num_classes=c['C'], N_nodes=c['n'], dataset=c['ds'], res_alpha=0.1)
    ref = O.OracleTeacherGNN(O.make_args(**kw), None)
    a = O.make_args(**kw); a.device = DEV
    model = TeacherGNN(a, None); model.load_state_dict(ref.state_dict(), strict=True); model.to(DEV)
    x = torch.randn(c['n'], c['F'], generator=torch.Generator().manual_seed(1))
    y = torch.randint(0, c['C'], (c['n'],), generator=torch.Generator().manual_seed(2))
    mask = torch.zeros(c['n'], dtype=torch.bool); mask[: c['n'] // 10] = True
    ref.train(); model.train()
    r = ref.get_3_embs(x, ei, mask)
    nll_r = F.nll_loss(F.log_softmax(r.emb4classi, 1), y[mask]); se_r = ref.se_reg_all
    res = model.get_3_embs(x.to(DEV), ei.to(DEV), mask.to(DEV))
    nll_g = F.nll_loss(F.log_softmax(res.emb4classi, 1), y.to(DEV)[mask.to(DEV)]); se_g = model.se_reg_all
    print(backend, 'nll ref', float(nll_r), 'ours', float(nll_g), '| se ref', float(se_r), 'ours', float(se_g),
          '| max logit diff', float((res.emb4classi_full.cpu() - r.emb4classi_full).abs().max()),
          'logit scale', float(r.emb4classi_full.abs().max()), 'dropout', a.dropout)
    lr = O.teacher_loss(ref, x, ei, y, mask, 0.5)
    print('   teacher_loss', float(lr), ' ours', float(nll_g + 0.5 * se_g))
