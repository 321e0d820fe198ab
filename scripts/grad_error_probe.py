"""Where does the gradient error at ogbn-arxiv shape come from?  Max |grad - fp64 oracle| / max |grad| per parameter
under the A/B switches of the CUDA path (dense backend, backward fusion, weight-gradient chain length)."""
import os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_tail_generalization_b200 import ops
from gnn_tail_generalization_b200.GNN_model.GNN_normalizations import TeacherGNN
from oracle import coldbrew_oracle as O

c = dict(n=169343, und=1157799, F=128, H=256, C=40, trick='InitialBatchNorm', se='100', ds='ogbn-arxiv')
torch.manual_seed(3)
ei = O.powerlaw_graph(c['n'], c['und'], seed=0)
kw = dict(type_trick=c['trick'], whetherHasSE=c['se'], num_layers=2, dim_hidden=c['H'], num_feats=c['F'],
          num_classes=c['C'], N_nodes=c['n'], dataset=c['ds'], res_alpha=0.1)
ref = O.OracleTeacherGNN(O.make_args(**kw), None)
x = torch.randn(c['n'], c['F'], generator=torch.Generator().manual_seed(1))
y = torch.randint(0, c['C'], (c['n'],), generator=torch.Generator().manual_seed(2))
mask = torch.zeros(c['n'], dtype=torch.bool); mask[: c['n'] // 10] = True
ref64 = O.OracleTeacherGNN(O.make_args(**kw), None); ref64.load_state_dict(ref.state_dict()); ref64.double().train()
O.teacher_loss(ref64, x.double(), ei, y, mask, 0.5).backward()
g64 = {k: p.grad for k, p in ref64.named_parameters() if p.grad is not None}


def run(tag):
    a = O.make_args(**kw); a.device = 'cuda'
    m = TeacherGNN(a, None); m.load_state_dict(ref.state_dict()); m.to('cuda').train()
    res = m.get_3_embs(x.cuda(), ei.cuda(), mask.cuda())
    F.nll_loss(F.log_softmax(res.emb4classi, 1), y.cuda()[mask.cuda()]).backward()
    out = {k.split('model.model.')[-1]: f'{float((p.grad.cpu().double() - g64[k]).abs().max() / g64[k].abs().max()):.1e}'
           for k, p in m.named_parameters() if k in g64}
    print(tag, out, flush=True)


run('default          ')
ops.set_dense_backend('cublas'); run('cublas GEMMs     '); ops.set_dense_backend('tcgen05')
ops.set_backward_fusion(False); run('no bwd fusion    '); ops.set_backward_fusion(True)
