"""Shared helpers for the test-suite (fixtures loading, args construction)."""
import glob
import json
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def golden_cases():
    return sorted(os.path.basename(p)[4:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, 'ref_*.npz'))
                  if not p.endswith('ref_kat_toy.npz'))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, f'ref_{name}.npz'), allow_pickle=False)
    return {k: z[k] for k in z.files}


def golden_args(z, make_args):
    """Rebuild the user-facing args of a golden case.  TeacherGNN.__init__ rewrote num_classes/num_feats in
    place before they were stored (GNN_normalizations.py:13-22), so restore the *_bkup values."""
    d = json.loads(str(z['args_json']))
    if 'num_classes_bkup' in d:
        d['num_classes'] = d.pop('num_classes_bkup')
    if 'num_feats_bkup' in d:
        d['num_feats'] = d.pop('num_feats_bkup')
    return make_args(**d)


def load_params(module, z, device=None):
    sd = {k[len('param/'):]: torch.from_numpy(v.copy()) for k, v in z.items() if k.startswith('param/')}
    if device is not None:
        sd = {k: v.to(device) for k, v in sd.items()}
    missing = module.load_state_dict(sd, strict=True)
    return missing
