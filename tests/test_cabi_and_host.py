"""CPU suite: the C-ABI library loads and exports what include/coldbrew_b200.h declares, and the host
side of the drop-in modules (construction, state_dict keys, RNG order) matches the reference fixtures.
No kernel is launched here."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

from tests.helpers import golden_args, golden_cases, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def built_lib():
    from gnn_tail_generalization_b200 import build
    return build.build()


def _declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'coldbrew_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(cb_[a-z0-9_]+)\s*\(', text)))


def test_header_declares_what_binding_binds():
    from gnn_tail_generalization_b200 import _cabi
    assert _declared_symbols() == sorted(_cabi.SYMBOLS)


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    for name in _declared_symbols():
        assert hasattr(lib, name), name
    lib.cb_abi_version.restype = ctypes.c_int
    from gnn_tail_generalization_b200 import _cabi
    assert lib.cb_abi_version() == _cabi.ABI_VERSION == 6
    lib.cb_last_error.restype = ctypes.c_char_p
    assert lib.cb_last_error() == b''
    lib.cb_launch_count.restype = ctypes.c_int64
    assert lib.cb_launch_count() >= 0


def test_argument_validation_without_gpu(built_lib):
    """Argument checks happen before any CUDA call, so they can run here."""
    from gnn_tail_generalization_b200 import _cabi as C
    lib = C.lib()
    out = ctypes.c_void_p()
    assert lib.cb_graph_create(None, -1, 10, 0, None, ctypes.byref(out)) == -1
    assert b'negative' in lib.cb_last_error()
    assert lib.cb_graph_create(None, 5, 10, 0, None, ctypes.byref(out)) == -1
    assert lib.cb_graph_create_sliced(None, 0, 10, 4, 2, 0, None, ctypes.byref(out)) == -1
    assert lib.cb_graph_create(None, 0, 2 ** 31, 0, None, ctypes.byref(out)) == -4
    assert lib.cb_agg_forward(None, None, 4, 4, None, None, 0.0, 0, None, None, None, 4, None, 0, None) == -1
    assert lib.cb_agg_gather(None, 0, None, 4, 4, None, None, None, 4, None, 0, None) == -1
    assert lib.cb_row_scale(None, None, 3, 0, None, None) == -1
    assert lib.cb_row_scale(None, None, 0, 4, None, None) == 0          # empty input: nothing to do
    assert lib.cb_graph_destroy(None) == 0
    with pytest.raises(C.ColdBrewError) as ei:
        C.call('cb_sumsq', None, 4, None, None, 0, None)
    assert ei.value.code == -1


def test_no_cpu_fallback():
    from gnn_tail_generalization_b200 import graph, ops
    with pytest.raises(ValueError):
        graph.GraphHandle(torch.zeros(2, 3, dtype=torch.int64), 4)
    with pytest.raises(RuntimeError):
        ops.row_scale_raw(torch.ones(2, 2), torch.ones(2))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'gnn_tail_generalization_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', text, flags=re.M), f
                assert 'libcb_oracle' not in text, f


@pytest.mark.parametrize('name', golden_cases())
def test_state_dict_keys_and_shapes_match_reference(name):
    from gnn_tail_generalization_b200.GNN_model.GNN_normalizations import TeacherGNN
    from oracle.coldbrew_oracle import make_args
    z = load_golden(name)
    model = TeacherGNN(golden_args(z, make_args), None)
    want = {k[len('param/'):]: v.shape for k, v in z.items() if k.startswith('param/')}
    got = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    assert got == {k: tuple(s) for k, s in want.items()}


# seeds of tests/golden/make_golden.py: torch.manual_seed(1000 + seed) precedes x, y and the model
_SEEDS = {'nores_se000_L2': 1, 'nores_se111_L3': 2, 'nores_se100_L4': 3, 'initial_se111_L2': 4,
          'initial_se100_L3': 5, 'residual_se010_L3': 6, 'dense_concat_L2': 7, 'dense_maxpool_L2': 8,
          'dense_attention_L2': 9, 'jumping_concat_L3': 10, 'jumping_maxpool_L2': 11, 'exact_batchnorm_L2': 12,
          'exact_pairnorm_L3': 13, 'learnable_input_L2': 14, 'odd_dims_L2': 15, 'initial_se000_L2': 16,
          'featureless_se111_L2': 17, 'initial_jumping_L3': 18, 'residual_jumping_L2': 19}


@pytest.mark.parametrize('name', golden_cases())
def test_rng_order_reproduces_reference_initial_weights(name):
    """Same seed, same draw order => the replacement modules start from the reference's weights."""
    from gnn_tail_generalization_b200.GNN_model.GNN_normalizations import TeacherGNN
    from oracle.coldbrew_oracle import make_args
    z = load_golden(name)
    a = golden_args(z, make_args)
    torch.manual_seed(1000 + _SEEDS[name])
    n = z['x'].shape[0]
    x = torch.randn(n, z['x'].shape[1])
    y = torch.randint(0, a.num_classes, (n,))
    assert np.array_equal(x.numpy(), z['x']) and np.array_equal(y.numpy(), z['y'])
    model = TeacherGNN(a, None)
    for k, v in model.state_dict().items():
        assert np.array_equal(v.numpy(), z['param/' + k]), k


def test_teacher_rewrites_args_like_reference():
    from gnn_tail_generalization_b200.GNN_model.GNN_normalizations import TeacherGNN
    from oracle.coldbrew_oracle import make_args
    a = make_args(N_nodes=9, num_classes=5, dim_commonEmb=11, dim_learnable_input=4, num_feats=7,
                  type_trick='Initial', whetherHasSE='010')
    m = TeacherGNN(a, None)
    assert (a.num_classes, a.num_classes_bkup, a.num_feats, a.num_feats_bkup) == (11, 5, 4, 7)
    assert m.embs.shape == (9, 4) and m.model.model.layers_MLP[-1].out_features == 11
    assert m.model.model.layers_GCN[0].le.shape == (9, a.dim_hidden)


def test_sampling_tricks_are_rejected():
    from gnn_tail_generalization_b200.GNN_model.GCN import TricksComb
    from oracle.coldbrew_oracle import make_args
    with pytest.raises(NotImplementedError):
        TricksComb(make_args(N_nodes=4, type_trick='DropEdge'))


def test_dropped_edges_answers_any_layer():
    from gnn_tail_generalization_b200.GNN_model.drop_tricks import DropoutTrick
    from oracle.coldbrew_oracle import make_args
    ei = torch.zeros(2, 3, dtype=torch.long)
    adjs = DropoutTrick(make_args(N_nodes=4))(ei)
    assert adjs[0][0] is ei and adjs[5][0] is ei and adjs[5][1] is None


def test_golden_args_roundtrip():
    from oracle.coldbrew_oracle import make_args
    z = load_golden('initial_se111_L2')
    d = json.loads(str(z['args_json']))
    a = golden_args(z, make_args)
    assert a.type_trick == d['type_trick'] and a.TeacherGNN.whetherHasSE == [1, 1, 1]
