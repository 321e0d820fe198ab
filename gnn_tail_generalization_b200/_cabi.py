"""ctypes binding of libcoldbrew_b200.so (the C ABI declared in include/coldbrew_b200.h).

The product path has no CPU fallback: if the library is missing or a call fails, this module raises.
PyTorch is used only for device memory (``tensor.data_ptr()``) and the current CUDA stream.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# CB_LIB: another build of the same ABI (A/B measurements of a kernel change); default: the in-tree library
LIB_PATH = os.environ.get('CB_LIB') or os.path.join(_HERE, 'libcoldbrew_b200.so')

CB_OK = 0
ABI_VERSION = 6
CB_ACT_NONE, CB_ACT_RELU = 0, 1
CB_BY_DST, CB_BY_SRC = 0, 1
CB_F32, CB_BF16 = 0, 1
PREP_DEGREES, PREP_SYMMETRIZE, PREP_PARTIAL_SORTED_IDX, PREP_DEGREE_STATS, PREP_SORT_IDX_BY_VALUE, PREP_MASK_FROM_IDX, \
    PREP_DROP_EDGES = range(7)
CB_PANEL_SHIFT = 7      # source-panel blocks are 128 rows (include/coldbrew_b200.h)

Q_NUM_NODES, Q_NUM_EDGES, Q_ROW_BEGIN, Q_ROW_END, Q_HAS_ZERO_IN_DEG, Q_HUB_CHUNK, Q_SRC_PANELS = 0, 1, 2, 3, 4, 5, 6
Q_DST_ROWPTR_EXP, Q_SRC_ROWPTR_EXP = 14, 25
Q_DST_ROWPTR, Q_DST_COL, Q_DST_PERM, Q_DST_NUM_HUB_CHUNKS = 10, 11, 12, 13
Q_SRC_ROWPTR, Q_SRC_COL, Q_SRC_PERM, Q_SRC_NUM_HUB_CHUNKS, Q_SRC_NUM_EDGES = 20, 21, 22, 23, 24
Q_DIN_INV_SQRT, Q_DOUT_INV_SQRT, Q_IN_DEGREE, Q_OUT_DEGREE = 30, 31, 32, 33

CB_PEER_HANDLE_BYTES, CB_MAX_PEERS = 64, 7


class PeerPush(ctypes.Structure):
    """cb_peer_push_t: where the rows of a kernel output go besides its local `out`."""
    _fields_ = [('n_peers', ctypes.c_int32), ('max_ctas', ctypes.c_int32),
                ('peer', ctypes.c_void_p * CB_MAX_PEERS), ('need', ctypes.c_void_p),
                ('row0', ctypes.c_int64), ('ld', ctypes.c_int64), ('row_live', ctypes.c_void_p),
                ('tile_first', ctypes.c_int32), ('tile_step', ctypes.c_int32)]


# every symbol include/coldbrew_b200.h declares: name -> (restype, argtypes)
_vp, _i64, _int, _dbl = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_double
SYMBOLS = {
    'cb_last_error': (ctypes.c_char_p, []),
    'cb_abi_version': (_int, []),
    'cb_graph_create': (_int, [_vp, _i64, _i64, _int, _vp, ctypes.POINTER(_vp)]),
    'cb_graph_create_sliced': (_int, [_vp, _i64, _i64, _i64, _i64, _int, _vp, ctypes.POINTER(_vp)]),
    'cb_graph_create_local': (_int, [_vp, _i64, _vp, _i64, _i64, _i64, _i64, _int, _vp, ctypes.POINTER(_vp)]),
    'cb_graph_create_panelled': (_int, [_vp, _i64, _vp, _i64, _i64, _i64, _i64, _int, _int, _int, _vp,
                                        ctypes.POINTER(_vp)]),
    'cb_graph_destroy': (_int, [_vp]),
    'cb_graph_query': (_int, [_vp, _int, _vp]),
    'cb_graph_workspace_bytes': (_i64, [_vp, _int, _i64]),
    'cb_agg_forward': (_int, [_vp, _vp, _i64, _i64, _vp, _vp, _dbl, _int, _vp, _vp, _vp, _i64, _vp, _i64, _vp]),
    'cb_agg_gather': (_int, [_vp, _int, _vp, _i64, _i64, _vp, _vp, _vp, _i64, _vp, _i64, _vp]),
    'cb_agg_forward_bf16': (_int, [_vp, _vp, _i64, _i64, _vp, _vp, _dbl, _int, _vp, _vp, _vp, _i64, _vp, _i64, _vp]),
    'cb_agg_gather_bf16': (_int, [_vp, _int, _vp, _i64, _i64, _vp, _vp, _vp, _i64, _vp, _i64, _vp]),
    'cb_agg_forward_pass': (_int, [_vp, _int, _vp, _i64, _i64, _vp, _vp, _dbl, _int, _vp, _vp, _vp, _i64, _int, _vp, _vp,
                                   _i64, _vp]),
    'cb_agg_gather_pass': (_int, [_vp, _int, _int, _vp, _i64, _i64, _vp, _vp, _i64, _int, _vp, _vp, _i64, _vp]),
    'cb_graph_sort_edge_values': (_int, [_vp, _int, _vp, _i64, _vp, _vp]),
    'cb_agg_gather_weighted': (_int, [_vp, _int, _int, _vp, _i64, _i64, _vp, _vp, _vp, _i64, _vp, _i64, _vp]),
    'cb_agg_edge_dot': (_int, [_vp, _int, _int, _vp, _i64, _vp, _i64, _i64, _vp, _vp]),
    'cb_agg_propagate': (_int, [_vp, _int, _vp, _i64, _vp, _vp, _dbl, _dbl, _int, _dbl, _dbl, _vp, _vp, _vp, _vp, _i64, _vp]),
    'cb_graph_live_workspace_bytes': (_i64, [_vp, _int]),
    'cb_graph_compact_live': (_int, [_vp, _int, _vp, _vp, _i64, _vp]),
    'cb_agg_gather_compacted': (_int, [_vp, _int, _int, _vp, _i64, _i64, _vp, _vp, _vp, _i64, _vp, _i64, _vp]),
    'cb_prep_workspace_bytes': (_i64, [_i64, _i64]),
    'cb_agg_backward_prep': (_int, [_vp, _vp, _vp, _i64, _vp, _vp, _int, _int, _dbl, _vp, _vp, _vp, _int, _vp,
                                    _i64, _vp]),
    'cb_agg_backward_prep_bf16': (_int, [_vp, _vp, _vp, _i64, _vp, _vp, _int, _int, _dbl, _vp, _vp, _vp, _int, _vp,
                                         _i64, _vp]),
    'cb_agg_backward_prep_ex': (_int, [_vp, _int, _vp, _vp, _i64, _vp, _vp, _int, _int, _dbl, _vp, _dbl, _vp, _vp, _vp,
                                       _int, _vp, _vp, _i64, _vp]),
    'cb_se_adam_step': (_int, [_vp, _vp, _int, _vp, _vp, _vp, _i64, _dbl, _dbl, _dbl, _dbl, _dbl, _i64, _vp, _dbl, _vp]),
    'cb_to_bf16': (_int, [_vp, _i64, _vp, _vp]),
    'cb_row_scale': (_int, [_vp, _vp, _i64, _i64, _vp, _vp]),
    'cb_sumsq_workspace_bytes': (_i64, []),
    'cb_sumsq': (_int, [_vp, _i64, _vp, _vp, _i64, _vp]),
    'cb_gemm_split_weight': (_int, [_vp, _i64, _i64, _int, _vp, _vp, _vp]),
    'cb_gemm_rows_supported': (_int, [_i64, _i64, _i64]),
    'cb_gemm_rows': (_int, [_vp, _i64, _i64, _i64, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _int, _vp, _i64, _vp, _vp,
                            _i64, _vp, _vp]),
    'cb_peer_alloc': (_int, [_i64, ctypes.POINTER(_vp), _vp]),
    'cb_peer_open': (_int, [_vp, ctypes.POINTER(_vp)]),
    'cb_peer_close': (_int, [_vp]),
    'cb_peer_free': (_int, [_vp]),
    'cb_gemm_rows_grad_workspace_bytes': (_i64, [_i64, _i64]),
    'cb_gemm_rows_masked': (_int, [_int, _vp, _i64, _i64, _i64, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _int, _vp, _i64, _vp, _vp,
                                   _i64, _vp, _i64, _vp, _vp]),
    'cb_gemm_rows_grad': (_int, [_vp, _i64, _i64, _i64, _vp, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _i64, _int, _dbl,
                                 _vp, _i64, _int, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp]),
    'cb_row_any_nonzero': (_int, [_vp, _int, _i64, _i64, _i64, _vp, _vp]),
    'cb_gemm_tn_supported': (_int, [_i64, _i64, _i64]),
    'cb_gemm_tn_workspace_bytes': (_i64, [_i64, _i64, _i64]),
    'cb_gemm_tn': (_int, [_vp, _i64, _vp, _i64, _i64, _i64, _i64, _vp, _int, _vp, _i64, _vp, _i64, _vp]),
    'cb_gemm_weight_to_bf16': (_int, [_vp, _i64, _i64, _int, _vp, _vp]),
    'cb_gemm_rows_supported_bf16': (_int, [_i64, _i64, _i64]),
    'cb_gemm_rows_bf16': (_int, [_vp, _i64, _i64, _i64, _vp, _i64, _vp, _vp, _vp, _i64, _int, _vp, _i64, _vp, _vp, _i64,
                                 _vp, _vp]),
    'cb_gemm_rows_grad_bf16': (_int, [_vp, _i64, _i64, _i64, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _i64, _int, _dbl,
                                      _vp, _i64, _int, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp]),
    'cb_gemm_tn_supported_bf16': (_int, [_i64, _i64, _i64]),
    'cb_gemm_tn_workspace_bytes_bf16': (_i64, [_i64, _i64, _i64]),
    'cb_gemm_tn_bf16': (_int, [_vp, _i64, _vp, _i64, _i64, _i64, _i64, _vp, _i64, _vp, _i64, _vp]),
    'cb_prep_graph_workspace_bytes': (_i64, [_int, _i64]),
    'cb_prep_degrees': (_int, [_vp, _i64, _i64, _vp, _vp, _vp, _i64, _vp]),
    'cb_prep_symmetrize': (_int, [_vp, _i64, _vp, ctypes.POINTER(_i64), _vp, _i64, _vp]),
    'cb_prep_partial_sorted_idx': (_int, [_vp, _i64, _int, _int, _vp, ctypes.POINTER(_i64), _vp, _i64, _vp]),
    'cb_prep_degree_stats': (_int, [_vp, _i64, ctypes.POINTER(_dbl), _vp, _i64, _vp]),
    'cb_prep_sort_idx_by_value': (_int, [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _vp]),
    'cb_prep_mask_from_idx': (_int, [_vp, _i64, _i64, _vp, _vp, _i64, _vp]),
    'cb_prep_drop_edges': (_int, [_vp, _i64, _vp, _i64, _vp, ctypes.POINTER(_i64), _vp, _i64, _vp]),
    'cb_topk_merge': (_int, [_vp, _i64, _i64, _i64, _i64, _int, _vp, _vp, _int, _vp]),
    'cb_topk_softmax_mix': (_int, [_vp, _vp, _i64, _int, _vp, _i64, _i64, _vp, _vp]),
    'cb_launch_count': (_i64, []),
}


class ColdBrewError(RuntimeError):
    """A C-ABI call returned a negative status."""

    def __init__(self, fn, code, text):
        super().__init__(f'{fn} failed with code {code}: {text}')
        self.code = code


_lib = None


def lib():
    """The loaded library.  Raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f'{LIB_PATH} is missing: build it with `python -m gnn_tail_generalization_b200.build` '
                f'(nvcc, sm_100a).  There is no CPU or PyTorch fallback for this path.')
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)          # AttributeError if the symbol is not exported
            fn.restype, fn.argtypes = res, args
        if handle.cb_abi_version() != ABI_VERSION:
            raise ImportError(f'{LIB_PATH}: ABI version {handle.cb_abi_version()} != {ABI_VERSION}, rebuild')
        _lib = handle
    return _lib


def check(fn_name, rc):
    if rc != CB_OK:
        raise ColdBrewError(fn_name, rc, lib().cb_last_error().decode('utf-8', 'replace'))


def call(fn_name, *args):
    check(fn_name, getattr(lib(), fn_name)(*args))


def ptr(t):
    """Device address of a tensor, or NULL for None."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def launch_count():
    return int(lib().cb_launch_count())
