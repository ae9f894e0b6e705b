"""ctypes binding of libgrand_b200.so (include/grand_b200.h).

The product path has no fallback: if the library is missing or a symbol is absent, importing this module's
`lib()` raises.  PyTorch is used around it only for device memory (`tensor.data_ptr()`) and streams.
"""
import ctypes as C
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('GRAND_LIB') or os.path.join(HERE, 'libgrand_b200.so')   # GRAND_LIB: A/B builds

GRAND_TILE = 32
GRAND_SEG_EDGES = int(os.environ.get('GRAND_SEG_EDGES', '512'))   # rows longer than this are cut into segments
GRAND_MAX_HEADS = 8
GRAND_MAX_TERMS = 8
ABI_VERSION = 3

SCORE_TYPES = {'scaled_dot': 0, 'cosine_sim': 1, 'pearson': 2, 'exp_kernel': 3, 'exp_kernel2': 4, 'gat': 5}
NORM_SOFTMAX, NORM_SQUAREPLUS = 0, 1
METHODS = {'euler': 0, 'midpoint': 1, 'rk4': 2}

c_float_p = C.POINTER(C.c_float)
c_int32_p = C.POINTER(C.c_int32)


class GrandGraph(C.Structure):
  _fields_ = [('n_rows', C.c_int32), ('n_cols', C.c_int32), ('n_edges', C.c_int32), ('n_tiles', C.c_int32),
              ('n_split_rows', C.c_int32), ('n_split_tiles', C.c_int32),
              ('rowptr', C.c_void_p), ('col', C.c_void_p), ('erow', C.c_void_p), ('eid', C.c_void_p),
              ('win_hdr', C.c_void_p), ('win_rec', C.c_void_p),
              ('split_tile0', C.c_void_p), ('colptr', C.c_void_p), ('csc_pos', C.c_void_p),
              ('n_lunits', C.c_int32), ('n_munits', C.c_int32), ('n_lrows', C.c_int32),
              ('seg_desc', C.c_void_p), ('lrow_slot0', C.c_void_p), ('lrow_row', C.c_void_p)]


class GrandAttention(C.Structure):
  _fields_ = [('d', C.c_int32), ('ad', C.c_int32), ('heads', C.c_int32), ('score_type', C.c_int32),
              ('norm_type', C.c_int32), ('norm_idx', C.c_int32),
              ('wq', C.c_void_p), ('bq', C.c_void_p), ('wk', C.c_void_p), ('bk', C.c_void_p),
              ('kernel_scalars', C.c_void_p), ('edge_weight_csr', C.c_void_p)]


GRAND_MAX_PEERS = 8


class GrandPeers(C.Structure):
  _fields_ = [('world', C.c_int32), ('rank', C.c_int32), ('arena', C.c_void_p * GRAND_MAX_PEERS),
              ('flags_offset', C.c_int64), ('push_ptr', C.c_void_p), ('push_dst', C.c_void_p),
              ('push_src', C.c_void_p), ('n_push', C.c_int32)]


class GrandEpilogue(C.Structure):
  _fields_ = [('alpha_train', C.c_void_p), ('beta_train', C.c_void_p), ('x0', C.c_void_p),
              ('no_alpha_sigmoid', C.c_int32), ('n_terms', C.c_int32),
              ('f_out', C.c_void_p), ('y_out', C.c_void_p), ('coef_f', C.c_float),
              ('coef', C.c_float * GRAND_MAX_TERMS), ('term', C.c_void_p * GRAND_MAX_TERMS),
              ('peers', C.POINTER(GrandPeers)), ('coef_dev', C.c_void_p)]


GRAND_DOPRI5_COEFS = 72


class GrandDopri5State(C.Structure):
  _fields_ = [('t', C.c_double), ('t_prev', C.c_double), ('dt', C.c_double), ('t_end', C.c_double),
              ('acc', C.c_double * 3), ('h0', C.c_double), ('rtol', C.c_float), ('atol', C.c_float),
              ('numel', C.c_int32), ('attempts', C.c_int32), ('accepted', C.c_int32), ('nfe', C.c_int32),
              ('max_attempts', C.c_int32), ('accept', C.c_int32), ('have_interp', C.c_int32), ('status', C.c_int32),
              ('coef', C.c_float * GRAND_DOPRI5_COEFS)]


class GrandWorkspace(C.Structure):
  _fields_ = [('q', C.c_void_p), ('k', C.c_void_p), ('scores', C.c_void_p), ('seg_stats', C.c_void_p),
              ('split_stats', C.c_void_p), ('row_stats', C.c_void_p), ('split_partial', C.c_void_p), ('split_counter', C.c_void_p),
              ('gmax', C.c_void_p), ('lrow_stats', C.c_void_p), ('lpartial', C.c_void_p), ('lcounter', C.c_void_p),
              ('coop_sync', C.c_void_p)]


# every symbol include/grand_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SIGNATURES = {
  'grand_abi_version': (C.c_int, []),
  'grand_last_error': (C.c_char_p, []),
  'grand_launch_count': (C.c_int64, []),
  'grand_workspace_sizes': (C.c_int, [C.POINTER(GrandGraph), C.POINTER(GrandAttention)] + [C.POINTER(C.c_int64)] * 10),
  'grand_project_qk': (C.c_int, [_P, C.c_int32, C.c_int32, C.POINTER(GrandAttention), _P, _P, _P]),
  'grand_plan_create': (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
  'grand_plan_describe': (C.c_int, [C.c_void_p, C.POINTER(GrandGraph), C.POINTER(C.c_int32), C.POINTER(C.c_int64),
                                    C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
  'grand_plan_image': (C.c_void_p, [C.c_void_p]),
  'grand_plan_bind': (C.c_int, [C.c_void_p, _P, C.POINTER(GrandGraph)]),
  'grand_plan_destroy': (None, [C.c_void_p]),
  'grand_linear': (C.c_int, [_P, C.c_int32, C.c_int32, _P, _P, C.c_int32, _P, C.c_int32, _P]),
  'grand_set_projection_impl': (C.c_int, [C.c_int32]),
  'grand_rhs_attention': (C.c_int, [C.POINTER(GrandGraph), C.POINTER(GrandAttention), _P, C.POINTER(GrandWorkspace),
                                     C.POINTER(GrandEpilogue), _P]),
  'grand_rhs_attention_edges': (C.c_int, [C.POINTER(GrandGraph), C.POINTER(GrandAttention), _P,
                                           C.POINTER(GrandWorkspace), C.POINTER(GrandEpilogue), _P]),
  'grand_rhs_attention_backward': (C.c_int, [C.POINTER(GrandGraph), C.POINTER(GrandAttention), _P, _P,
                                              C.POINTER(GrandWorkspace), _P, C.c_int32, _P, _P, _P, _P, _P]),
  'grand_edge_dot': (C.c_int, [C.POINTER(GrandGraph), C.c_int32, _P, _P, _P, C.c_int32, _P, _P]),
  'grand_attention_weights_backward': (C.c_int, [C.POINTER(GrandGraph), C.POINTER(GrandAttention), _P,
                                                  C.POINTER(GrandWorkspace), _P, _P, _P, _P, _P]),
  'grand_rhs_laplacian': (C.c_int, [C.POINTER(GrandGraph), C.c_int32, _P, _P, C.POINTER(GrandWorkspace),
                                     C.POINTER(GrandEpilogue), _P]),
  'grand_attention_weights': (C.c_int, [C.POINTER(GrandGraph), C.POINTER(GrandAttention), _P,
                                         C.POINTER(GrandWorkspace), _P, _P, _P, _P]),
  'grand_edge_values_to_csr': (C.c_int, [C.POINTER(GrandGraph), _P, C.c_int32, _P, _P]),
  'grand_integrate_fixed': (C.c_int, [C.POINTER(GrandGraph), C.POINTER(GrandAttention), _P, C.c_int32,
                                       C.POINTER(GrandWorkspace), C.POINTER(GrandEpilogue), C.c_int32,
                                       c_float_p, C.c_int32, _P, _P, _P]),
  'grand_pack_rows': (C.c_int, [_P, _P, C.c_int64, C.c_int32, _P, _P]),
  'grand_lincomb': (C.c_int, [C.c_int64, C.c_int32, C.POINTER(C.c_void_p), c_float_p, _P, _P]),
  'grand_error_sumsq': (C.c_int, [C.c_int64, C.c_int32, C.POINTER(C.c_void_p), c_float_p, _P, _P, C.c_float,
                                   C.c_float, _P, _P]),
  'grand_peer_alloc': (C.c_int, [C.c_int64, C.POINTER(C.c_void_p)]),
  'grand_peer_free': (C.c_int, [_P]),
  'grand_peer_export': (C.c_int, [_P, _P]),
  'grand_peer_open': (C.c_int, [_P, C.POINTER(C.c_void_p)]),
  'grand_peer_close': (C.c_int, [_P]),
  'grand_peer_push_rows': (C.c_int, [C.POINTER(GrandPeers), _P, C.c_int32, C.c_int32, _P]),
  'grand_peer_barrier': (C.c_int, [C.POINTER(GrandPeers), _P]),
  'grand_peer_error': (C.c_int, [C.POINTER(GrandPeers), C.POINTER(C.c_int32)]),
  'grand_integrate_fixed_sharded': (C.c_int, [C.POINTER(GrandGraph), C.POINTER(GrandAttention), C.POINTER(GrandWorkspace),
                                               C.POINTER(GrandEpilogue), C.c_int32, c_float_p, C.c_int32, _P, _P, _P, _P,
                                               _P, _P, C.POINTER(C.c_void_p), C.POINTER(C.c_int32),
                                               C.POINTER(GrandPeers), _P]),
  'grand_dopri5_create': (C.c_int, [C.POINTER(GrandGraph), C.POINTER(GrandAttention), _P, C.c_int32,
                                     C.POINTER(GrandWorkspace), C.POINTER(GrandEpilogue), _P, _P, _P, _P, C.c_double,
                                     C.c_double, C.c_float, C.c_float, C.c_int32, C.POINTER(C.c_void_p)]),
  'grand_dopri5_launch': (C.c_int, [_P, _P]),
  'grand_dopri5_destroy': (C.c_int, [_P]),
  'grand_plan_pack_rows': (C.c_int64, [_P, C.c_int64, C.c_int32, C.c_int32, _P, _P]),
}


class GrandError(RuntimeError):
  pass


_lib = None


def lib():
  """The loaded library; raises (loudly) when it has not been built -- there is no fallback path."""
  global _lib
  if _lib is None:
    if not os.path.exists(LIB_PATH):
      raise GrandError('libgrand_b200.so is not built: run `python graph-neural-pde_b200/build_native.py` '
                       '(or __graft_entry__.build()); this package has no non-CUDA fallback')
    handle = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
      fn = getattr(handle, name)  # AttributeError if the header and the library disagree
      fn.restype = restype
      fn.argtypes = argtypes
    got = handle.grand_abi_version()
    if got != ABI_VERSION:
      raise GrandError('libgrand_b200.so ABI %d != binding ABI %d' % (got, ABI_VERSION))
    _lib = handle
  return _lib


def check(rc):
  if rc != 0:
    msg = lib().grand_last_error()
    raise GrandError('grand_b200 error %d: %s' % (rc, msg.decode() if msg else '?'))


def ptr(t):
  """Device pointer of a tensor (None -> NULL).  The caller keeps the tensor alive."""
  if t is None:
    return None
  assert t.is_contiguous(), 'grand_b200 kernels need contiguous tensors'
  return t.data_ptr()


def current_stream(device):
  return torch.cuda.current_stream(device).cuda_stream


def require_cuda(t, what):
  if not t.is_cuda:
    raise GrandError('%s lives on %s: the grand_b200 product path is CUDA-only (sm_100a) and has no CPU fallback'
                     % (what, t.device))
  if t.dtype != torch.float32:
    raise GrandError('%s has dtype %s: the kernels compute in fp32' % (what, t.dtype))
