"""Graph plan: the caller's COO `edge_index` re-laid out once for the kernels (include/grand_b200.h `grand_graph`).

  * CSR by edge_index[0] -- the row torch_sparse.spmm sums into == the query node of the attention
    (reference src/function_transformer_attention.py:190-191); a STABLE sort, so edges of a row
    keep the caller's relative order and `eid` maps every CSR position back to the caller's edge position
    (PyG appends self loops at the end of the list, so the caller's order is not sorted).
  * windows ("tiles") of <= GRAND_TILE edges, one warp each, one edge per lane: rows with more edges are cut into
    chunks -- those tiles come first (their index is their partial-sum slot) --, the remaining rows are packed, in row
    order, several whole rows to a window (<= max_rows).  Every row belongs to a tile, also rows without edges (their
    output is -alpha x + beta x0).
  * CSC view (colptr, csc_pos) for attention_norm_idx = 1.
  * work list of the persistent edge kernel (long segments, medium rows, packed windows).

The build itself is native: `grand_plan_create` (csrc/plan_host.cu, host code -- counting sorts and one pass per array)
produces ONE int32 image; this module copies it to the device in one transfer and exposes the arrays as views.  An edge
list that lives on the GPU (the hard-attention block re-selects its edges on the device every training forward,
reference src/block_transformer_hard_attention.py:60-90) is brought to the host first.  tests/_plan_reference.py keeps
the torch-ops restatement this replaced; tests/test_host_logic.py compares the two array by array.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _native as N

PLAN_ARRAYS = ('rowptr', 'col', 'erow', 'eid', 'win_hdr', 'win_rec', 'split_tile0', 'colptr', 'csc_pos', 'seg_desc',
               'lrow_slot0', 'lrow_row', 'tile_row', 'tile_beg', 'tile_nrows', 'tile_split')   # enum grand_plan_array
_SHAPES = {'win_hdr': (-1, 4), 'win_rec': (-1, N.GRAND_TILE, 2), 'seg_desc': (-1, 4)}


def pack_short_rows(deg, tile, max_rows):
  """Greedy packing of consecutive short rows into windows: -> (first row, row count) per window (int64 arrays); the
  sequential pass of the plan (grand_plan_pack_rows, csrc/plan_host.cu).  Rows with more than `tile` edges are skipped
  (they are chunked separately) and close the current window, because a window covers a contiguous CSR range."""
  deg = np.ascontiguousarray(deg, dtype=np.int64)
  n = len(deg)
  first = np.empty(max(n, 1), dtype=np.int64)
  count = np.empty(max(n, 1), dtype=np.int64)
  w = N.lib().grand_plan_pack_rows(deg.ctypes.data, n, int(tile), int(max_rows), first.ctypes.data, count.ctypes.data)
  if w < 0:
    raise N.GrandError('grand_plan_pack_rows failed (%d): bad degrees / tile / max_rows' % w)
  return first[:w].copy(), count[:w].copy()


def default_seg_edges(n_edges):
  """Large graphs: a warp takes a whole row of up to 512 in-edges (fewer partial sums, no statistics pre-pass for them;
  ogbn-arxiv, TMA-ring kernel: 64 -> 362, 128 -> 288, 256 -> 265, 512 -> 254, 1024 -> 257, 2048 -> 296 us per launch).
  Small graphs (everything L2-resident, a handful of units per warp): the launch lasts as long as its longest unit, so
  long rows are cut finer and spread over more warps (measured on B200: Photo 165 -> 115 us per evaluation, Cora
  56 -> 46 with 64-edge segments; ogbn-arxiv 342 us with 256 against 376 us with 128, profiles/README.md)."""
  return N.GRAND_SEG_EDGES if (n_edges >= (1 << 20) or 'GRAND_SEG_EDGES' in os.environ) else 64


class GraphPlan:
  def __init__(self, edge_index, num_nodes, device=None, num_cols=None, tile=N.GRAND_TILE, max_rows=8,
               seg_edges=None):
    assert edge_index.dim() == 2 and edge_index.shape[0] == 2, 'edge_index must be [2, E]'
    device = torch.device(device) if device is not None else edge_index.device
    n_rows = int(num_nodes)
    n_cols = int(num_cols) if num_cols is not None else n_rows
    ei = edge_index.detach().to(device='cpu', dtype=torch.long).contiguous()
    e = int(ei.shape[1])
    assert e < 2 ** 31 and n_cols < 2 ** 31
    if seg_edges is None:
      seg_edges = default_seg_edges(e)
    assert seg_edges >= tile and seg_edges % tile == 0
    lib = N.lib()
    handle = C.c_void_p()
    rc = lib.grand_plan_create(ei.data_ptr() if e > 0 else None, e, n_rows, n_cols, int(tile), int(max_rows),
                               int(seg_edges), C.byref(handle))
    if rc != 0:
      msg = lib.grand_last_error().decode()
      if 'out of range' in msg:
        raise AssertionError('edge_index out of range: ' + msg)
      N.check(rc)
    try:
      counts = N.GrandGraph()
      max_degree, image_elems = C.c_int32(), C.c_int64()
      offsets, elems = (C.c_int64 * len(PLAN_ARRAYS))(), (C.c_int64 * len(PLAN_ARRAYS))()
      N.check(lib.grand_plan_describe(handle, C.byref(counts), C.byref(max_degree), C.byref(image_elems), offsets, elems))
      host = np.ctypeslib.as_array(C.cast(lib.grand_plan_image(handle), C.POINTER(C.c_int32)),
                                   shape=(int(image_elems.value),))
      # ONE transfer, straight from the library's buffer (a synchronous copy: the buffer may be freed right after); every
      # array below is a view into the device image.  A host-side plan (tests) takes its own copy instead.
      self.image = torch.from_numpy(host).to(device) if device.type != 'cpu' else torch.from_numpy(host.copy())
    finally:
      lib.grand_plan_destroy(handle)
    for i, name in enumerate(PLAN_ARRAYS):
      view = self.image[int(offsets[i]):int(offsets[i]) + int(elems[i])]
      setattr(self, name, view.view(*_SHAPES[name]) if name in _SHAPES else view)
    self.device = device
    self.n_rows, self.n_cols, self.n_edges = n_rows, n_cols, e
    self.n_tiles, self.n_split_rows, self.n_split_tiles = counts.n_tiles, counts.n_split_rows, counts.n_split_tiles
    self.n_lunits, self.n_munits, self.n_lrows = counts.n_lunits, counts.n_munits, counts.n_lrows
    self.max_degree = int(max_degree.value)
    self.seg_edges = int(seg_edges)
    self.eid_long = self.eid.long()                # CSR position -> caller position (for torch-side permutations)
    self._struct = None

  def struct(self):
    """ctypes `grand_graph` pointing at this plan's device arrays (the plan must outlive every use)."""
    if self._struct is None:
      g = N.GrandGraph()
      g.n_rows, g.n_cols, g.n_edges, g.n_tiles = self.n_rows, self.n_cols, self.n_edges, self.n_tiles
      g.n_split_rows, g.n_split_tiles = self.n_split_rows, self.n_split_tiles
      g.n_lunits, g.n_munits, g.n_lrows = self.n_lunits, self.n_munits, self.n_lrows
      for name in ('rowptr', 'col', 'erow', 'eid', 'win_hdr', 'win_rec', 'split_tile0', 'colptr', 'csc_pos',
                   'seg_desc', 'lrow_slot0', 'lrow_row'):
        setattr(g, name, getattr(self, name).data_ptr())
      self._struct = g
    return self._struct

  def to_csr_order(self, per_edge):
    """Per-edge values in the caller's order -> CSR order (torch-side; the device kernel is
    grand_edge_values_to_csr)."""
    return per_edge[self.eid_long].contiguous()
