"""TEST INFRASTRUCTURE: the graph plan restated with torch index ops (this was the product builder until the native
`grand_plan_create` replaced it).  tests/test_host_logic.py holds the native plan to this restatement array by array."""
import numpy as np
import torch


def pack_short_rows_ref(deg, tile, max_rows):
  first, count = [], []
  rows = edges = 0
  start = -1
  for r, dr in enumerate(np.asarray(deg).tolist()):
    if dr > tile:
      if rows:
        first.append(start); count.append(rows); rows = 0
      continue
    if rows and (edges + dr > tile or rows == max_rows):
      first.append(start); count.append(rows); rows = 0
    if rows == 0:
      start, edges = r, 0
    edges += dr
    rows += 1
  if rows:
    first.append(start); count.append(rows)
  return np.asarray(first, dtype=np.int64), np.asarray(count, dtype=np.int64)


class TorchPlan:
  def __init__(self, edge_index, num_nodes, device=None, num_cols=None, tile=32, max_rows=8,
               seg_edges=None):
    assert edge_index.dim() == 2 and edge_index.shape[0] == 2, 'edge_index must be [2, E]'
    device = torch.device(device) if device is not None else edge_index.device
    ei = edge_index.to(device=device, dtype=torch.long)
    n_rows = int(num_nodes)
    n_cols = int(num_cols) if num_cols is not None else n_rows
    e = ei.shape[1]
    if e > 0:
      assert int(ei[0].max()) < n_rows and int(ei[1].max()) < n_cols and int(ei.min()) >= 0, 'edge_index out of range'
    assert e < 2 ** 31 and n_cols < 2 ** 31
    row, col = ei[0], ei[1]
    perm = torch.argsort(row, stable=True)
    deg = torch.bincount(row, minlength=n_rows)
    rowptr = torch.zeros(n_rows + 1, dtype=torch.long, device=device)
    rowptr[1:] = torch.cumsum(deg, 0)

    n_tiles_of_row = torch.clamp((deg + tile - 1) // tile, min=1)
    split_rows = torch.nonzero(n_tiles_of_row > 1).flatten()
    # longest rows first: their chunks are scheduled first and hide the tail
    if split_rows.numel() > 0:
      order = torch.argsort(deg[split_rows], descending=True, stable=True)
      split_rows = split_rows[order]
    nt_split = n_tiles_of_row[split_rows]
    split_tile0 = torch.zeros(split_rows.numel() + 1, dtype=torch.long, device=device)
    split_tile0[1:] = torch.cumsum(nt_split, 0)
    n_split_tiles = int(split_tile0[-1])
    split_index = torch.arange(split_rows.numel(), device=device)
    st_split = torch.repeat_interleave(split_index, nt_split)                 # split-row index of each split tile
    st_k = torch.arange(n_split_tiles, device=device) - split_tile0[:-1][st_split]
    st_row = split_rows[st_split]
    st_beg = rowptr[st_row] + st_k * tile
    # short rows (<= tile edges, also empty ones): consecutive rows are packed greedily into windows of <= tile
    # edges and <= max_rows rows; a long row ends the current window
    pr0, pn = pack_short_rows_ref(deg.cpu().numpy(), tile, max_rows)
    pack_row = torch.from_numpy(pr0).to(device)
    pack_nrows = torch.from_numpy(pn).to(device)
    tile_row = torch.cat([st_row, pack_row])
    tile_beg = torch.cat([st_beg, rowptr[pack_row]])
    tile_nrows = torch.cat([torch.ones_like(st_row), pack_nrows])
    tile_split = torch.cat([st_split, torch.full_like(pack_row, -1)])

    col_csr = col[perm]
    # window header (int4) and per-lane record (int2), see csrc/edge_common.cuh
    n_t = int(tile_row.numel())
    is_split = tile_split >= 0
    t_end = torch.where(is_split, torch.minimum(tile_beg + tile, rowptr[tile_row + 1]), rowptr[tile_row + tile_nrows])
    t_cnt = t_end - tile_beg
    win_hdr = torch.stack([tile_row, tile_beg, tile_nrows | (t_cnt << 8), tile_split], dim=1)
    lane = torch.arange(tile, device=device).unsqueeze(0)                       # [1, 32]
    pos = tile_beg.unsqueeze(1) + lane                                         # CSR position of every lane
    lane_valid = lane < t_cnt.unsqueeze(1)
    pos_c = torch.where(lane_valid, pos, torch.zeros_like(pos)).clamp(max=max(e - 1, 0))
    lane_col = torch.where(lane_valid, col_csr[pos_c] if e > 0 else torch.zeros_like(pos), torch.zeros_like(pos))
    lane_row = (row[perm][pos_c] if e > 0 else torch.zeros_like(pos))
    kk = torch.where(lane_valid & ~is_split.unsqueeze(1), lane_row - tile_row.unsqueeze(1), torch.zeros_like(pos))
    # lane range of the lane's own row inside the window
    own_row = tile_row.unsqueeze(1) + kk
    seg_lo = torch.where(is_split.unsqueeze(1), torch.zeros_like(pos), rowptr[own_row] - tile_beg.unsqueeze(1))
    seg_hi = torch.where(is_split.unsqueeze(1), t_cnt.unsqueeze(1).expand_as(pos), rowptr[own_row + 1] - tile_beg.unsqueeze(1))
    seg_lo = torch.where(lane_valid, seg_lo, lane.expand_as(pos))
    seg_hi = torch.where(lane_valid, seg_hi, lane.expand_as(pos))
    # lane range of row number `lane` of the window (lane < nrows)
    has_row = lane < tile_nrows.unsqueeze(1)
    rrow = (tile_row.unsqueeze(1) + lane).clamp(max=max(n_rows - 1, 0))
    row_lo = torch.where(has_row & ~is_split.unsqueeze(1), rowptr[rrow] - tile_beg.unsqueeze(1), torch.zeros_like(pos))
    row_hi = torch.where(has_row & ~is_split.unsqueeze(1), rowptr[rrow + 1] - tile_beg.unsqueeze(1), torch.zeros_like(pos))
    packed = kk | (seg_lo << 8) | (seg_hi << 14) | (row_lo << 20) | (row_hi << 26)
    packed = torch.where(packed >= 2 ** 31, packed - 2 ** 32, packed)           # two's complement into int32
    win_rec = torch.stack([lane_col, packed], dim=2)                            # [n_tiles, 32, 2]
    assert n_t == 0 or (int(t_cnt.max()) <= tile and int(t_cnt.min()) >= 0)
    perm_c = torch.argsort(col_csr, stable=True)
    cdeg = torch.bincount(col, minlength=n_cols)
    colptr = torch.zeros(n_cols + 1, dtype=torch.long, device=device)
    colptr[1:] = torch.cumsum(cdeg, 0)

    # ---- work list of the persistent edge kernel (csrc/rhs_stream.cu): rows with more than `tile` in-edges are either
    #      "medium" (<= seg_edges: one warp takes the whole row) or "long" (cut into segments of seg_edges edges whose
    #      partial sums are merged by the last segment to finish).  Longest first: the tail of the kernel is made of the
    #      small packed windows.
    assert seg_edges is not None   # the default (by graph size) is chosen by graph_plan.GraphPlan
    assert seg_edges >= tile and seg_edges % tile == 0
    long_rows = torch.nonzero(deg > seg_edges).flatten()
    long_rows = long_rows[torch.argsort(deg[long_rows], descending=True, stable=True)]
    med_rows = torch.nonzero((deg > tile) & (deg <= seg_edges)).flatten()
    med_rows = med_rows[torch.argsort(deg[med_rows], descending=True, stable=True)]
    nseg = (deg[long_rows] + seg_edges - 1) // seg_edges
    lrow_slot0 = torch.zeros(long_rows.numel() + 1, dtype=torch.long, device=device)
    lrow_slot0[1:] = torch.cumsum(nseg, 0)
    n_lunits = int(lrow_slot0[-1])
    l_index = torch.repeat_interleave(torch.arange(long_rows.numel(), device=device), nseg)
    l_k = torch.arange(n_lunits, device=device) - lrow_slot0[:-1][l_index]
    l_row = long_rows[l_index]
    l_beg = rowptr[l_row] + l_k * seg_edges
    l_cnt = torch.minimum(torch.full_like(l_beg, seg_edges), rowptr[l_row + 1] - l_beg)
    seg_desc = torch.cat([torch.stack([l_row, l_beg, l_cnt, l_index], dim=1),
                          torch.stack([med_rows, rowptr[med_rows], deg[med_rows], torch.full_like(med_rows, -1)], dim=1)])

    i32 = lambda t: t.to(torch.int32).contiguous()
    self.device = device
    self.n_rows, self.n_cols, self.n_edges = n_rows, n_cols, e
    self.n_tiles = int(tile_row.numel())
    self.n_split_rows = int(split_rows.numel())
    self.n_split_tiles = n_split_tiles
    self.max_degree = int(deg.max()) if n_rows > 0 and e > 0 else 0
    self.rowptr, self.col, self.erow, self.eid = i32(rowptr), i32(col_csr), i32(row[perm]), i32(perm)
    self.tile_row, self.tile_beg, self.tile_split = i32(tile_row), i32(tile_beg), i32(tile_split)
    self.tile_nrows = i32(tile_nrows)
    self.win_hdr, self.win_rec = i32(win_hdr), i32(win_rec)
    self.split_tile0 = i32(split_tile0)
    self.colptr, self.csc_pos = i32(colptr), i32(perm_c)
    self.seg_edges = int(seg_edges)
    self.n_lunits, self.n_munits, self.n_lrows = n_lunits, int(med_rows.numel()), int(long_rows.numel())
    self.seg_desc, self.lrow_slot0, self.lrow_row = i32(seg_desc), i32(lrow_slot0), i32(long_rows)
    self.eid_long = perm  # CSR position -> caller position (int64, for torch-side permutations)
    self._struct = None

