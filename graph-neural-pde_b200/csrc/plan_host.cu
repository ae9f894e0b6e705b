// The graph plan, built natively on the host (no device code): grand_plan_create and its helpers.  graph_plan.py
// only wraps the resulting image in tensors.
#include <stdint.h>

#include "common.cuh"

extern "C" {

// Greedy packing of consecutive short rows into windows of <= tile edges and <= max_rows rows (include/grand_b200.h).
// A row with more than `tile` edges is chunked elsewhere and closes the current window, because a window covers a
// contiguous CSR range.  One pass, O(n_rows).
int64_t grand_plan_pack_rows(const int64_t* deg, int64_t n_rows, int32_t tile, int32_t max_rows, int64_t* first,
                             int64_t* count) {
  if (n_rows < 0 || tile <= 0 || max_rows <= 0 || max_rows > GRAND_MAX_PACKED_ROWS || (n_rows > 0 && (!deg || !first || !count)))
    return GRAND_EINVAL;
  int64_t w = 0, start = -1, edges = 0;
  int32_t rows = 0;
  for (int64_t r = 0; r < n_rows; ++r) {
    const int64_t dr = deg[r];
    if (dr < 0) return GRAND_EINVAL;
    if (dr > tile) {                      // long row: flush and skip
      if (rows) { first[w] = start; count[w] = rows; ++w; rows = 0; }
      continue;
    }
    if (rows && (edges + dr > tile || rows == max_rows)) { first[w] = start; count[w] = rows; ++w; rows = 0; }
    if (rows == 0) { start = r; edges = 0; }
    edges += dr;
    ++rows;
  }
  if (rows) { first[w] = start; count[w] = rows; ++w; }
  return w;
}

// ---- the whole graph plan, natively (include/grand_b200.h: grand_plan_create) -----------------------------------------
}  // extern "C"

#include <algorithm>
#include <memory>
#include <new>
#include <thread>
#include <vector>
#ifdef GRAND_PLAN_TIMING
#include <chrono>
#define GRAND_PLAN_TICK(what) { auto now = std::chrono::steady_clock::now(); fprintf(stderr, "plan %-10s %.1f ms\n", what, std::chrono::duration<double, std::milli>(now - tick).count()); tick = now; }
#else
#define GRAND_PLAN_TICK(what)
#endif

struct grand_plan_host {
  grand_graph counts;                       // scalar fields; pointer fields stay null (grand_plan_bind fills a copy)
  int32_t max_degree, seg_edges;
  int64_t off[GRAND_PLAN_ARRAYS], cnt[GRAND_PLAN_ARRAYS];
  std::unique_ptr<int32_t[]> image;         // all arrays back to back, every start a multiple of 16 bytes; allocated
  int64_t image_elems;                      // uninitialised and written exactly once (no zero-fill pass, no staging copies)
};

namespace {

// rows [of the given list] ordered by degree, longest first, ties in row order (== torch.argsort(descending, stable))
void sort_longest_first(std::vector<int32_t>& rows, const std::vector<int64_t>& deg) {
  std::stable_sort(rows.begin(), rows.end(), [&](int32_t a, int32_t b) { return deg[a] > deg[b]; });
}

// [0, n) cut into contiguous chunks, one host thread each (the calling thread takes the first)
inline int plan_threads(int64_t n) {
  if (n < (1 << 16)) return 1;
  const unsigned hw = std::thread::hardware_concurrency();
  return (int)std::min<unsigned>(8u, hw == 0 ? 1u : hw);
}
template <class F>
void parallel_for(int64_t n, F body) {
  const int T = plan_threads(n);
  if (T == 1) { body(0, n); return; }
  std::vector<std::thread> pool;
  for (int t = 1; t < T; ++t) pool.emplace_back([=]() { body(n * t / T, n * (t + 1) / T); });
  body(0, n / T);
  for (auto& th : pool) th.join();
}

// Stable counting sort in two steps, each split over a few host threads (per-thread histograms over contiguous chunks
// of the input keep equal keys in input order).  count(): ptr[k] = first output position of key k (ptr has n_keys + 1
// entries).  scatter(): order[p] = input index at output position p.  Between the two the caller can size its outputs.
struct CountingSort {
  int T = 1;
  int64_t n = 0;
  int32_t n_keys = 0;
  std::vector<std::vector<int32_t>> hist;

  void chunk(int t, int64_t& lo, int64_t& hi) const { lo = n * t / T; hi = n * (t + 1) / T; }

  template <class Key>
  void count(int64_t n_, int32_t n_keys_, Key key, int32_t* ptr) {
    n = n_; n_keys = n_keys_; T = plan_threads(n);
    hist.assign(T, std::vector<int32_t>((size_t)n_keys + 1, 0));
    std::vector<std::thread> pool;
    auto body = [&](int t) { int64_t lo, hi; chunk(t, lo, hi); for (int64_t i = lo; i < hi; ++i) ++hist[t][key(i)]; };
    for (int t = 1; t < T; ++t) pool.emplace_back(body, t);
    body(0);
    for (auto& th : pool) th.join();
    int32_t run = 0;
    for (int32_t k = 0; k < n_keys; ++k) {       // hist[t][k] becomes thread t's first position for key k
      ptr[k] = run;
      for (int t = 0; t < T; ++t) { const int32_t c = hist[t][k]; hist[t][k] = run; run += c; }
    }
    ptr[n_keys] = run;
  }

  template <class Key>
  void scatter(Key key, int32_t* order) {
    std::vector<std::thread> pool;
    auto body = [&](int t) { int64_t lo, hi; chunk(t, lo, hi); for (int64_t i = lo; i < hi; ++i) order[hist[t][key(i)]++] = (int32_t)i; };
    for (int t = 1; t < T; ++t) pool.emplace_back(body, t);
    body(0);
    for (auto& th : pool) th.join();
    hist.clear();
  }
};

}  // namespace

extern "C" {

int grand_plan_create(const int64_t* edge_index, int64_t n_edges, int32_t n_rows, int32_t n_cols, int32_t tile,
                      int32_t max_rows, int32_t seg_edges, grand_plan_host** out) {
  using grand::set_error;
  if (!out) return GRAND_EINVAL;
  *out = nullptr;
  GRAND_REQUIRE(n_edges >= 0 && n_rows >= 0 && n_cols >= 0 && (n_edges == 0 || edge_index), GRAND_EINVAL,
                "plan_create: null edge list or negative sizes");
  GRAND_REQUIRE(n_edges < ((int64_t)1 << 31) && tile == GRAND_TILE && max_rows > 0 && max_rows <= GRAND_MAX_PACKED_ROWS,
                GRAND_EUNSUPPORTED, "plan_create: needs < 2^31 edges, tile == %d, max_rows <= %d", GRAND_TILE,
                GRAND_MAX_PACKED_ROWS);
  GRAND_REQUIRE(seg_edges >= tile && seg_edges % tile == 0, GRAND_EINVAL, "plan_create: seg_edges must be a multiple of the tile");
  const int64_t* row = edge_index;
  const int64_t* col = edge_index + n_edges;
  const int64_t E = n_edges;
  try {
#ifdef GRAND_PLAN_TIMING
    auto tick = std::chrono::steady_clock::now();
#endif
    // 1. validate; row degrees and CSR row pointer (first half of the stable counting sort by edge_index[0])
    for (int64_t e = 0; e < E; ++e)
      GRAND_REQUIRE(row[e] >= 0 && row[e] < n_rows && col[e] >= 0 && col[e] < n_cols, GRAND_EINVAL,
                    "plan_create: edge %lld = (%lld, %lld) out of range [%d, %d)", (long long)e, (long long)row[e],
                    (long long)col[e], n_rows, n_cols);
    std::vector<int32_t> rowptr_v((size_t)n_rows + 1);
    CountingSort by_row;
    by_row.count(E, n_rows, [&](int64_t e) { return (int32_t)row[e]; }, rowptr_v.data());
    std::vector<int64_t> deg(n_rows);
    int64_t max_degree = 0;
    for (int32_t r = 0; r < n_rows; ++r) { deg[r] = rowptr_v[r + 1] - rowptr_v[r]; max_degree = std::max(max_degree, deg[r]); }
    GRAND_PLAN_TICK("degrees")

    // 2. everything that only needs the degrees: which rows are cut into chunks (longest first), the packed windows, the
    //    work list's long / medium rows -> every array size
    std::vector<int32_t> split_rows, long_rows, med_rows;
    for (int32_t r = 0; r < n_rows; ++r) {
      if (deg[r] > tile) split_rows.push_back(r);
      if (deg[r] > seg_edges) long_rows.push_back(r);
      else if (deg[r] > tile) med_rows.push_back(r);
    }
    sort_longest_first(split_rows, deg);
    sort_longest_first(long_rows, deg);
    sort_longest_first(med_rows, deg);
    std::vector<int32_t> split_tile0(split_rows.size() + 1, 0), lrow_slot0(long_rows.size() + 1, 0);
    for (size_t i = 0; i < split_rows.size(); ++i)
      split_tile0[i + 1] = split_tile0[i] + (int32_t)((deg[split_rows[i]] + tile - 1) / tile);
    for (size_t i = 0; i < long_rows.size(); ++i)
      lrow_slot0[i + 1] = lrow_slot0[i] + (int32_t)((deg[long_rows[i]] + seg_edges - 1) / seg_edges);
    const int32_t n_split_tiles = split_tile0.back(), n_lunits = lrow_slot0.back();
    std::vector<int64_t> pfirst(std::max<int32_t>(n_rows, 1)), pcount(std::max<int32_t>(n_rows, 1));
    const int64_t n_pack = grand_plan_pack_rows(deg.data(), n_rows, tile, max_rows, pfirst.data(), pcount.data());
    if (n_pack < 0) return (int)n_pack;
    const int64_t n_tiles = (int64_t)n_split_tiles + n_pack;
    GRAND_REQUIRE(n_tiles < ((int64_t)1 << 31) / (2 * tile), GRAND_EUNSUPPORTED, "plan_create: too many windows");
    GRAND_PLAN_TICK("lists")

    // 3. the image: sizes -> offsets -> ONE uninitialised allocation; every array below is written in place
    grand_plan_host* P = new (std::nothrow) grand_plan_host();
    GRAND_REQUIRE(P != nullptr, GRAND_ECUDA, "plan_create: out of host memory");
    std::unique_ptr<grand_plan_host> guard(P);
    const int64_t sizes[GRAND_PLAN_ARRAYS] = {
      (int64_t)n_rows + 1, E, E, E, n_tiles * 4, n_tiles * tile * 2, (int64_t)split_tile0.size(), (int64_t)n_cols + 1, E,
      ((int64_t)n_lunits + (int64_t)med_rows.size()) * 4, (int64_t)lrow_slot0.size(), (int64_t)long_rows.size(), n_tiles,
      n_tiles, n_tiles, n_tiles};
    int64_t total = 0;
    for (int i = 0; i < GRAND_PLAN_ARRAYS; ++i) {
      P->off[i] = total;
      P->cnt[i] = sizes[i];
      total += (sizes[i] + 3) & ~(int64_t)3;
    }
    P->image_elems = std::max<int64_t>(total, 4);
    P->image.reset(new int32_t[(size_t)P->image_elems]);
    int32_t* const img = P->image.get();
    for (int i = 0; i < GRAND_PLAN_ARRAYS; ++i)                    // the <= 3 padding elements behind every array
      for (int64_t k = P->off[i] + P->cnt[i]; k < (i + 1 < GRAND_PLAN_ARRAYS ? P->off[i + 1] : P->image_elems); ++k) img[k] = 0;
    int32_t* const rowptr = img + P->off[GRAND_PLAN_ROWPTR];
    int32_t* const col_csr = img + P->off[GRAND_PLAN_COL];
    int32_t* const erow = img + P->off[GRAND_PLAN_EROW];
    int32_t* const eid = img + P->off[GRAND_PLAN_EID];
    int32_t* const win_hdr = img + P->off[GRAND_PLAN_WIN_HDR];
    int32_t* const win_rec = img + P->off[GRAND_PLAN_WIN_REC];
    int32_t* const colptr = img + P->off[GRAND_PLAN_COLPTR];
    int32_t* const csc_pos = img + P->off[GRAND_PLAN_CSC_POS];
    int32_t* const seg_desc = img + P->off[GRAND_PLAN_SEG_DESC];
    int32_t* const tile_row = img + P->off[GRAND_PLAN_TILE_ROW];
    int32_t* const tile_beg = img + P->off[GRAND_PLAN_TILE_BEG];
    int32_t* const tile_nrows = img + P->off[GRAND_PLAN_TILE_NROWS];
    int32_t* const tile_split = img + P->off[GRAND_PLAN_TILE_SPLIT];
    std::copy(rowptr_v.begin(), rowptr_v.end(), rowptr);
    std::copy(split_tile0.begin(), split_tile0.end(), img + P->off[GRAND_PLAN_SPLIT_TILE0]);
    std::copy(lrow_slot0.begin(), lrow_slot0.end(), img + P->off[GRAND_PLAN_LROW_SLOT0]);
    std::copy(long_rows.begin(), long_rows.end(), img + P->off[GRAND_PLAN_LROW_ROW]);

    // 4. CSR (second half of the sort by row: eid; then columns / rows in CSR order) and the CSC view (sort of the CSR
    //    positions by column)
    by_row.scatter([&](int64_t e) { return (int32_t)row[e]; }, eid);
    parallel_for(E, [&](int64_t lo, int64_t hi) {
      for (int64_t p = lo; p < hi; ++p) { col_csr[p] = (int32_t)col[eid[p]]; erow[p] = (int32_t)row[eid[p]]; }
    });
    {
      CountingSort by_col;
      by_col.count(E, n_cols, [&](int64_t p) { return col_csr[p]; }, colptr);
      by_col.scatter([&](int64_t p) { return col_csr[p]; }, csc_pos);
    }
    GRAND_PLAN_TICK("csr+csc")

    // 5. tiles: chunks of the rows with more than `tile` edges first (longest rows first), then the packed windows
    for (size_t i = 0; i < split_rows.size(); ++i) {
      const int32_t r = split_rows[i];
      for (int32_t t = split_tile0[i], k = 0; t < split_tile0[i + 1]; ++t, ++k) {
        tile_row[t] = r; tile_beg[t] = rowptr[r] + k * tile; tile_nrows[t] = 1; tile_split[t] = (int32_t)i;
      }
    }
    for (int64_t w = 0; w < n_pack; ++w) {
      const int64_t t = n_split_tiles + w;
      tile_row[t] = (int32_t)pfirst[w]; tile_beg[t] = rowptr[pfirst[w]]; tile_nrows[t] = (int32_t)pcount[w]; tile_split[t] = -1;
    }
    // window header (int4) and per-lane record (int2), see csrc/edge_common.cuh
    parallel_for(n_tiles, [&](int64_t t_lo, int64_t t_hi) {
    for (int64_t t = t_lo; t < t_hi; ++t) {
      const int32_t r0 = tile_row[t], beg = tile_beg[t], nr = tile_nrows[t];
      const bool is_split = tile_split[t] >= 0;
      const int32_t end = is_split ? std::min(beg + tile, rowptr[r0 + 1]) : rowptr[r0 + nr];
      const int32_t cnt = end - beg;
      win_hdr[4 * t + 0] = r0; win_hdr[4 * t + 1] = beg; win_hdr[4 * t + 2] = nr | (cnt << 8); win_hdr[4 * t + 3] = tile_split[t];
      for (int32_t lane = 0; lane < tile; ++lane) {
        const bool valid = lane < cnt;
        int32_t lane_col = 0, kk = 0, seg_lo = lane, seg_hi = lane, row_lo = 0, row_hi = 0;
        if (valid) {
          const int32_t p = beg + lane;
          lane_col = col_csr[p];
          if (is_split) {
            seg_lo = 0; seg_hi = cnt;
          } else {
            kk = erow[p] - r0;
            seg_lo = rowptr[r0 + kk] - beg; seg_hi = rowptr[r0 + kk + 1] - beg;
          }
        }
        if (lane < nr && !is_split) {
          const int32_t rr = std::min(r0 + lane, n_rows - 1);
          row_lo = rowptr[rr] - beg; row_hi = rowptr[rr + 1] - beg;
        }
        const uint32_t packed = (uint32_t)kk | ((uint32_t)seg_lo << 8) | ((uint32_t)seg_hi << 14) |
                                ((uint32_t)row_lo << 20) | ((uint32_t)row_hi << 26);
        win_rec[(t * tile + lane) * 2 + 0] = lane_col;
        win_rec[(t * tile + lane) * 2 + 1] = (int32_t)packed;
      }
    }
    });
    GRAND_PLAN_TICK("windows")
    // 6. work list of the persistent edge kernel: long segments (longest row first), medium rows (longest first)
    for (size_t i = 0; i < long_rows.size(); ++i) {
      const int32_t r = long_rows[i];
      for (int32_t u = lrow_slot0[i], k = 0; u < lrow_slot0[i + 1]; ++u, ++k) {
        const int32_t b = rowptr[r] + k * seg_edges;
        seg_desc[4 * u + 0] = r; seg_desc[4 * u + 1] = b; seg_desc[4 * u + 2] = std::min(seg_edges, rowptr[r + 1] - b);
        seg_desc[4 * u + 3] = (int32_t)i;
      }
    }
    for (size_t i = 0; i < med_rows.size(); ++i) {
      const size_t u = (size_t)n_lunits + i;
      const int32_t r = med_rows[i];
      seg_desc[4 * u + 0] = r; seg_desc[4 * u + 1] = rowptr[r]; seg_desc[4 * u + 2] = (int32_t)deg[r]; seg_desc[4 * u + 3] = -1;
    }
    GRAND_PLAN_TICK("worklist")
    grand_graph& g = P->counts;
    g = grand_graph();
    g.n_rows = n_rows; g.n_cols = n_cols; g.n_edges = (int32_t)E; g.n_tiles = (int32_t)n_tiles;
    g.n_split_rows = (int32_t)split_rows.size(); g.n_split_tiles = n_split_tiles;
    g.n_lunits = n_lunits; g.n_munits = (int32_t)med_rows.size(); g.n_lrows = (int32_t)long_rows.size();
    guard.release();
    P->max_degree = (int32_t)max_degree;
    P->seg_edges = seg_edges;
    *out = P;
    return GRAND_OK;
  } catch (const std::bad_alloc&) {
    set_error("plan_create: out of host memory");
    return GRAND_ECUDA;
  }
}

int grand_plan_describe(const grand_plan_host* plan, grand_graph* counts, int32_t* max_degree, int64_t* image_elems,
                        int64_t* offsets, int64_t* elems) {
  if (!plan) return GRAND_EINVAL;
  if (counts) *counts = plan->counts;
  if (max_degree) *max_degree = plan->max_degree;
  if (image_elems) *image_elems = plan->image_elems;
  for (int i = 0; i < GRAND_PLAN_ARRAYS; ++i) {
    if (offsets) offsets[i] = plan->off[i];
    if (elems) elems[i] = plan->cnt[i];
  }
  return GRAND_OK;
}

const int32_t* grand_plan_image(const grand_plan_host* plan) { return plan ? plan->image.get() : nullptr; }

int grand_plan_bind(const grand_plan_host* plan, const int32_t* image, grand_graph* g) {
  if (!plan || !image || !g) return GRAND_EINVAL;
  *g = plan->counts;
  const int32_t** slots[12] = {&g->rowptr, &g->col, &g->erow, &g->eid, &g->win_hdr, &g->win_rec, &g->split_tile0,
                               &g->colptr, &g->csc_pos, &g->seg_desc, &g->lrow_slot0, &g->lrow_row};
  for (int i = 0; i < 12; ++i) *slots[i] = image + plan->off[i];
  return GRAND_OK;
}

void grand_plan_destroy(grand_plan_host* plan) { delete plan; }

}  // extern "C"
