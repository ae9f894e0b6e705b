// Host-side part of the graph plan (no device code): the one step of graph_plan.py that is inherently sequential.
// Everything else of the plan is data-parallel index arithmetic done with torch ops on the device.
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

}  // extern "C"
