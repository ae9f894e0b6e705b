/*
 * grand_b200.h -- C ABI of the B200-native GRAND / BLEND ODE right-hand-side engine.
 *
 * The reference (twitter-research/graph-neural-pde) is pure Python and has no FFI of its own; its "operator API"
 * for this path is the Python class contract `ODEFunc.forward(t, x)` / `ODEblock.forward(x)` selected by
 * `set_function` / `set_block` (reference src/model_configurations.py:17-44).  This header is the boundary a
 * maintainer binds UNDER that contract (ctypes stub in INTEGRATION.md): plain C, raw device pointers and sizes,
 * one cudaStream_t (passed as void*), no torch types.
 *
 * Rules common to every entry point
 *   - returns 0 on success, a negative GRAND_E* code otherwise; `grand_last_error()` has the message;
 *   - never allocates, never synchronises, keeps no global state: every buffer (inputs, outputs, workspace) is
 *     caller-owned device memory, so every call is CUDA-Graph capturable on `stream`;
 *   - all floating-point data is fp32 row-major, all indices int32; sm_100a only.
 *
 * Each entry point cites the reference code it replaces (paths relative to the reference root).
 */
#ifndef GRAND_B200_H_
#define GRAND_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GRAND_ABI_VERSION 3

/* error codes */
#define GRAND_OK 0
#define GRAND_EINVAL (-1)      /* bad argument (null pointer, negative size, unknown enum)            */
#define GRAND_EUNSUPPORTED (-2)/* shape outside what the kernels are built for (see grand_limits)     */
#define GRAND_ECUDA (-3)       /* a CUDA runtime call failed; message has cudaGetErrorString          */

/* edges per window (one warp, one edge per lane): short rows are packed several to a window, rows with more
 * in-edges are cut into chunks of this size */
#define GRAND_TILE 32
#define GRAND_MAX_PACKED_ROWS 8  /* whole rows packed into one window: at most this many */
#define GRAND_MAX_HEADS 8
#define GRAND_MAX_TERMS 8
#define GRAND_MAX_PEERS 8        /* GPUs of one NVLink / NVSwitch domain that share a node partition */

/* attention_type, reference src/function_transformer_attention.py:193-206 */
enum grand_score_type {
  GRAND_SCORE_SCALED_DOT = 0,  /* q.k / sqrt(d_k)                                   :195-196 */
  GRAND_SCORE_COSINE = 1,      /* CosineSimilarity(eps=1e-5)                        :197-199 */
  GRAND_SCORE_PEARSON = 2,     /* cosine of mean-centred q, k                       :200-206 */
  GRAND_SCORE_EXP_KERNEL = 3,  /* ov^2 exp(-|q-k|^2 / (2 ls^2))                     :193-194 */
  GRAND_SCORE_EXP_KERNEL2 = 4, /* beltrami: product of a feature and a positional exp kernel :133-169 */
  GRAND_SCORE_GAT = 5          /* GAT additive scorer leaky_relu(a_src.h_i + a_dst.h_j), the same `a` for every head:
                                  reference src/function_GAT_attention.py:105-114.  q = k = x W (wq = wk = W^T, zero
                                  biases); kernel_scalars = device [2 d_k + 1]: a_src, a_dst, negative slope */
};

/* normaliser, reference :210-213 and src/utils.py:179-208 */
enum grand_norm_type {
  GRAND_NORM_SOFTMAX = 0,      /* exp(s - max_seg) / (sum_seg + 1e-16), PyG softmax            */
  GRAND_NORM_SQUAREPLUS = 1    /* o = s - max_GLOBAL; (o + sqrt(o^2 + 4))/2 / (sum_seg + 1e-16) */
};

/*
 * The graph the ODE function owns (`odefunc.edge_index`, reference src/block_constant.py:24), re-laid out once:
 * CSR by edge_index[0] (the row that torch_sparse.spmm sums into == the query node), stable so
 * that edges of a row keep the caller's order, cut into row tiles of <= GRAND_TILE edges, plus a CSC view for
 * attention_norm_idx = 1.  All arrays are device int32.  Built by the host mirror (graph_plan.py); the row ids may
 * be a shard's local rows while `col` indexes the shard's [own | halo] feature rows (multi-GPU).
 */
typedef struct grand_graph {
  int32_t n_rows;              /* rows to produce (nodes owned)                                       */
  int32_t n_cols;              /* rows addressable through `col` (own + halo); == n_rows on one GPU   */
  int32_t n_edges;             /* directed edges incl. self loops                                     */
  int32_t n_tiles;             /* row tiles                                                           */
  int32_t n_split_rows;        /* rows cut into more than one tile                                    */
  int32_t n_split_tiles;       /* tiles of split rows; they are tiles [0, n_split_tiles), slot == tile */
  const int32_t* rowptr;       /* [n_rows + 1]                                                        */
  const int32_t* col;          /* [n_edges] edge_index[1] in CSR order                                */
  const int32_t* erow;         /* [n_edges] edge_index[0] in CSR order                                */
  const int32_t* eid;          /* [n_edges] CSR position -> position in the caller's edge list        */
  const int32_t* win_hdr;      /* [n_tiles, 4] per window: first row, first CSR position, nrows | cnt << 8,
                                  split-row index or -1                                                */
  const int32_t* win_rec;      /* [n_tiles, 32, 2] per lane: column; row-in-window | lo << 8 | hi << 14 |
                                  row-lane-range (lo << 20 | hi << 26) -- see csrc/edge_common.cuh      */
  const int32_t* split_tile0;  /* [n_split_rows + 1] first split-tile slot of each split row          */
  const int32_t* colptr;       /* [n_cols + 1] CSC by edge_index[1] (may be null if norm_idx 1 unused)*/
  const int32_t* csc_pos;      /* [n_edges] CSC position -> CSR position                              */
  /* Work list of the persistent edge kernel (csrc/rhs_stream.cu).  Rows with <= GRAND_TILE in-edges are the packed
   * windows [n_split_tiles, n_tiles) above; a row with more is ONE "medium" unit (<= seg_edges in-edges, a warp takes
   * the whole row) or is cut into "long" segments of seg_edges edges.  Units are ordered long segments (longest row
   * first), medium rows (longest first), packed windows (row order).  seg_desc may be null when n_split_tiles == 0. */
  int32_t n_lunits;            /* long-row segments; unit index == partial-sum slot                   */
  int32_t n_munits;            /* medium rows                                                         */
  int32_t n_lrows;             /* rows cut into long segments                                         */
  const int32_t* seg_desc;     /* [n_lunits + n_munits, 4]: row, first CSR position, edge count, long-row index or -1 */
  const int32_t* lrow_slot0;   /* [n_lrows + 1] first segment of each long row                        */
  const int32_t* lrow_row;     /* [n_lrows] row id of each long row                                   */
} grand_graph;

/*
 * Attention layer description == SpGraphTransAttentionLayer's state + opt
 * (reference src/function_transformer_attention.py:59-126).  Weights are nn.Linear layout [out, in].
 */
typedef struct grand_attention {
  int32_t d;                   /* in_features == state width (hidden_dim)                             */
  int32_t ad;                  /* attention_dim                                                       */
  int32_t heads;               /* opt['heads'], ad % heads == 0 (:80-81)                              */
  int32_t score_type;          /* enum grand_score_type                                               */
  int32_t norm_type;           /* enum grand_norm_type (opt['square_plus'])                           */
  int32_t norm_idx;            /* opt['attention_norm_idx'] 0: segments share edge[0], 1: edge[1]     */
  const float* wq;             /* Q.weight [ad, d]  (EXP_KERNEL2: block-diagonal [2 ad, d], see host)  */
  const float* bq;             /* Q.bias   [ad]                                                       */
  const float* wk;             /* K.weight [ad, d]                                                    */
  const float* bk;             /* K.bias   [ad]                                                       */
  const float* kernel_scalars; /* device [4]: output_var, lengthscale (, output_var_p, lengthscale_p);
                                  GRAND_SCORE_GAT: [2 d_k + 1], see the enum                            */
  const float* edge_weight_csr;/* [n_edges] reweight_attention weights in CSR order, or null (:208-209)*/
} grand_attention;

/*
 * Node-partitioned multi-GPU (new design; the reference has no distributed path).  Every rank owns a
 * contiguous range of rows; the rows of other ranks that its edges reference -- the HALO -- live behind its own rows in
 * every state / K buffer ([own rows | halo rows grouped by owner]).  All such buffers of a rank sit in ONE device
 * allocation, the ARENA, laid out identically on every rank and mapped into every other rank's address space (CUDA
 * IPC over NVLink): a rank delivers a row by plain stores through arena[peer] at the same byte offset the buffer has
 * in its own arena.  push_ptr / push_dst list, for every own row, which peers hold it in their halo and at which row.
 */
typedef struct grand_peers {
  int32_t world;               /* ranks (<= GRAND_MAX_PEERS)                                          */
  int32_t rank;
  void* arena[GRAND_MAX_PEERS];/* this process's mapping of every rank's arena; arena[rank] is its own */
  int64_t flags_offset;        /* byte offset inside every arena of int32 flags[GRAND_MAX_PEERS] (arrivals, written by
                                  the peers), int32 epoch (barrier counter), int32 error: zero at start  */
  const int32_t* push_ptr;     /* [n_rows + 1]                                                        */
  const int32_t* push_dst;     /* [push_ptr[n_rows], 2]: peer, destination row in the peer's buffers  */
  const int32_t* push_src;     /* [push_ptr[n_rows]] own row of every push entry (the CSR above, flattened) */
  int32_t n_push;              /* push_ptr[n_rows]                                                    */
} grand_peers;

/*
 * What happens to the RHS value f_i once a row is finished -- the residual of
 * reference src/function_transformer_attention.py:46-53 / src/function_laplacian_diffusion.py:43-51
 *     f = alpha (ax - x) [+ beta x0],  alpha = sigmoid(*alpha_train) unless no_alpha_sigmoid
 * fused with the Runge-Kutta stage combine torchdiffeq does in separate kernels (rk4_alt_step_func,
 * _runge_kutta_step):   y_out = sum_t coef[t] * term[t] + coef_f * f.
 */
typedef struct grand_epilogue {
  const float* alpha_train;    /* device scalar                                                       */
  const float* beta_train;     /* device scalar (read only if x0 != null)                             */
  const float* x0;             /* [n_rows, d] or null (opt['add_source'])                             */
  int32_t no_alpha_sigmoid;
  int32_t n_terms;             /* <= GRAND_MAX_TERMS                                                  */
  float* f_out;                /* [n_rows, d] or null                                                 */
  float* y_out;                /* [n_rows, d] or null                                                 */
  float coef_f;
  float coef[GRAND_MAX_TERMS];
  const float* term[GRAND_MAX_TERMS];
  const grand_peers* peers;    /* or null.  Non-null: y_out lies in the arena and every finished row of it is ALSO
                                  stored into the halo of the peers that reference it, from the same kernel, as the
                                  row is produced (the exchange overlaps the rest of the launch)         */
  const float* coef_dev;       /* or null.  Non-null: device [1 + GRAND_MAX_TERMS] = coef_f, coef[0..] read BY THE KERNEL
                                  instead of the host values above (device-side step-size control, grand_dopri5_*) */
} grand_epilogue;

/* caller-owned scratch; sizes from grand_workspace_sizes */
typedef struct grand_workspace {
  float* q;                    /* [n_rows, ad_eff]                                                    */
  float* k;                    /* [n_cols, ad_eff]                                                    */
  float* scores;               /* [n_edges, heads]   general path only                                */
  float* seg_stats;            /* [n_cols, heads, 2] general path only: (max, sum) per segment        */
  float* split_stats;          /* [n_split_tiles, heads, 2] per chunk of a long row: (max, sum-exp)   */
  float* row_stats;            /* [n_split_rows, heads, 2] the same merged per long row               */
  float* split_partial;        /* [n_split_tiles, d]                                                  */
  int32_t* split_counter;      /* [n_split_rows], zero on first use; kernels leave it zero            */
  float* gmax;                 /* [1] global max for squareplus (ordered-int encoded during reduction)*/
  float* lrow_stats;           /* [n_lunits, heads, 2] softmax (max, sum-exp) of every long-row segment */
  float* lpartial;             /* [n_lunits, max(heads,1), d] partial sums of the long-row segments   */
  int32_t* lcounter;           /* [n_lrows], zero on first use; the kernel leaves it zero (atomicInc wraps) */
  int32_t* coop_sync;          /* [4] zero on first use, or null: barrier words of the one-launch integration of small
                                  graphs (csrc/rhs_coop.cu); null = always the per-evaluation launches          */
} grand_workspace;

int grand_abi_version(void);
/* number of kernel launches (and async copies) this library has issued so far in this process */
int64_t grand_launch_count(void);
const char* grand_last_error(void);

/* element counts of each workspace array for a graph / attention shape; any pointer may be null */
int grand_workspace_sizes(const grand_graph* g, const grand_attention* a, int64_t* n_q, int64_t* n_k,
                          int64_t* n_scores, int64_t* n_seg_stats, int64_t* n_split_stats,
                          int64_t* n_split_partial, int64_t* n_split_counter, int64_t* n_lrow_stats,
                          int64_t* n_lpartial, int64_t* n_lcounter);

/* Q = x Wq^T + bq over n_rows rows, K = x Wk^T + bk over n_cols rows.
 * Replaces the Q/K nn.Linear calls, reference src/function_transformer_attention.py:174-175 (V, :176, is unused on
 * the mix_features=False path and is not computed). */
int grand_project_qk(const float* x, int32_t n_rows, int32_t n_cols, const grand_attention* a, float* q, float* k,
                     void* stream);

/* y = x w^T + b: the dense layers of the model shell around the ODE block -- encoder m1 (and m11 / m12 with use_mlp),
 * decoder m2 (reference src/GNN.py:23-31, 60-71; src/base_classes.py:110-127).  x [n_rows, d_in], w [d_out, d_in]
 * (nn.Linear layout), b [d_out] or null, y [n_rows, ldy].  d_in <= 128 (multiple of 4) with d_out a multiple of 4 runs on
 * the tcgen05 / TMEM / TMA kernel of csrc/project_tc.cu (3xTF32, fp32-class accuracy), other shapes on an exact fp32
 * tiled kernel. */
int grand_linear(const float* x, int32_t n_rows, int32_t d_in, const float* w, const float* b, int32_t d_out, float* y,
                 int32_t ldy, void* stream);

/* Measurement switch (process-wide): which kernel computes the projections -- 3 tcgen05 with the split X chunks in tensor
 * memory, 2 tcgen05 with all operands in shared memory (both csrc/project_tc.cu), 1 mma.sync
 * (csrc/project_mma.cu), 0 fp32 FFMA (csrc/project.cu), -1 back to the default (or the GRAND_PROJ environment variable).
 * A kernel that does not cover a shape hands over to the next. */
int grand_set_projection_impl(int32_t impl);

/* One GRAND-nl RHS evaluation: ODEFuncTransformerAtt.forward (reference :38-53) = projection, per-edge scores,
 * segment softmax / squareplus, head mean, (A (.) a) X aggregation, residual, RK stage combine.  Picks the fused
 * single-pass kernel when the configuration allows it (norm_idx 0, softmax, heads <= GRAND_MAX_HEADS), the general
 * scores -> segment-stats -> aggregate pipeline otherwise. */
int grand_rhs_attention(const grand_graph* g, const grand_attention* a, const float* x, const grand_workspace* ws,
                        const grand_epilogue* ep, void* stream);

/* The edge phase of grand_rhs_attention alone: ws->q / ws->k already hold the projections (grand_project_qk).  Used
 * where a halo exchange sits between projection and aggregation (node-partitioned multi-GPU) and by bench.py to time
 * the dominant kernel on its own. */
int grand_rhs_attention_edges(const grand_graph* g, const grand_attention* a, const float* x, const grand_workspace* ws,
                              const grand_epilogue* ep, void* stream);

/* Vector-Jacobian product of grand_rhs_attention (edge phase) for the fused configuration (scaled_dot, softmax,
 * attention_norm_idx 0); the reference obtains it from autograd over src/function_transformer_attention.py:38-53.
 * Given grad_f = dL/df [n_rows, d]:  abar_out[E] (head-mean attention, CALLER edge order -- feed it to the transposed
 * aggregation alpha (Abar^T g - g) = grand_rhs_laplacian on the flipped graph), dq [n_rows, ad] = dL/dQ and
 * dk [n_cols, ad] = dL/dK (both must be zero on entry; the dense tails dX += dQ Wq + dK Wk, dWq = dQ^T X ... are plain
 * GEMMs left to the caller).  row_dot: [n_split_rows, heads] scratch, zero on entry.  Recomputes Q/K into ws. */
int grand_rhs_attention_backward(const grand_graph* g, const grand_attention* a, const float* x, const float* grad_f,
                                 const grand_workspace* ws, const float* alpha_train, int32_t no_alpha_sigmoid,
                                 float* row_dot, float* abar_out, float* dq, float* dk, void* stream);

/* Gradients through FROZEN attention (AttODEblock, reference src/block_transformer_attention.py:36-39: the attention is
 * evaluated once at x(0) and the Laplacian function integrates with its head mean).  The reference gets them from
 * autograd; here
 *   grand_edge_dot                     c[p] = alpha * grad_f[row_p] . x[col_p] (CSR order): the gradient of ONE Laplacian
 *                                      RHS evaluation w.r.t. its per-edge weights; autograd sums it over the evaluations,
 *   grand_attention_weights_backward   given that sum c = dL/d(mean_h attention) in CSR order: softmax + head-mean backward
 *                                      of SpGraphTransAttentionLayer.forward at x0 -> dq [n_rows, ad], dk [n_cols, ad]
 *                                      (zero on entry; dense tails dWq = dq^T x0 ... are GEMMs left to the caller).
 *                                      row_dot: [n_split_rows, heads] scratch, zero on entry.  scaled_dot + softmax +
 *                                      attention_norm_idx 0 only. */
int grand_edge_dot(const grand_graph* g, int32_t d, const float* x, const float* grad_f, const float* alpha_train,
                   int32_t no_alpha_sigmoid, float* c_csr_out, void* stream);
int grand_attention_weights_backward(const grand_graph* g, const grand_attention* a, const float* x0,
                                     const grand_workspace* ws, const float* c_csr, float* row_dot, float* dq, float* dk,
                                     void* stream);

/* One GRAND-l RHS evaluation: LaplacianODEFunc.forward (reference src/function_laplacian_diffusion.py:38-51);
 * `w_csr` is the per-edge weight sparse_multiply would use (:28-36) in CSR order. */
int grand_rhs_laplacian(const grand_graph* g, int32_t d, const float* x, const float* w_csr, const grand_workspace* ws,
                        const grand_epilogue* ep, void* stream);

/* SpGraphTransAttentionLayer.forward (reference :128-214) as a standalone op: attention[E, H] and prods[E, H] in the
 * CALLER's edge order (either may be null), and optionally the head-mean in CSR order for the Laplacian path
 * (reference src/function_laplacian_diffusion.py:29-31). */
int grand_attention_weights(const grand_graph* g, const grand_attention* a, const float* x, const grand_workspace* ws,
                            float* attention_out, float* prods_out, float* mean_csr_out, void* stream);

/* w_csr[p] = w[eid[p]] (heads == 1) or mean_h w[eid[p], h]: edge weights / frozen attention in caller order -> CSR. */
int grand_edge_values_to_csr(const grand_graph* g, const float* w, int32_t heads, float* w_csr, void* stream);

/* Fixed-grid integration, torchdiffeq FixedGridODESolver.integrate with method euler / midpoint / rk4 (= the 3/8
 * rule rk4_alt_step_func; in-repo witness reference src/early_stop_solver.py:150-155,163-184).  `dt[n_steps]` are
 * the host-side grid increments.  y (in: y(t0), out: y(t_end)) and 5 stage buffers of [n_rows, d] are caller-owned.
 * a == null selects the Laplacian function with `w_csr`.  Every stage is one RHS launch whose epilogue writes the
 * next stage's input; no axpy kernels. */
enum grand_method { GRAND_EULER = 0, GRAND_MIDPOINT = 1, GRAND_RK4 = 2 };
int grand_integrate_fixed(const grand_graph* g, const grand_attention* a, const float* w_csr, int32_t d,
                          const grand_workspace* ws, const grand_epilogue* residual, int32_t method,
                          const float* dt, int32_t n_steps, float* y, float* stage_buffers, void* stream);

/* out[r, :] = x[idx[r], :] for r < n_idx: packs the boundary rows other shards need into a contiguous send buffer
 * (node-partitioned multi-GPU halo exchange; new design, the reference has no distributed path). */
int grand_pack_rows(const float* x, const int32_t* idx, int64_t n_idx, int32_t d, float* out, void* stream);

/* y_out = sum_t coef[t] * term[t]  (n elements); the dense-output / interpolation arithmetic of the solvers. */
int grand_lincomb(int64_t n, int32_t n_terms, const float* const* term, const float* coef, float* y_out,
                  void* stream);

/* dopri5 error ratio: sum over elements of (err / (atol + rtol max(|y0|,|y1|)))^2 accumulated in double into
 * *sumsq (caller zeroes it), err = sum_t coef[t] * k[t]   (torchdiffeq _compute_error_ratio with rms norm). */
int grand_error_sumsq(int64_t n, int32_t n_terms, const float* const* k, const float* coef, const float* y0,
                      const float* y1, float atol, float rtol, double* sumsq, void* stream);

/*
 * dopri5 with the controller ON THE DEVICE (torchdiffeq RKAdaptiveStepsizeODESolver semantics, witness reference
 * src/early_stop_solver.py:30-86): initial step selection, the attempt loop, accept / reject, the step-size update and
 * the final dense-output evaluation at t1 are ONE CUDA Graph whose attempt loop is a conditional WHILE node; the host
 * launches it and reads nothing back until it wants the counters.  Integrates y(t0) -> y(t1) for the two-point time
 * vector the reference's blocks use.  All buffers caller-owned: `planes` = 13 x [n_rows, d] floats (ycur, ya, yb, ymid,
 * yprev, fprev, k0..k6), `state` = device memory of sizeof(grand_dopri5_state); y0 is read, `out` receives y(t1).
 * a == null selects the Laplacian function with w_csr.  create / destroy are the only calls that allocate (the graph).
 */
#define GRAND_DOPRI5_COEFS 72
typedef struct grand_dopri5_state {
  double t, t_prev, dt, t_end;   /* time of ycur, start of the last accepted step, next step size, target        */
  double acc[3];                 /* sum-of-squares accumulators of the norm kernels                              */
  double h0;                     /* probe size of the initial step selection                                     */
  float rtol, atol;
  int32_t numel, attempts, accepted, nfe, max_attempts, accept, have_interp, status;  /* status != 0: dt underflow */
  float coef[GRAND_DOPRI5_COEFS];/* every dt-dependent coefficient of the current attempt (stage combines, error, dense output) */
} grand_dopri5_state;
typedef struct grand_dopri5_plan grand_dopri5_plan;
int grand_dopri5_create(const grand_graph* g, const grand_attention* a, const float* w_csr, int32_t d,
                        const grand_workspace* ws, const grand_epilogue* residual, const float* y0, float* out,
                        float* planes, grand_dopri5_state* state, double t0, double t1, float rtol, float atol,
                        int32_t max_attempts, grand_dopri5_plan** plan);
int grand_dopri5_launch(grand_dopri5_plan* plan, void* stream);
int grand_dopri5_destroy(grand_dopri5_plan* plan);

/* ---- node-partitioned multi-GPU over peer-mapped memory (see grand_peers) ------------------------------------------
 * Arena life cycle (the only entry points of this library that allocate): cudaMalloc / cudaIpcGetMemHandle /
 * cudaIpcOpenMemHandle / cudaIpcCloseMemHandle / cudaFree.  `handle` is the 64-byte cudaIpcMemHandle_t. */
int grand_peer_alloc(int64_t bytes, void** ptr);
int grand_peer_free(void* ptr);
int grand_peer_export(void* ptr, void* handle64);
int grand_peer_open(const void* handle64, void** ptr);
int grand_peer_close(void* ptr);
/* Stores the own rows of `src` ([n_rows, width] inside the own arena, width % 4 == 0) into the halo sections of the
 * peers that reference them (push lists of `peers`); one thread per (push entry, 16-byte piece), so every delivery is
 * an independent load -> store.  Used for the K rows after the projection and for the first stage input. */
int grand_peer_push_rows(const grand_peers* peers, const float* src, int32_t n_rows, int32_t width, void* stream);
/* All ranks of `peers` wait for each other ON THE STREAM: bumps the local epoch, publishes it into every peer's flags
 * (release, system scope) and spins until every peer's epoch has arrived (acquire).  Everything the peers stored
 * before their barrier is visible afterwards.  A peer that does not arrive within ~2 s sets `error` instead of hanging
 * the GPU (grand_peer_error reads it back after a synchronize). */
int grand_peer_barrier(const grand_peers* peers, void* stream);
/* HOST-synchronous read of the barrier error flag (0 = every barrier so far completed). */
int grand_peer_error(const grand_peers* peers, int32_t* error_out);
/* grand_integrate_fixed on a node partition: per evaluation  project own rows -> push K rows -> barrier -> edge kernel
 * (its epilogue pushes the next stage's rows).  y / yb / yc and k_bufs[0..1] ([n_cols, d] / [n_cols, ad]) lie in the
 * arena; k1..k3 ([n_rows, d]) are local.  *k_toggle selects the K buffer of the first evaluation and is updated. */
int grand_integrate_fixed_sharded(const grand_graph* g, const grand_attention* a, const grand_workspace* ws,
                                  const grand_epilogue* residual, int32_t method, const float* dt, int32_t n_steps,
                                  float* y, float* yb, float* yc, float* k1, float* k2, float* k3, float* const* k_bufs,
                                  int32_t* k_toggle, const grand_peers* peers, void* stream);

/* HOST function (plain CPU code, no CUDA call): the sequential step of building a grand_graph -- greedy packing of
 * consecutive rows with <= tile in-edges into windows of <= tile edges and <= max_rows (<= GRAND_MAX_PACKED_ROWS) rows;
 * rows with more edges are skipped (they are cut into chunks) and close the current window.  deg[n_rows] are the row
 * degrees; first[w] / count[w] (capacity n_rows each) receive the first row and the row count of window w.  Returns the
 * number of windows or a negative error code.  The reference has no counterpart (torch_sparse.spmm takes raw COO). */
int64_t grand_plan_pack_rows(const int64_t* deg, int64_t n_rows, int32_t tile, int32_t max_rows, int64_t* first,
                             int64_t* count);

/* ---- the graph plan, natively (csr_build / plan_create / plan_destroy) -------------------------
 * HOST functions (plain CPU code, no CUDA call).  grand_plan_create re-lays the caller's COO edge list
 * (edge_index [2, n_edges], int64, host memory; what the reference hands to torch_sparse.spmm / PyG softmax at
 * src/function_transformer_attention.py:25-36, 190-213) into everything `grand_graph` points at: CSR by edge_index[0]
 * (stable), windows of <= GRAND_TILE edges (chunks of long rows first, longest row first, then packed short rows), the
 * per-lane records, the CSC view, and the work list of the persistent edge kernel.  The result is ONE int32 image;
 * grand_plan_describe gives its length, the scalar counts and the offset / element count of each array in the order of
 * enum grand_plan_array; the caller copies the image to the device and grand_plan_bind fills a grand_graph whose pointers
 * point into that copy.  The hard-attention block rebuilds its plan through this on every training forward
 * (reference src/block_transformer_hard_attention.py:60-90 re-selects the edges there). */
enum grand_plan_array {
  GRAND_PLAN_ROWPTR = 0, GRAND_PLAN_COL, GRAND_PLAN_EROW, GRAND_PLAN_EID, GRAND_PLAN_WIN_HDR, GRAND_PLAN_WIN_REC,
  GRAND_PLAN_SPLIT_TILE0, GRAND_PLAN_COLPTR, GRAND_PLAN_CSC_POS, GRAND_PLAN_SEG_DESC, GRAND_PLAN_LROW_SLOT0,
  GRAND_PLAN_LROW_ROW, GRAND_PLAN_TILE_ROW, GRAND_PLAN_TILE_BEG, GRAND_PLAN_TILE_NROWS, GRAND_PLAN_TILE_SPLIT,
  GRAND_PLAN_ARRAYS
};
typedef struct grand_plan_host grand_plan_host;
int grand_plan_create(const int64_t* edge_index, int64_t n_edges, int32_t n_rows, int32_t n_cols, int32_t tile,
                      int32_t max_rows, int32_t seg_edges, grand_plan_host** out);
/* any output pointer may be null; offsets / elems have GRAND_PLAN_ARRAYS entries (int32 elements) */
int grand_plan_describe(const grand_plan_host* plan, grand_graph* counts, int32_t* max_degree, int64_t* image_elems,
                        int64_t* offsets, int64_t* elems);
const int32_t* grand_plan_image(const grand_plan_host* plan);
int grand_plan_bind(const grand_plan_host* plan, const int32_t* image, grand_graph* g);
void grand_plan_destroy(grand_plan_host* plan);
#ifdef __cplusplus
}
#endif
#endif /* GRAND_B200_H_ */
