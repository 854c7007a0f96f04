/*
 * pgsd_b200 -- C ABI of the B200-native signed/directed message-passing hot path.
 *
 * The reference (PyGSD 1.1.1, pure Python) has no FFI of its own: the boundary its models,
 * examples and tests call is the Python class API (MagNetConv.forward, DiGCNConv.forward,
 * SGCNConv.forward, Conv_Base.forward, DIMPA.forward).  The drop-in Python classes in
 * `pytorch_geometric_signed_directed_b200/nn/` keep those signatures and bind the entry
 * points below through ctypes (see INTEGRATION.md for the stub a reference maintainer
 * would add).  Each entry point cites the reference code it replaces; paths are relative to
 * /root/reference/torch_geometric_signed_directed/.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in `_host`;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it;
 *     the plan builders additionally synchronise that stream once to return `nnz`;
 *   - functions return 0 (PGSD_OK) or a PGSD_ERR_* code; pgsd_last_error() returns the
 *     message for the calling thread; nothing throws across the ABI;
 *   - the library never allocates or frees caller memory: scratch space is sized by the
 *     *_workspace_bytes queries and passed in;
 *   - indices inside plans are int32 (N, nnz < 2^31); user edge lists are int64 as in PyG.
 */
#ifndef PGSD_B200_H
#define PGSD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PGSD_ABI_VERSION 1

#if defined(__GNUC__)
#define PGSD_API __attribute__((visibility("default")))
#else
#define PGSD_API
#endif

#define PGSD_OK 0
#define PGSD_ERR_INVALID 1     /* bad argument (null pointer, negative size, unsupported F) */
#define PGSD_ERR_CUDA 2        /* a CUDA runtime call or kernel launch failed              */
#define PGSD_ERR_WORKSPACE 3   /* workspace too small                                      */
#define PGSD_ERR_RANGE 4       /* N or nnz exceeds the int32 plan format                   */

typedef void* pgsd_stream_t;

#define PGSD_F32 0
#define PGSD_BF16 1

PGSD_API int pgsd_abi_version(void);
PGSD_API const char* pgsd_last_error(void);
/* sm count / compute capability of the current device (cc must be 10.x for the kernels). */
PGSD_API int pgsd_device_info(int* sm_count_host, int* cc_major_host, int* cc_minor_host);
/* sizeof(pgsd_spmm_args) / sizeof(pgsd_dense_args) as compiled: lets a foreign-language binding
 * verify its struct mirror before the first call. */
PGSD_API int pgsd_sizeof_args(size_t* spmm_args_bytes_host, size_t* dense_args_bytes_host);

/* ------------------------------------------------------------------------------------
 * Plan builders: user COO (int64) -> CSR-by-destination (int32), built once per graph and
 * cached by the Python layer under the reference's own cache rules.
 * ---------------------------------------------------------------------------------- */

/* Scratch bytes sufficient for any builder below on (num_nodes, num_edges). */
PGSD_API int pgsd_plan_workspace_bytes(int64_t num_nodes, int64_t num_edges, size_t* bytes_host);

/* Generic aggregation plan: out[dst] (+)= w * x[src]; duplicates and self-loops are kept as
 * separate entries, entries of a row stay in edge order (stable).
 * Replaces the index plumbing of PyG MessagePassing.propagate as used by
 *   nn/directed/DiGCNConv.py:86-89 (weighted add, dst = edge_index[1]),
 *   nn/signed/SGCNConv.py:101-118,128-129 (unweighted mean, pass edge_weight = NULL).
 * row_ptr [num_dst+1], col [E], val [E] (ignored when edge_weight is NULL). */
PGSD_API int pgsd_build_csr(const int64_t* edge_src, const int64_t* edge_dst, const float* edge_weight,
                   int64_t num_edges, int64_t num_dst, int64_t num_src,
                   int32_t* row_ptr, int32_t* col, float* val,
                   void* workspace, size_t workspace_bytes, pgsd_stream_t stream);

/* Random-walk normalised plan with "remaining" self loops:
 *   A_hat = A(off-diagonal) + diag(existing self-loop weight, else fill_value);
 *   w' = w / rowsum_dst(A_hat);   out[dst] = diag[dst]*x[dst] + sum w' * x[src].
 * Replaces conv_norm_rw + Conv_Base.message, nn/general/conv_base.py:12-31,98-117
 * (dst = edge_index[0], src = edge_index[1]; DIMPA's transposed call swaps them,
 * nn/directed/DIMPA.py:50-53).  edge_weight may be NULL (ones).
 * add_self_loops = 0 reproduces conv_norm_rw(add_self_loops=False): self-loop edges stay
 * ordinary entries, diag = 0, fill_value is ignored.
 * row_ptr [N+1], col/val [E] (capacity), diag [N]; *nnz_host = stored entries. */
PGSD_API int pgsd_build_csr_rw_norm(const int64_t* edge_dst, const int64_t* edge_src,
                           const float* edge_weight, int64_t num_edges, int64_t num_nodes,
                           float fill_value, int add_self_loops, int32_t* row_ptr,
                           int32_t* col, float* val, float* diag, int64_t* nnz_host,
                           void* workspace, size_t workspace_bytes, pgsd_stream_t stream);

/* Symmetrically normalised plan with "remaining" self loops (PyG gcn_norm as called by
 * nn/directed/DGCNConv.py:75-77): deg = rowsum_dst(A_hat), w' = deg^-1/2[src] * w * deg^-1/2[dst],
 * diag = loop weight / deg.  Same conventions as pgsd_build_csr_rw_norm. */
PGSD_API int pgsd_build_csr_sym_norm(const int64_t* edge_dst, const int64_t* edge_src,
                            const float* edge_weight, int64_t num_edges, int64_t num_nodes,
                            float fill_value, int add_self_loops, int32_t* row_ptr,
                            int32_t* col, float* val, float* diag, int64_t* nnz_host,
                            void* workspace, size_t workspace_bytes, pgsd_stream_t stream);

/* Magnetic (Hermitian) Laplacian plan, scaled for the Chebyshev recurrence:
 *   L~ = 2 L / lambda_max - I,  L = I - D^-1/2 A_s D^-1/2 (.) exp(i 2 pi q Theta)   ('sym')
 *                               L = D - A_s (.) exp(i 2 pi q Theta)                (none)
 * with A_s = (A + A^T)/2, Theta = A - A^T, self-loops dropped, duplicate edges summed.
 * Replaces utils/directed/get_magnetic_Laplacian.py:44-87 + MagNetConv.__norm__
 * (nn/directed/MagNetConv.py:78-120); signed_mode 1/2 replaces
 * utils/general/get_magnetic_signed_Laplacian.py:45-92 (absolute_degree True/False) +
 * nn/general/MSConv.py:76-119.
 * The plan is stored for the direction the reference actually aggregates in
 * (source_to_target): row a lists sources b with value L~[b,a]; entries are in
 * (a,b)-sorted order, which is also the order of the reference's coalesced list, so
 * `cached_result` can be re-materialised from it (indices bit-exact).
 *   val_real/val_imag [2E] capacity, diag_real [N] (= 2*diag(L)/lambda_max - 1; the
 *   imaginary diagonal is identically 0).  normalization: 1 = 'sym', 0 = None. */
PGSD_API int pgsd_build_magnetic_laplacian(const int64_t* edge_row, const int64_t* edge_col,
                                  const float* edge_weight, int64_t num_edges,
                                  int64_t num_nodes, double q, int normalization,
                                  float lambda_max, int signed_mode,
                                  int32_t* row_ptr, int32_t* col, float* val_real,
                                  float* val_imag, float* diag_real, int64_t* nnz_host,
                                  void* workspace, size_t workspace_bytes,
                                  pgsd_stream_t stream);

/* Same plan, additionally keeping theta[u] = the coalesced antisymmetric weight Theta of every
 * stored entry (plan orientation: val_real = -m cos(2 pi q theta), val_imag = m sin(2 pi q theta)),
 * which pgsd_magnetic_q_grad needs when q is trainable (MagNetConv.py:58-59,141-142). */
PGSD_API int pgsd_build_magnetic_laplacian_theta(const int64_t* edge_row, const int64_t* edge_col,
                                  const float* edge_weight, int64_t num_edges,
                                  int64_t num_nodes, double q, int normalization,
                                  float lambda_max, int signed_mode,
                                  int32_t* row_ptr, int32_t* col, float* val_real,
                                  float* val_imag, float* diag_real, float* theta,
                                  int64_t* nnz_host, void* workspace, size_t workspace_bytes,
                                  pgsd_stream_t stream);

/* Row-range build of the same plan for the node-range sharded path (no reference
 * counterpart -- the reference is single-process): rank r builds only rows [row_lo, row_hi) from the
 * edges incident to that node range (edges touching no node of the range are ignored; passing the
 * whole list is allowed), in two phases around ONE all-gather:
 *   begin : row_ptr [row_hi-row_lo+1], col / sym / theta [2E capacity] (sym = coalesced A+A^T weight,
 *           theta = coalesced A-A^T of each stored entry, same (a,b)-sorted order and the same
 *           in-edge-order duplicate sums as the full build) and deg_local [row_hi-row_lo] = row sums;
 *   (caller all-gathers deg_local into deg_all [num_nodes])
 *   finish: val_real / val_imag [nnz] (may alias sym / theta) and diag_real [row_hi-row_lo].
 * Rows produced this way are bit-identical to the same rows of pgsd_build_magnetic_laplacian. */
PGSD_API int pgsd_build_magnetic_rows_begin(const int64_t* edge_row, const int64_t* edge_col,
                                  const float* edge_weight, int64_t num_edges, int64_t num_nodes,
                                  int64_t row_lo, int64_t row_hi, int signed_mode,
                                  int32_t* row_ptr, int32_t* col, float* sym, float* theta,
                                  float* deg_local, int64_t* nnz_host, void* workspace,
                                  size_t workspace_bytes, pgsd_stream_t stream);
PGSD_API int pgsd_build_magnetic_rows_finish(const int32_t* row_ptr, const int32_t* col, const float* sym,
                                  const float* theta, const float* deg_all, int64_t num_nodes,
                                  int64_t row_lo, int64_t row_hi, double q, int normalization,
                                  float lambda_max, float* val_real, float* val_imag,
                                  float* diag_real, pgsd_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Sparse aggregation (the hot loop): for every destination row r and operator k < n_ops
 *   agg_k[r] = diag_k[r] * x_k[r] + sum_{e in row r} val_k[e] * x_k[col[e]]
 *   agg_k[r] *= 1 / max(row_len(r), 1)                         if mean
 *   y_k[r]   = alpha * agg_k[r] + beta * z_k[r] + bias         (z_k, bias optional)
 * The n_ops (1 or 2) operators share (row_ptr, col) and are evaluated in one pass.
 * Replaces PyG propagate = index_select + message + scatter_add as called from
 *   nn/directed/MagNetConv.py:196-236,251-252 (n_ops = 2: real & imaginary operator;
 *     alpha = 2, beta = -1, z = T_{k-2} fuses the Chebyshev recurrence of :214-216),
 *   nn/directed/DiGCNConv.py:86-94 (bias = update()),
 *   nn/signed/SGCNConv.py:101-118 (mean, val = NULL),
 *   nn/general/conv_base.py:111-117.
 * x/z/y are row-major with leading dimensions in ELEMENTS so column slices of wider
 * buffers can be aggregated/written in place (SGCNConv.py:109-119 half-width slices).
 * n_ops = 2 with x[0] == x[1] and ldx[0] == ldx[1] (fp32; how examples/magnet_node.py:61-62
 * calls the first layer: X_real and X_img are one tensor) gathers every neighbour row ONCE for
 * both operators; results are bit-identical to the two-gather evaluation.  z[k] may alias
 * y[k] (accumulating epilogue: every row is read and written by the same lanes).
 * ---------------------------------------------------------------------------------- */
typedef struct pgsd_spmm_args {
  int64_t n_rows;            /* destination rows                                      */
  int32_t feat;              /* F: feature columns aggregated                         */
  int32_t n_ops;             /* 1 or 2                                                */
  int32_t dtype;             /* PGSD_F32 or PGSD_BF16 (x, z, y); accumulation is fp32 */
  int32_t mean;              /* bit 0: divide by max(row length, 1); bit 1: tanh of the finished row
                                (after beta * z and bias; hub rows excluded)           */
  const int32_t* row_ptr;    /* [n_rows + 1]                                          */
  const int32_t* col;        /* [nnz]                                                 */
  const float* val[2];       /* [nnz] or NULL (implicit 1.0)                          */
  const float* diag[2];      /* [n_rows] or NULL (use diag_const)                     */
  float diag_const[2];       /* 0 disables the diagonal term                          */
  const void* x[2];          /* [n_src, ldx] gathered features                        */
  int64_t ldx[2];
  float alpha, beta;
  const void* z[2];          /* [n_rows, ldz] or NULL                                 */
  int64_t ldz[2];
  void* y[2];                /* [n_rows, ldy]                                         */
  int64_t ldy[2];
  const float* bias;         /* [F] or NULL                                           */
  int32_t variant;           /* 0 = library default; bits 0-3 loads in flight (2/4/8), 0x10 /
                                0x20 prefer 128- / 256-bit gathers, 0x80 warp-per-row kernel
                                instead of the default group-per-row kernel, 0x400 gather twice
                                even when both operators read one tensor (A/B timing); 0x800 bulk-copy (TMA)
                                gathers + segmented reduction (fp32 rows of 128/256/512 bytes); bits 12-14:
                                preferred shared-memory carve-out of the launch in steps of 14 % of
                                the SM's 228 KB (0 = driver default), so that a kernel which needs
                                shared memory (bulk-copy shard push) can be co-resident            */
  int32_t diag_row_offset;   /* x row holding destination row 0 (diag term only): lets x span a
                                larger node range than the plan's rows (row-sharded plans)  */
  float op_scale[2];         /* per-operator multiplier of alpha (0 is read as 1): -1 on the
                                antisymmetric imaginary operator gives the transposed
                                aggregation of the backward pass from the same plan        */
  /* hub rows (optional; fp32 only): rows longer than long_row_threshold are listed in
   * long_rows[n_long_rows]; their entries are aggregated in long_chunk-entry slices by a
   * second launch (fp32 atomics), long_chunk_ptr[n_long_rows + 1] = exclusive prefix of the
   * slices per hub row.  Leave n_long_rows = 0 to process every row in the main kernel.       */
  const int32_t* long_rows;
  const int32_t* long_chunk_ptr;
  int32_t n_long_rows;
  int32_t long_row_threshold;
  int32_t long_chunk;
  int32_t grid_reserve;      /* resident-CTA slots to leave free for a kernel running beside this
                                one (the shard-push collective of the row-sharded path); 0 = none */
} pgsd_spmm_args;

PGSD_API int pgsd_spmm_csr(const pgsd_spmm_args* args, pgsd_stream_t stream);
/* Name of the kernel template instantiation the calling thread's last pgsd_spmm_csr launched (e.g.
 * "spmm_groups_kernel<4,16,2,2,4,0,256,3>" = <words per load, lanes per row, operators, gathered matrices, loads in
 * flight, bf16, threads, min CTAs per SM>): lets a harness check that a recorded profile belongs to the kernel it times. */
PGSD_API const char* pgsd_last_spmm_kernel(void);

/* ------------------------------------------------------------------------------------
 * Dense feature transform next to the aggregation:
 *   acc_g[r, :] = sum_{t : group[t] == g} X_t[r, :k_t] @ W_t          (W_t is [k_t, n_out])
 *   combine == 0:  y0 = acc_0 + bias
 *   combine == 1:  y0 = acc_0 - acc_1 + bias ;  y1 = acc_0 + acc_1 + bias
 * combine == 1 is MagNetConv's real/imag mixing: out_real = A - B + b, out_imag = A + B + b
 * with A = sum_k T_k(L~_r) x_real W_k (group 0), B = sum_k T_k(L~_i) x_imag W_k (group 1)
 * (nn/directed/MagNetConv.py:189-247; two of its four chains are duplicates).
 * combine == 0 replaces torch.matmul / nn.Linear at nn/directed/DiGCNConv.py:66,
 * nn/directed/DiGCN_Inception_Block.py:44, nn/signed/SGCNConv.py:102-121 (terms = the
 * column blocks of the concatenated input, so no torch.cat is materialised).
 * W_t element (k, n) is w[t][k * ldw_k[t] + n * ldw_n[t]] which covers both the
 * [in, out] Parameter layout and nn.Linear's [out, in] layout (and column slices of it).
 * fp32 in / fp32 accumulate; PGSD_BF16 reads/writes bf16 X/Y with fp32 weights & accumulate.
 * ---------------------------------------------------------------------------------- */
#define PGSD_DENSE_MAX_TERMS 16
typedef struct pgsd_dense_args {
  int64_t n_rows;
  int32_t n_out;
  int32_t n_terms;
  int32_t dtype;
  int32_t combine;
  const void* x[PGSD_DENSE_MAX_TERMS];
  int64_t ldx[PGSD_DENSE_MAX_TERMS];
  int32_t k[PGSD_DENSE_MAX_TERMS];
  int32_t group[PGSD_DENSE_MAX_TERMS];
  const float* w[PGSD_DENSE_MAX_TERMS];
  int64_t ldw_k[PGSD_DENSE_MAX_TERMS];
  int64_t ldw_n[PGSD_DENSE_MAX_TERMS];
  const float* bias;         /* [n_out] or NULL */
  void* y[2];
  int64_t ldy[2];
  int32_t relu_mode;         /* 0 none; 1 complex ReLU (mask = y0 >= 0 applied to y0, y1):
                                nn/directed/complex_relu.py:17-34 (combine == 1 only);
                                2 tanh (combine == 0 only): nn/signed/SGCN.py:93-96     */
  int32_t variant;           /* 0 = auto: TMA-fed warp-specialised tcgen05 kernel, else the
                                register-staged tcgen05 kernel, else FFMA; 1 = FFMA; 2 / 4 / 8 =
                                register-staged tcgen05 (synchronous / warp-specialised / deep
                                prefetch); 16 = require the TMA-fed kernel (experiment knobs of that
                                kernel ride in the upper bits: 8-11 lo slots, 12-15 landing stages,
                                16-18 timing-only knock-outs, 20-22 converter groups, 24-25 epilogue
                                staging layout; 0 everywhere = measured defaults)            */
} pgsd_dense_args;

PGSD_API int pgsd_dense_transform(const pgsd_dense_args* args, pgsd_stream_t stream);

/* ------------------------------------------------------------------------------------
 * One MagNetConv / MSConv layer of Chebyshev order K = 1 in ONE launch (SURVEY 8b
 * "magnet_layer_fused_f32"): aggregation and transform fused, so T = L~ x never goes to HBM.
 *   T_k[r]   = diag_k[r] * x_k[r] + sum_{e in row r} val_k[e] * x_k[col[e]]     k = real, imag
 *   A        = x_real W0 + T_real W1,   B = x_imag W0 + T_imag W1
 *   out_real = A - B + bias,  out_imag = A + B + bias       (+ complex ReLU mask when relu_mode = 1)
 * Replaces the whole body of MagNetConv.forward for weight [2, F_in, F_out]
 * (nn/directed/MagNetConv.py:185-249: four propagates, eight matmuls, the real/imag mix and the
 * bias add; nn/general/MSConv.py:181-246 likewise).  Persistent CTAs: lane groups aggregate
 * rows into a shared-memory ring, a consumer warpgroup splits finished 128-row tiles into TF32
 * hi/lo operands and runs 3xTF32 tcgen05.mma against the resident weights, epilogue from TMEM.
 * fp32 only; feat_in = 64, feat_out = 64; plans without hub rows (see pgsd_spmm_args).
 * pgsd_magnet_fused_supported() returns 1 when the shape is inside that envelope.
 * ---------------------------------------------------------------------------------- */
typedef struct pgsd_magnet_fused_args {
  int64_t n_rows;            /* N (square operator: destination rows == source rows)   */
  int32_t feat_in, feat_out;
  const int32_t* row_ptr;    /* [N + 1]                                                */
  const int32_t* col;        /* [nnz]                                                  */
  const float* val[2];       /* real / imaginary operator values                       */
  const float* diag[2];      /* [N] or NULL (use diag_const)                           */
  float diag_const[2];
  const float* x[2];         /* x_real, x_imag [N, ldx]                                */
  int64_t ldx[2];
  const float* w[2];         /* W0, W1: element (k, n) at w[i][k * ldw_k[i] + n * ldw_n[i]] */
  int64_t ldw_k[2], ldw_n[2];
  const float* bias;         /* [feat_out] or NULL                                     */
  float* y[2];               /* out_real, out_imag [N, ldy]                            */
  int64_t ldy[2];
  int32_t relu_mode;         /* as pgsd_dense_args                                     */
  int32_t variant;           /* producer warps x gathers in flight: 0 = 16x4, 1 = 20x4, 2 = 20x2, 3 = 24x2 */
} pgsd_magnet_fused_args;

PGSD_API int pgsd_magnet_fused_supported(int32_t feat_in, int32_t feat_out, int32_t dtype);
/* sizeof(pgsd_magnet_fused_args) as compiled (struct-mirror check of foreign bindings). */
PGSD_API size_t pgsd_sizeof_magnet_fused_args(void);
PGSD_API int pgsd_magnet_layer_fused(const pgsd_magnet_fused_args* args, pgsd_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Segment-softmax attention over CSR-by-destination plans.
 *   t_e     = act(s_src[p][j] + s_dst[p][i])        edge e = (j -> i) of type p (one plan per type)
 *   alpha_e = exp(t_e - max_i) / (sum over ALL entries of row i, both types, + 1e-16)
 * Replaces SNEAConv.message + PyG utils.softmax (nn/signed/SNEAConv.py:135-146): the
 * reference's Linear(2*out -> 1) on [x_j || x_i] splits into two per-node scalars
 * (s_src = X a_j, s_dst = X a_i + c), so the per-edge work is scalar.
 *   y != NULL        : y[i] = xd[0][i] * sum_{type 0} alpha + xd[1][i] * sum_{type 1} alpha
 *                      (SNEAConv aggregates the TARGET feature times alpha)
 *   alpha_out[p] set : alpha_e written per stored entry (GATConv-style layers then run
 *                      pgsd_spmm_csr with val = alpha_out[p]).
 * act: 0 = tanh, 1 = leaky_relu(slope).  fp32 only.
 * ---------------------------------------------------------------------------------- */
typedef struct pgsd_attn_args {
  int64_t n_rows;
  int32_t feat;
  int32_t n_types;           /* 1 or 2 */
  int32_t act;
  float slope;
  const int32_t* row_ptr[2];
  const int32_t* col[2];
  const float* s_src[2];     /* [n_src] */
  const float* s_dst[2];     /* [n_rows] */
  const float* xd[2];        /* [n_rows, ldxd] target-side features (mode y) */
  int64_t ldxd[2];
  float* y;                  /* [n_rows, ldy] or NULL */
  int64_t ldy;
  float* alpha_out[2];       /* [nnz_p] or NULL */
} pgsd_attn_args;

PGSD_API int pgsd_edge_softmax(const pgsd_attn_args* args, pgsd_stream_t stream);

/* Backward of pgsd_edge_softmax (training through SNEAConv.message, nn/signed/SNEAConv.py:135-146, and through the
 * attention of PyG GATConv as used by nn/signed/SDGNN.py:35-64).  dL/dalpha arrives per stored entry (dalpha[p],
 * GAT-style: see pgsd_sddmm_rows) or, when dalpha[p] is NULL, as one coefficient per row (row_coef[p], SNEAConv:
 * <gy[i], xd_p[i]>).  Writes g_s_dst[p] [n_rows]; ADDS into g_s_src[p] [n_src] (caller zero-initialises; fp32
 * atomics); type_sum[p] (optional, [n_rows]) receives the per-row sum of alpha over the entries of type p. */
typedef struct pgsd_attn_bwd_args {
  int64_t n_rows;
  int32_t n_types, act;
  float slope;
  int32_t reserved;
  const int32_t* row_ptr[2];
  const int32_t* col[2];
  const float* s_src[2];
  const float* s_dst[2];
  const float* dalpha[2];
  const float* row_coef[2];
  float* g_s_src[2];
  float* g_s_dst[2];
  float* type_sum[2];
} pgsd_attn_bwd_args;
PGSD_API size_t pgsd_sizeof_attn_bwd_args(void);
PGSD_API int pgsd_edge_softmax_backward(const pgsd_attn_bwd_args* args, pgsd_stream_t stream);

/* out[k] = <gy[row(k)], h[col[k]]> for every stored entry k of a CSR plan: dL/dalpha of y[i] = sum_k alpha_k h[col_k]
 * (sampled dense-dense product).  fp32, feat % 4 == 0 (<= 256), 16-byte aligned rows. */
PGSD_API int pgsd_sddmm_rows(const int32_t* row_ptr, const int32_t* col, const float* gy, int64_t ldg,
                             const float* h, int64_t ldh, int64_t n_rows, int32_t feat, float* out,
                             pgsd_stream_t stream);

/* GATConv aggregation with the edge softmax inside (PyG GATConv, heads = 1, as used by nn/signed/SDGNN.py:35-61 and
 * nn/signed/SiGAT.py:59-64):  y[i] = sum_e alpha_e h[col_e] (+ bias) (+ beta * z[i]) over the entries of row i,
 * alpha = softmax_i(leaky_relu(s_src[col_e] + s_dst[i])) with PyG's 1e-16 in the denominator.  fp32, feat a
 * multiple of 4 (<= 128), 16-byte aligned rows; z may alias y (accumulating epilogue). */
PGSD_API int pgsd_gat_aggregate(const int32_t* row_ptr, const int32_t* col, const float* s_src,
                                  const float* s_dst, float negative_slope, const float* h, int64_t ldh,
                                  int32_t feat, int64_t n_rows, const float* bias, const float* z,
                                  int64_t ldz, float beta, float* y, int64_t ldy, pgsd_stream_t stream);

/* Weight / bias gradient of y = X W (backward pass):
 *   dw[k, n] += sum_r x[r, k] * g[r, n];   db[n] += sum_r g[r, n]   (db may be NULL)
 * x: [n_rows, k], g: [n_rows, n] (fp32 or bf16), dw/db fp32 accumulated with atomics: the
 * caller zero-initialises them.  Replaces autograd's matmul backward for the call sites listed
 * under pgsd_dense_transform. */
PGSD_API int pgsd_xtg_accumulate(const void* x, int64_t ldx, const void* g, int64_t ldg, int64_t n_rows,
                                 int32_t k, int32_t n, int32_t dtype, float* dw, int64_t lddw, float* db,
                                 pgsd_stream_t stream);

/* Gradient of the loss w.r.t. a trainable magnetic charge q through ONE aggregation
 *   y_real = alpha * L~_r^T x_real,  y_imag = alpha * L~_i^T x_imag   (pgsd_spmm_csr, n_ops = 2):
 *   *dq += scale * sum_e theta[e] * ( val_imag[e] * <gy_real[row(e)], x_real[col[e]]>
 *                                   - val_real[e] * <gy_imag[row(e)], x_imag[col[e]]> ),
 * scale = 2 pi alpha  (d val_real/dq = 2 pi theta val_imag, d val_imag/dq = -2 pi theta val_real).
 * A sampled dense-dense product (SDDMM) reduced to a scalar; fp32 features, double accumulator the
 * caller zero-initialises.  Replaces autograd through torch.exp(1j*2*pi*q*theta)
 * (utils/directed/get_magnetic_Laplacian.py:68) + the four propagates of MagNetConv.py:196-236. */
PGSD_API int pgsd_magnetic_q_grad(const int32_t* row_ptr, const int32_t* col, const float* val_real,
                                  const float* val_imag, const float* theta, int64_t n_rows, int32_t feat,
                                  const float* gy_real, int64_t ldgr, const float* gy_imag, int64_t ldgi,
                                  const float* x_real, int64_t ldxr, const float* x_imag, int64_t ldxi,
                                  double scale, double* dq, pgsd_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Sparse preprocessing primitives.  The reference prepares DiGCN / DGCN inputs with
 * DENSE N x N matrices on the CPU (a dense (N+1)^2 eigen-decomposition and dense products):
 *   utils/directed/get_adjs_DiGCN.py:113-190 (get_appr_directed_adj), :193-254 (get_second_directed_adj),
 *   utils/directed/features_in_out.py:44-46 (directed_features_in_out).
 * pytorch_geometric_signed_directed_b200/utils/directed.py composes the same operators from:
 * ---------------------------------------------------------------------------------- */

/* Sort COO entries by key (= row * n + col, < 2^key_bits) and sum the values of equal keys
 * (torch.sparse_coo_tensor(...).to_dense() / torch.nonzero order of the reference: row-major).
 * keys_out / vals_out need capacity n_entries; *n_unique_host = entries written. */
PGSD_API int pgsd_coalesce_workspace_bytes(int64_t n_entries, size_t* bytes_host);
PGSD_API int pgsd_coo_coalesce(const int64_t* keys, const float* vals, int64_t n_entries, int key_bits,
                               int64_t* keys_out, float* vals_out, int64_t* n_unique_host, void* workspace,
                               size_t workspace_bytes, pgsd_stream_t stream);

/* Expand step of C = B^T diag(scale) B for a CSR matrix B [n_rows x n_cols]: row k with len_k entries
 * emits its len_k^2 products (key = col[a] * n_cols + col[b], value = scale[k] * val[a] * val[b]);
 * product_offsets [n_rows + 1] = exclusive prefix of len_k^2 (int64), n_products = its last element.
 * Followed by pgsd_coo_coalesce this is an expand-sort-compress SpGEMM: P^T P and P P^T of
 * get_adjs_DiGCN.py:231-232, A^T D^-1 A and A D^-1 A^T of features_in_out.py:44-46.
 * val / scale may be NULL (ones). */
PGSD_API int pgsd_gram_expand(const int32_t* row_ptr, const int32_t* col, const float* val, const float* scale,
                              const int64_t* product_offsets, int64_t n_rows, int64_t n_cols,
                              int64_t n_products, int64_t* keys_out, float* vals_out, pgsd_stream_t stream);

/* Stationary vector of the personalised-PageRank chain of get_adjs_DiGCN.py:147-160 (the dense
 * scipy.linalg.eig of an (N+1) x (N+1) matrix) by sparse fp64 power iteration:
 *   pi = (1 - alpha) P^T pi + alpha / ((1 + alpha) N),   pi <- pi / sum(pi) is left to the caller.
 * (row_ptr_dst, col_src, p) is P in CSR by DESTINATION (row j lists sources i with p[i -> j]).
 * pi, scratch: [n] doubles.  The error contracts by (1 - alpha) per iteration. */
PGSD_API int pgsd_ppr_stationary(const int32_t* row_ptr_dst, const int32_t* col_src, const float* p,
                                 int64_t n, double alpha, int32_t n_iter, double* pi, double* scratch,
                                 pgsd_stream_t stream);

/* Halo pack for the node-range sharded path (no reference counterpart: the reference is
 * single-device): out[i, :] = x[index[i], :]  -- rows another rank asked for. */
PGSD_API int pgsd_gather_rows(const void* x, int64_t ldx, const int32_t* index, int64_t n_index,
                     int32_t feat, int32_t dtype, void* out, int64_t ldo,
                     pgsd_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Shard push: the all-gather of the node-range sharded path as ONE kernel over
 * NVLink peer memory -- no reference counterpart (the reference is single-device).
 * Rank `rank` owns `n_rows` feature rows of 1 or 2 matrices (x_real, x_imag).  Every 16-byte
 * piece is read ONCE from local HBM and stored to every peer's receive buffer (peer-mapped
 * addresses, e.g. torch symmetric memory) straight from registers, so sending costs one read
 * of the shard instead of world-1.  The rows travel in `n_slices` row slices (slice s = rows
 * [slice_row[s], slice_row[s+1])); when a slice is complete on all peers (per-CTA
 * fence.sys, device-scope arrival counter, last CTA), the word flag[p][s] on peer p is set to
 * `seq` with a system-scope release store.  The consumer orders its reads behind the flags with
 * pgsd_wait_flags on its own stream (a kernel boundary then separates the wait from the
 * non-coherent gathers of the aggregation).  `seq` must grow from call to call (wrap-around safe
 * comparison); `counters` is n_slices zeroed words of LOCAL scratch that the kernel leaves zeroed.
 * mc_dst != NULL: one multimem.st per piece to an NVSwitch multicast address instead of world-1
 * unicast stores (the switch replicates to every rank, including the sender).
 * The kernel is meant to run beside the aggregation on its own high-priority stream:
 * it occupies n_ctas resident-CTA slots for the whole exchange, which the aggregation launches
 * leave free through pgsd_spmm_args.grid_reserve.
 * ---------------------------------------------------------------------------------- */
#define PGSD_MAX_RANKS 16
#define PGSD_MAX_SLICES 16
typedef struct pgsd_push_args {
  int32_t world, rank;
  int32_t n_tensors;                 /* 1 or 2                                                   */
  int32_t row_bytes;                 /* bytes of one feature row, multiple of 16                 */
  int64_t n_rows;                    /* rows of the local shard                                  */
  const void* src[2];                /* local rows, 16-byte aligned                              */
  int64_t ld_src_bytes[2];
  void* dst[2][PGSD_MAX_RANKS];      /* dst[t][p]: where row 0 of THIS rank's shard lands on rank p */
  int64_t ld_dst_bytes[2];
  void* mc_dst[2];                   /* multicast alias of dst (same offset on every rank) or NULL */
  int32_t n_slices;
  int32_t n_ctas;                    /* grid size in 256-thread CTAs (0 = 64); engine 1: 32-thread CTAs */
  int64_t slice_row[PGSD_MAX_SLICES + 1];
  uint32_t* flag[PGSD_MAX_RANKS];    /* flag[p][s] on rank p for source = this rank              */
  uint32_t* counters;                /* [PGSD_MAX_SLICES + 1] local scratch, zero on entry and on exit */
  uint32_t seq;
  int32_t include_self;              /* 1: also store to dst[t][rank] and set flag[rank]          */
  int32_t engine;                    /* 0: LSU kernel (any row stride, multicast); 1: bulk-copy (TMA) kernel --
                                        cp.async.bulk global->shared->peer, contiguous rows only, else falls
                                        back to 0                                                   */
  int32_t chunk_bytes, stages;       /* engine 1: tile size (0 = 16384) and ring depth (0 = 4)    */
  int32_t reserved;
  uint32_t* started;                 /* LOCAL word set to seq once every CTA of the push is resident, or NULL:
                                        pgsd_wait_flags on it before launching the SM-filling aggregation, so
                                        the push always gets its SM slots (and shared memory) first          */
} pgsd_push_args;

PGSD_API size_t pgsd_sizeof_push_args(void);
PGSD_API int pgsd_shard_push(const pgsd_push_args* args, pgsd_stream_t stream);

/* Content fingerprint of a device buffer: out2[0] += sum of its 32-bit words, out2[1] += position-weighted sum
 * (both wrap around; the caller zero-initialises out2).  The Python plan cache compares it before it reuses a plan
 * for edge tensors whose identity (pointer, shape, version counter) has not changed -- the reference itself
 * re-normalises on every call (nn/general/conv_base.py:102-108, MagNetConv.py:158-181 with cached=False). */
PGSD_API int pgsd_fingerprint(const void* data, int64_t n_bytes, uint64_t* out2, pgsd_stream_t stream);

/* Copy-engine variant of the same exchange (no SM involved): a stream-ordered device-to-device / peer copy and a
 * one-word signal (fence.sys + st.release.sys of `seq`) that is enqueued behind the copies of a slice. */
PGSD_API int pgsd_peer_copy(void* dst, const void* src, size_t bytes, pgsd_stream_t stream);
PGSD_API int pgsd_signal_flag(uint32_t* flag, uint32_t seq, pgsd_stream_t stream);

/* Stream-ordered wait: returns (on the stream) once flags[index_host[i]] has reached `seq` for all
 * i < n (n <= 64), polling with ld.acquire.sys.  After timeout_ns the kernel gives up, writes 1 to
 * *status (device int32, may be NULL) and lets the stream continue -- the caller checks status at
 * the end of the step instead of hanging the GPU. */
PGSD_API int pgsd_wait_flags(const uint32_t* flags, const int32_t* index_host, int32_t n, uint32_t seq,
                             uint64_t timeout_ns, int32_t* status, pgsd_stream_t stream);

/* Signed-triangle motif counts (SDGNN / SiGAT preprocessing; replaces the Python set loops of
 * nn/signed/SDGNN.py:153-254 get_features()/build_adj_lists() and nn/signed/SiGAT.py:93-186).
 * row_ptr4 / col4: HOST arrays of 4 DEVICE pointers -- the CSR neighbour lists pos_out, pos_in, neg_out,
 * neg_in of every node, each row sorted and duplicate-free (the reference keeps Python sets).
 * counts[e*16 + t] = |N_lu(t)(edge_u[e]) intersect N_lv(t)(edge_v[e])| in the order of the reference's tuple
 * (d1_1..d1_4, d2_1..d2_4, d3_1..d3_4, d4_1..d4_4).  Exact integers. */
PGSD_API int pgsd_signed_triangle_counts(const int32_t* const* row_ptr4, const int32_t* const* col4,
                                  const int64_t* edge_u, const int64_t* edge_v, int64_t n_edges,
                                  int64_t n_nodes, int32_t* counts, pgsd_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PGSD_B200_H */
