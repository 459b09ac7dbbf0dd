/*
 * gala_b200.h -- C-ABI of the B200-native (sm_100a) sparse aggregation library that
 * replaces the kernels GALA's code generator emits as text into gala.cu.
 *
 * The reference has no FFI: its "operator interface" for this path is the set of
 * free functions that CUDAGenerator::generateCudaCodeForCNode writes into every
 * generated program (src/codegen/cuda.h:170-955 of ADAPT-uiuc/GALA, cited below as
 * cuda.h:LINE) and that the emitted autograd classes call
 * (src/codegen/common.h:622-977, cited as common.h:LINE).  Each entry point here
 * names the emitted function / kernel it replaces.  The libtorch shim that
 * re-creates the emitted names on top of this ABI is
 * gala-gnn-acceleration-language_b200/host/gala_b200_torch.h; INTEGRATION.md shows the
 * codegen edit.
 *
 * Conventions
 *   - plain pointers and sizes only; every data pointer is DEVICE memory unless
 *     the parameter is documented as host;
 *   - int32 indices, float32 values, row-major dense matrices
 *     (common.h:1682-1693);
 *   - every function returns int: 0 = success, > 0 = a cudaError_t,
 *     < 0 = a GALA_ERR_* argument error.  Nothing exits or throws across the ABI
 *     (the reference printf+exit()s, cuda.h:980-998);
 *   - work is enqueued on the given stream (a cudaStream_t passed as void*;
 *     NULL = the legacy default stream); no hidden synchronisation, no allocation
 *     of caller-visible memory, no global state;
 *   - outputs and workspaces are caller-provided.
 *
 * Graph layout ("column-tiled CSR", src/ops/tiling.h:222-283, Appendix B of
 * SURVEY.md): `segments` consecutive LOCAL row-pointer arrays of length nrows+1
 * in `offsets`; `cols` (and every per-edge value array) is segment-major,
 * row-major inside a segment; segment s owns edges
 * [bounds[2s], bounds[2s+1]) and row i of segment s owns
 * bounds[2s] + offsets[s*(nrows+1)+i .. i+1).  A plain CSR is segments = 1.
 * `bounds` is a HOST array exactly as in the generated code (cuda.h:472-475 reads
 * it on the host); it may be NULL when segments == 1.  Graphs with more than 64 segments
 * also pass a device copy of it (bounds_dev).
 */
#ifndef GALA_B200_H
#define GALA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define GALA_B200_ABI_VERSION 4   /* v2: row pitches ldx / ldy, gala_pad_rows_f32; v3: edge tiles in gala_plan_t;
                                     v4: gala_gat_forward_col_f32, gala_reflection_f32 (additions only) */

#define GALA_OK 0
#define GALA_ERR_NULL_POINTER (-1)
#define GALA_ERR_BAD_SHAPE (-2)
#define GALA_ERR_UNSUPPORTED (-3)
#define GALA_ERR_WORKSPACE (-4)
#define GALA_ERR_MISALIGNED (-5)

typedef void *gala_stream_t; /* cudaStream_t */

typedef struct gala_graph {
    const int32_t *offsets; /* device, [segments * (nrows + 1)]                      */
    const int32_t *cols;    /* device, [nvals]                                       */
    const int32_t *bounds;  /* HOST,   [2 * segments]; NULL allowed iff segments == 1 */
    int32_t nrows;          /* rows of the sparse matrix == rows of every output     */
    int32_t ncols;          /* columns == rows of the gathered dense operand         */
    int32_t segments;       /* S >= 1                                                */
    int64_t nvals;          /* E                                                     */
    const int32_t *bounds_dev; /* device copy of bounds [2 * segments]; required iff segments > 64
                                  (up to 64 segment starts travel in the kernel parameters), else may be NULL */
} gala_graph_t;

/*
 * Load-balance plan for power-law graphs: the list of "hub" rows (total degree
 * above `hub_threshold`) that are executed by a whole thread block instead of a
 * single warp, and the remaining rows in degree-descending order (so the warps of
 * a block carry rows of equal length and long rows are scheduled first).  Built
 * once per graph on the device, reused by every call.  All kernels accept
 * plan == NULL (pure warp-per-row in natural row order).
 *
 * Edge tiles (single-segment graphs): the edge arrays cut into spans of `tile_edges`
 * consecutive edges, tile t owning the rows that start inside it (an nnz split of the
 * rows, cf. nnz_ord_row_tile_info, src/ops/tiling.h:1656-1708).  With them the streaming
 * edge kernels (row sum, row scaling, edge-softmax forward / backward) run
 * edge-parallel: persistent thread blocks stage tile after tile in shared memory with
 * bulk copies, whatever the degree distribution.  tile_rows == NULL: row-structured
 * kernels.
 */
typedef struct gala_plan {
    const int32_t *hub_rows;  /* device, [n_hub]: rows run by a whole CTA              */
    const int32_t *row_order; /* device, [n_ordered]: all other rows, longest first    */
    int32_t n_hub;
    int32_t n_ordered;        /* == nrows - n_hub                                      */
    int32_t hub_threshold;
    const int32_t *tile_rows; /* device, [2 * (n_tiles + 1)]: (first row, first edge) of every edge
                                 tile, closed by (nrows, nvals); 8-byte aligned; or NULL        */
    int32_t n_tiles;          /* ceil(nvals / tile_edges)                              */
    int32_t tile_edges;       /* edges per tile the table was built for                */
    int32_t tile_policy;      /* GALA_TILES_AUTO: edge-parallel kernels where they measured
                                 faster (row scaling; every op when the mean degree is below
                                 96); GALA_TILES_ALWAYS: whenever the table is usable      */
} gala_plan_t;
#define GALA_TILES_AUTO 0
#define GALA_TILES_ALWAYS 1

/* Fused epilogue / prologue of the SpMM (all fields optional):                   */
/*   Y[i,:] = act( row_scale[i] * sum_e w_e * col_scale[col_e] * X[col_e,:]       */
/*                 + (accumulate ? Y[i,:] : 0) )                                   */
/* row_scale/col_scale fold the `norm * res` passes the generated GCN/SAGE models */
/* run as separate ATen kernels (common.h:1128-1180, codegen/gala.cu:441-457).    */
typedef struct gala_epilogue {
    const float *row_scale; /* device [nrows] or NULL */
    const float *col_scale; /* device [ncols] or NULL */
    int32_t accumulate;     /* 0: Y is overwritten; 1: Y += (reference semantics
                               of the per-segment `C = C + ...`, cuda.h:309-351)  */
    int32_t relu;           /* apply max(.,0) last                                 */
    int32_t schedule;       /* column-tiled graphs: GALA_SCHEDULE_AUTO picks
                               segment-major (one launch per column segment, so the
                               slice of X a segment gathers from stays in L2) when
                               ncols*K*4 exceeds the L2, row-major (one launch)
                               otherwise; or force either                            */
    int64_t ldx;            /* row pitch of X in elements; 0 = K (packed rows).  A pitch that
                               is a multiple of 4 on a 16-byte aligned base lets every row be
                               gathered with 128-bit loads whatever K is (K = 41 stored with
                               pitch 44: the reference's K % 32 remainder kernels,
                               cuda.h:58-168, become one pass); gala_pad_rows_f32 re-pitches  */
    int64_t ldy;            /* row pitch of Y in elements; 0 = K                              */
    const struct gala_multi_out *multi_out; /* nullable: the finished rows go to every GPU (packed, pitch K)
                               instead of Y -- see gala_multi_out_t below                     */
} gala_epilogue_t;
#define GALA_SCHEDULE_AUTO 0
#define GALA_SCHEDULE_ROW_MAJOR 1
#define GALA_SCHEDULE_SEGMENT_MAJOR 2

int gala_b200_abi_version(void);
const char *gala_b200_error_string(int code);

/* Measurement utility (no reference counterpart): every SM streams `bytes` of `buf`   */
/* `repeats` times with 128-bit ld.global.cg loads.  A buffer that fits the L2 gives   */
/* the L2->SM read bandwidth the gather kernels are bounded by, a large one the HBM    */
/* read bandwidth.  *sink (device, 16 bytes) receives a value that depends on the data. */
int gala_b200_probe_read(const void *buf, size_t bytes, int32_t repeats, void *sink,
                         gala_stream_t stream);

/* ---- plan ------------------------------------------------------------------ */
/* Bytes of device workspace gala_plan_build needs for this graph.               */
size_t gala_plan_workspace_bytes(const gala_graph_t *g);
/* Scans the row pointers on the device, writes the hub list into `workspace`    */
/* and fills *plan (host struct).  Synchronises `stream` once to read the count. */
int gala_plan_build(const gala_graph_t *g, int32_t hub_threshold, void *workspace,
                    size_t workspace_bytes, gala_plan_t *plan, gala_stream_t stream);

/* ---- node aggregation ------------------------------------------------------- */
/* Replaces aggregate_node_mul_sum[_direct][_coarseN]_call + kernels K1/K2        */
/* (cuda.h:282-503, launch tree :58-168; sample codegen/gala.cu:65-390) and the   */
/* cuSPARSE flavour (cuda.h:211-280).  Y[nrows,K] = A * X[ncols,K]; `vals` is the  */
/* per-edge weight array (value_graph) or NULL for an unweighted graph            */
/* (cuda.h:292-295).  One launch covers every column segment and any K.          */
int gala_spmm_f32(const gala_graph_t *g, const float *vals, const float *X, int32_t K, float *Y,
                  const gala_epilogue_t *ep, const gala_plan_t *plan, gala_stream_t stream);

/* Replaces the sampled flavour K1s (cuda.h:313-320, 389-396): per row and per    */
/* segment, sum over j = (ra*ji + rb) % deg for ji in [0, nsamples).              */
/* (ra, rb) = global_ra/global_rb (common.h:817-830).  Y is overwritten unless    */
/* accumulate != 0.                                                              */
/* ldx / ldy: row pitches of X / Y in elements, 0 = K.                              */
int gala_spmm_sampled_f32(const gala_graph_t *g, const float *vals, const float *X, int32_t K,
                          float *Y, int32_t nsamples, int32_t ra, int32_t rb, int32_t accumulate,
                          int64_t ldx, int64_t ldy, gala_stream_t stream);

/* Re-pitch a dense matrix: Xp[r, 0:K] = X[r, 0:K], Xp[r, K:ld_out] = 0 (X with row pitch   */
/* ld_in >= K, Xp with ld_out >= K).  What a producer that cannot write pitched rows      */
/* itself runs once (N*K*8 bytes) before aggregating at a width that is not a multiple of */
/* 4 -- the class counts 41 / 47 of the shipped schedules.                                 */
int gala_pad_rows_f32(const float *X, int64_t nrows, int32_t K, int64_t ld_in, float *Xp,
                      int64_t ld_out, gala_stream_t stream);

/* ---- optional reduced-precision feature storage (not in the reference: its path is fp32 throughout) ---- */
/* Same kernels with the GATHERED rows stored as bf16 (X: [ncols, K] bf16, row pitch 2K bytes): half the      */
/* bytes per edge on the L2/HBM-bound gather.  Edge values, attention logits, accumulation and Y stay fp32;   */
/* the result equals the fp32 entry point run on bf16-rounded X (tests), i.e. within ~4e-3 of the fp32 result */
/* (north-star bound for bf16 features: 1e-2).  K in {8,16,32,64,128,256}; X and Y 16-byte aligned.           */
int gala_spmm_bf16(const gala_graph_t *g, const float *vals, const uint16_t *X_bf16, int32_t K, float *Y,
                   const gala_epilogue_t *epilogue, const gala_plan_t *plan, gala_stream_t stream);
int gala_gat_forward_bf16(const gala_graph_t *g, const float *aL, const float *aR, const uint16_t *X_bf16,
                          int32_t K, float slope, float *Y, float *alpha_out, int32_t relu,
                          const gala_plan_t *plan, gala_stream_t stream);

/* ---- edge kernels ------------------------------------------------------------ */
/* Replaces node_spmv_backward_of_sddmm_{nln,eaggr} + K3 (cuda.h:505-524,         */
/* 565-600, 659-678, 737-772): out[i] = sum_s (seed + sum_{e in row i, seg s}     */
/* vals[e]).  The reference seeds with 1e-12f once per segment.                   */
int gala_edge_rowsum_f32(const gala_graph_t *g, const float *vals, float *out, float seed,
                         const gala_plan_t *plan, gala_stream_t stream);

/* Replaces inplace_softmax_sddvv[_mult] + K4 (cuda.h:525-562, 601-656):          */
/* vals[e] *= rowval[row(e)] in place.                                            */
int gala_edge_scale_rows_f32(const gala_graph_t *g, float *vals, const float *rowval,
                             const gala_plan_t *plan, gala_stream_t stream);

/* Replaces edge_sddvv + K5 (cuda.h:679-698, 773-807) when op == GALA_SDDVV_ADD   */
/* and aggregate_edge_mul[_dir] + K7 (cuda.h:848-952) when op == GALA_SDDVV_MUL:  */
/* out[e] = A[row(e)] (+|*) B[col(e)].  If leaky_slope != 1, LeakyReLU(slope) is   */
/* applied to the result (the ATen pass at common.h:1180 fused in).               */
#define GALA_SDDVV_ADD 0
#define GALA_SDDVV_MUL 1
int gala_sddvv_f32(const gala_graph_t *g, const float *A, const float *B, float *out, int32_t op,
                   float leaky_slope, const gala_plan_t *plan, gala_stream_t stream);

/* Replaces edge_sddmm + K6 (cuda.h:699-734, 808-845): out[e] = sum_k A[row,k] *   */
/* B[col,k].  Implements the mathematical definition; the reference kernel's      */
/* shared-memory aliasing between the 8 rows of a block (cuda.h:706-714) is not   */
/* reproduced.                                                                    */
int gala_sddmm_f32(const gala_graph_t *g, const float *A, const float *B, int32_t K, float *out,
                   const gala_plan_t *plan, gala_stream_t stream);

/* The edge side of one GAT layer's backward as autograd runs it in the generated    */
/* program, in one kernel and without E-sized temporaries: softmax backward            */
/* (common.h:791-799), LeakyReLU backward, and the row sum the edge-sum backward         */
/* returns for BOTH attention inputs (common.h:630-675, node_spmv_backward_of_sddmm_eaggr): */
/*   ds = alpha*dalpha - alpha*(S*1e-12 + sum_row alpha*dalpha)                          */
/*   d_att[i] = S*1e-12 + sum_{e in row i} ds[e] * (aL[i] + aR[col[e]] > 0 ? 1 : slope)  */
int gala_gat_backward_att_f32(const gala_graph_t *g, const float *alpha, const float *dalpha,
                              const float *aL, const float *aR, float slope, float *d_att,
                              const gala_plan_t *plan, gala_stream_t stream);

/* Replaces the 5-pass forward of non_lnr_op_softmax_AutoGrad (common.h:760-773): */
/* alpha[e] = clamp(exp(x[e]),0,1e12) / (S*1e-12 + sum_row clamp(exp(x))).        */
/* x and alpha may alias.  recip (nullable, [nrows]) receives 1/rowsum.           */
int gala_edge_softmax_fwd_f32(const gala_graph_t *g, const float *x, float *alpha, float *recip,
                              const gala_plan_t *plan, gala_stream_t stream);

/* Replaces the backward of the same class (common.h:791-799):                    */
/* out[e] = alpha*dalpha - alpha * (S*1e-12 + sum_row alpha*dalpha).               */
int gala_edge_softmax_bwd_f32(const gala_graph_t *g, const float *alpha, const float *dalpha,
                              float *out, const gala_plan_t *plan, gala_stream_t stream);

/* ---- fused GAT layer (what the retargeted code generator emits) ---------------- */
/* One pass over the edges: e = aL[row]+aR[col] -> LeakyReLU(slope) -> edge-       */
/* softmax -> Y = alpha * X (-> ReLU if relu).  Equivalent to the emitted          */
/* sequence edge_sddvv, LeakyReLU, non_lnr_op_softmax, aggregate_node_mul_sum      */
/* (common.h:622-675, 735-810, 835-927).  alpha_out (nullable, [nvals]) receives   */
/* the attention values for the backward pass.                                     */
int gala_gat_forward_f32(const gala_graph_t *g, const float *aL, const float *aR, const float *X,
                         int32_t K, float slope, float *Y, float *alpha_out, int32_t relu,
                         const gala_plan_t *plan, gala_stream_t stream);

/* The same kernel with a fused dense epilogue on each finished output row y (K <= 128):    */
/*   att_w  [2,K] device, att_b[2]: att_out[row] = y.att_w[0] + att_b[0] and                 */
/*          att_out[nrows+row] = y.att_w[1] + att_b[1] -- the NEXT layer's attenL/attenR      */
/*          projections (Linear(h,1), frontend.y:987-994), att_out device [2*nrows];          */
/*   cls_wT [K,cls_n] device (the Linear weight transposed), cls_b [cls_n] nullable:          */
/*          cls_out[row,:] = y @ cls_wT + cls_b -- the classifier / FFN that follows the      */
/*          aggregation (FFN-recompute rewrite, middle-end.h:324-375).  Y may be NULL then.   */
/* Output rows pushed to every GPU of the node while they are produced (compute fused with  */
/* the all-gather that would follow): base[q] = address, in GPU q's peer-mapped copy of the  */
/* gathered buffer, where row 0 of THIS rank's slab lives; multicast_base = the same address */
/* in the NVLS multicast mapping (one multimem.st, replicated by the NVSwitch) or NULL.      */
/* The caller synchronises the GPUs (a barrier on the stream) before anyone reads the rows.  */
typedef struct gala_multi_out {
    float *base[8];
    float *multicast_base;
    int32_t count;
    const uint8_t *need_mask; /* device [rows of this rank] or NULL: bit q set = GPU q references the row and gets it.
                                 The "exchange only the rows a peer needs" of a 1-D partition as
                                 predicated peer stores; ignored on the multicast path (the switch replicates to all). */
} gala_multi_out_t;

typedef struct gala_dense_epilogue {
    const float *att_w;
    float att_b[2];
    float *att_out;
    const float *cls_wT;
    const float *cls_b;
    float *cls_out;
    int32_t cls_n;
    const gala_multi_out_t *multi_out; /* nullable: Y rows go to every GPU instead of Y */
    int64_t ldx, ldy;                  /* row pitches of X / Y in elements, 0 = K       */
    const gala_multi_out_t *att_multi_out; /* nullable: att_out[nrows+row] (the next layer's attenR, the scalar every
                                          GPU gathers by column) is also stored at element `row` of the given bases */
} gala_dense_epilogue_t;
int gala_gat_forward_ex_f32(const gala_graph_t *g, const float *aL, const float *aR, const float *X,
                            int32_t K, float slope, float *Y, float *alpha_out, int32_t relu,
                            const gala_dense_epilogue_t *ep, const gala_plan_t *plan,
                            gala_stream_t stream);

/* Same layer when the right-hand attention term is a linear function of the aggregated    */
/* features, aR[j] = dot(X[j,:], wR) + bR -- which is how the generated GAT computes it      */
/* (attenR = efc(res), common.h:1185-1281; with the layer-2 FFN-recompute rewrite wR = W1^T  */
/* w and bR = w.b1 + b, middle-end.h:324-375).  The kernel recomputes aR from the row it has */
/* just gathered, so the second random gather per edge disappears.  K % 4 == 0, K <= 32,     */
/* X and Y 16-byte aligned; otherwise GALA_ERR_UNSUPPORTED (use gala_gat_forward_f32).        */
int gala_gat_forward_dot_f32(const gala_graph_t *g, const float *aL, const float *wR, float bR,
                             const float *X, int32_t K, float slope, float *Y, float *alpha_out,
                             int32_t relu, const gala_plan_t *plan, gala_stream_t stream);

/* Same layer for features stored in a REFLECTED basis whose last column carries the right-hand */
/* attention term (ABI v4).  attenR = efc(res) is a linear function w.res + b of the rows the     */
/* layer aggregates (common.h:1185-1281), and the aggregation is linear in those rows, so the     */
/* producer may hand over X' = X H with H = I - 2 v v^T the Householder reflection that maps the  */
/* last unit vector onto -+ w/|w| (gala_reflection_f32 below builds v and sR = -+|w|): then       */
/*     aR[j] = sR * X'[j, K-1] + bR,     sum_j alpha_j X[j] = (sum_j alpha_j X'[j]) H,            */
/* H is orthogonal and its own inverse, and it folds into the weights of the transform that       */
/* produces X (W' = H W, b' = H b) at no cost.  The kernel reads the scalar from the 128-byte row  */
/* it gathers anyway -- 4 sectors per edge instead of 4 + 1, no second random gather, no dot       */
/* product -- and applies, per finished output row y (normalised sum):                             */
/*     reflect_in  (device [K], nullable): y <- y - 2 v (v.y), back to the original basis;         */
/*     ReLU if relu;                                                                               */
/*     reflect_out (device [K], nullable): the same with the NEXT layer's vector.                   */
/* ep (nullable) as for gala_gat_forward_ex_f32: row pitches, the dense epilogue on the FINAL row   */
/* (after reflect_out: att_w then holds the next layer's projections expressed in that basis; only   */
/* att_out[0:nrows] is needed by a next layer of this kind, and with att_w alone -- no cls_wT -- only  */
/* that half is written, straight from the registers holding the row), multi_out; att_multi_out unset */
/* (there is no scalar left to exchange).  K in {4, 8, 16, 32}; X, Y 16-byte aligned, row pitches     */
/* multiples of 4; otherwise GALA_ERR_UNSUPPORTED (use gala_gat_forward_f32).  alpha_out as above.   */
int gala_gat_forward_col_f32(const gala_graph_t *g, const float *aL, float sR, float bR, const float *X,
                             int32_t K, float slope, float *Y, float *alpha_out, int32_t relu,
                             const float *reflect_in, const float *reflect_out,
                             const gala_dense_epilogue_t *ep, const gala_plan_t *plan, gala_stream_t stream);

/* Host helper (no device work): the unit vector v[K] of the reflection H = I - 2 v v^T with      */
/* H e_{K-1} = -sign(w[K-1]) w/|w| and *sR = -sign(w[K-1]) |w|, so that (X H)[:, K-1] * sR = X w.  */
/* Accumulates in double.  w == 0 -> GALA_ERR_BAD_SHAPE (attention does not depend on the column). */
int gala_reflection_f32(const float *w, int32_t K, float *v, float *sR);

/* ---- dense feature transform on the tensor cores -------- */
/* Replaces torch::nn::Linear of the generated model (common.h:1185-1281; cuBLAS fp32 SIMT   */
/* through libtorch): Y[M,N] = X[M,K] * W[N,K]^T + bias, optional ReLU.  tcgen05.mma         */
/* kind::tf32 with the accumulator in TMEM, error-compensated (3xTF32) so that the result     */
/* stays within the fp32 parity bound.  N <= 256; the fused row epilogues (attention          */
/* projections, multi_out) need N <= 64 (hidden / class widths of the GNN layers).            */
/* Optional fused attention projections of a GAT layer (attenL/attenR = Linear(h,1)(res),      */
/* frontend.y:987-994): att_out[0:M] = Y_pre_relu . att_w[0,:] + att_b[0], att_out[M:2M] the   */
/* same with row 1.  att_w device [2,N], att_out device [2,M]; att_b [2] on the HOST, or on the */
/* device when att_b_on_device != 0 (trained biases: no host read-back per step); all nullable. */
/* row_scale (device [M], nullable) multiplies output row r before the ReLU: the `norm * res`   */
/* pass that follows the transform in the generated GCN (codegen/gala.cu:441-443).              */
/* multi_out (nullable): push the output rows to every GPU instead of Y (N % 4 == 0).          */
/* att_multi_out (nullable): att_out[M + r] (attenR of row r) is also stored at element r of     */
/* every GPU's gathered vector -- the one scalar per node a partitioned GAT layer exchanges.     */
int gala_linear_f32(const float *X, int64_t M, int32_t K, const float *W, const float *bias,
                    int32_t N, float *Y, const float *row_scale, int32_t relu, const float *att_w,
                    const float *att_b, int32_t att_b_on_device, float *att_out,
                    const struct gala_multi_out *multi_out, const struct gala_multi_out *att_multi_out,
                    gala_stream_t stream);

/* The narrow transforms that follow an aggregation (classifier Linear(h, classes), the two    */
/* Linear(h,1) attention projections; common.h:1185-1281): K <= 64, N <= 64, exact fp32 FMA,   */
/* weights in registers, one streaming pass.  transpose_out != 0 writes Y as [N, M] (each      */
/* projection a contiguous vector).                                                            */
int gala_linear_small_f32(const float *X, int64_t M, int32_t K, const float *W, const float *bias,
                          int32_t N, float *Y, int32_t relu, int32_t transpose_out,
                          gala_stream_t stream);

/* The same narrow transform with the row epilogues of gala_linear_f32 (row_scale before the ReLU, the  */
/* two attention projections of the pre-activation row into att_out [2, M] with att_b[2] on the HOST,    */
/* output rows / right-hand attention scalars pushed to every GPU through multi_out / att_multi_out).    */
/* A deliberately light kernel (256-thread blocks, at most max_ctas of them; 0 = default) that can run   */
/* on a side stream next to an aggregation kernel: the row-block pipeline of the partitioned layers      */
/* hides the push of layer l+1's transformed rows behind layer l's aggregation.   */
int gala_linear_small_ex_f32(const float *X, int64_t M, int32_t K, const float *W, const float *bias,
                             int32_t N, float *Y, const float *row_scale, int32_t relu,
                             const float *att_w, const float *att_b, float *att_out,
                             const struct gala_multi_out *multi_out,
                             const struct gala_multi_out *att_multi_out, int32_t max_ctas,
                             gala_stream_t stream);

/* The exchange step of a partitioned layer as its own light kernel: rows of the local matrix X [M, K]   */
/* (row pitch ldx, 0 = K; K % 4 == 0) and, optionally, one scalar per row are stored into every GPU that */
/* gathers them (multi_out / scalar_multi_out as above), whole 128-byte lines per store instruction.      */
/* For producers that finish row block b while block b+1 is still being computed (the copy runs on a      */
/* side stream), with at most max_ctas 256-thread blocks (0 = 148).                                        */
int gala_push_rows_f32(const float *X, int64_t M, int32_t K, int64_t ldx, const float *scalars,
                       const struct gala_multi_out *multi_out,
                       const struct gala_multi_out *scalar_multi_out, int32_t max_ctas,
                       gala_stream_t stream);

/* ---- format construction on the device ---------- */
/* All integer outputs are bit-exact against the reference functions named below.       */

/* Replaces CSRCMatrix::build, CSR branch (src/formats/csrc_matrix.h:148-282; count /    */
/* prefix / place / per-row sort, src/utils/mtx_sort.h:52-64,165-174,114-137,683-722):   */
/* COO (row_ids, col_ids[, vals]) in any order, duplicates kept -> offsets[nrows+1],     */
/* ids[nvals] sorted by (row, col), out_vals travelling with their edge (stable among    */
/* duplicates; the reference's order among duplicates is unspecified).  vals/out_vals    */
/* may be NULL (the pipeline sets every value to 1 afterwards, tests/common.h:363).      */
size_t gala_csr_from_coo_workspace_bytes(int32_t nrows, int32_t ncols, int64_t nvals);
int gala_csr_from_coo(int32_t nrows, int32_t ncols, int64_t nvals, const int32_t *row_ids,
                      const int32_t *col_ids, const float *vals, int32_t *offsets, int32_t *ids,
                      float *out_vals, void *workspace, size_t workspace_bytes,
                      gala_stream_t stream);

/* Replaces buildTranspose (tests/common.h:107-123) + get_sids (csrc_matrix.h:399-411). */
/* Workspace: gala_csr_from_coo_workspace_bytes(ncols, nrows, nvals).                   */
int gala_csr_transpose(int32_t nrows, int32_t ncols, int64_t nvals, const int32_t *offsets,
                       const int32_t *ids, const float *vals, int32_t *t_offsets, int32_t *t_ids,
                       float *t_vals, void *workspace, size_t workspace_bytes,
                       gala_stream_t stream);

/* Replaces rowReorderToAdj (src/ops/reordering.h:940-1013): perm[i] = new index of   */
/* node i; row perm[i] of the result holds (perm[col], val) of row i, sorted by (column, */
/* value).  Square graphs (the permutation relabels rows and columns).  perm must be a   */
/* permutation of [0, nrows).  Workspace: gala_csr_from_coo_workspace_bytes(n, n, nvals). */
int gala_csr_reorder(int32_t nrows, int64_t nvals, const int32_t *offsets, const int32_t *ids,
                     const float *vals, const int32_t *perm, int32_t *new_offsets, int32_t *new_ids,
                     float *new_vals, void *workspace, size_t workspace_bytes, gala_stream_t stream);

/* Replaces rowPermuteDenseTo (reordering.h:244-283; from = 0: Y[perm[i],:] = X[i,:])    */
/* and rowPermuteDenseFrom (:207-236; from = 1: Y[i,:] = X[perm[i],:]).  Out of place.   */
int gala_permute_rows_f32(const float *X, const int32_t *perm, float *Y, int32_t nrows, int32_t K,
                          int32_t from, gala_stream_t stream);

/* A permutation generator for the above (the reference ships only the identity and the  */
/* reversal, getAcendingOrder / getDecendingOrder reordering.h:1085-1103; its rabbit     */
/* order is commented out): nodes by DESCENDING degree, ties by node id.                 */
/* perm[i] = new index of node i ("to" form); order[k] (nullable) = node at new index k. */
size_t gala_degree_order_workspace_bytes(int32_t nrows);
int gala_degree_order(int32_t nrows, const int32_t *offsets, int32_t *perm, int32_t *order,
                      void *workspace, size_t workspace_bytes, gala_stream_t stream);

/* Replaces static_ord_col_breakpoints (src/ops/tiling.h:1594-1608) +                   */
/* ord_col_tiling_torch (:222-283).  segments = ceil(ncols / cols_per_partition);       */
/* out_offsets[segments*(nrows+1)], out_cols/out_vals[nvals] on the device,             */
/* bounds_host[2*segments] on the HOST (the call synchronises the stream once).         */
/* Requires column-sorted rows, as the reference does (tiling.h:273-276).               */
int32_t gala_col_tile_segments(int32_t ncols, int32_t cols_per_partition);
size_t gala_col_tile_workspace_bytes(int32_t nrows, int32_t ncols, int32_t cols_per_partition);
int gala_col_tile(int32_t nrows, int32_t ncols, int64_t nvals, const int32_t *offsets,
                  const int32_t *ids, const float *vals, int32_t cols_per_partition,
                  int32_t *out_offsets, int32_t *out_cols, float *out_vals, int32_t *bounds_host,
                  void *workspace, size_t workspace_bytes, gala_stream_t stream);

/* Replaces inplace_sample_graph_ab (src/ops/tiling.h:454-508): per row keep            */
/* sample_size edges at the sorted positions (ra*ji+rb) % deg; new_offsets[i] =         */
/* i*sample_size.  sample_size <= 128.  *status_dev (device int) is set to 1 if a row   */
/* has no edge (`% 0` in the reference).                                                */
int gala_sample_ab(int32_t nrows, const int32_t *offsets, const int32_t *ids, const float *vals,
                   int32_t sample_size, int32_t ra, int32_t rb, int32_t *new_offsets,
                   int32_t *new_ids, float *new_vals, int32_t *status_dev, gala_stream_t stream);

/* One layer of getMaskSubgraphs (tests/common.h:20-105): forward sub-graph = rows with  */
/* mask > 0 kept whole (new_ids/new_vals sized for nvals), *new_nvals_host = its edge    */
/* count (the call synchronises once); next_mask (nullable) = one-hop growth of the      */
/* mask (maxAgg gSpMM into a ZEROED buffer; the reference leaves it uninitialised).      */
/* The backward graph is gala_csr_transpose of the result.                               */
size_t gala_mask_subgraph_workspace_bytes(int32_t nrows);
int gala_mask_subgraph(int32_t nrows, const int32_t *offsets, const int32_t *ids, const float *vals,
                       const uint8_t *mask, int32_t *new_offsets, int32_t *new_ids, float *new_vals,
                       int64_t *new_nvals_host, uint8_t *next_mask, void *workspace,
                       size_t workspace_bytes, gala_stream_t stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* GALA_B200_H */
