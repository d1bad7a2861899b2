/*
 * glass_b200.h -- C ABI of libglass_b200.so: hand-written sm_100a kernels for the GLASS
 * labeled message-passing hot path (reference: Xi-yuanWang/GLASS, impl/models.py, impl/utils.py).
 *
 * The reference is pure Python on PyTorch/PyG and has no FFI of its own; every entry point below
 * replaces a group of ATen / PyG library calls at the cited reference lines.  The Python host
 * layer (glass_b200/ops.py) binds these with ctypes and registers them as torch custom ops;
 * INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * Conventions
 *   - All pointers are DEVICE pointers unless the name ends in _host.  No allocation, no free,
 *     no host synchronisation inside any entry point except glass_csr_build (init path, which
 *     returns the de-duplicated entry count to the host).  Everything else is CUDA-graph capturable.
 *   - Dense matrices are row-major fp32 with an explicit leading dimension `ld*` in elements.
 *   - Index arrays handed in by the reference API are int64 (torch default); the CSR produced and
 *     consumed here uses int32 (n_node, nnz < 2^31).
 *   - `stream` is a cudaStream_t passed as void*; NULL is the legacy default stream.
 *   - Return value: 0 on success, a negative glass_status otherwise; glass_last_error() gives
 *     the message of the last failure on the calling thread.  No entry point ever falls back
 *     to a CPU path.
 */
#ifndef GLASS_B200_H_
#define GLASS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GLASS_B200_ABI_VERSION 2

typedef enum {
    GLASS_OK = 0,
    GLASS_ERR_BAD_ARG = -1,      /* shape / enum / alignment not supported */
    GLASS_ERR_CUDA = -2,         /* a CUDA runtime call or launch failed */
    GLASS_ERR_WORKSPACE = -3,    /* workspace too small */
    GLASS_ERR_UNSUPPORTED = -4   /* configuration outside the kernel's envelope */
} glass_status;

typedef enum { GLASS_AGGR_MEAN = 0, GLASS_AGGR_SUM = 1, GLASS_AGGR_GCN = 2 } glass_aggr;       /* impl/models.py:95-109 */
typedef enum { GLASS_ACT_NONE = 0, GLASS_ACT_RELU = 1, GLASS_ACT_ELU = 2 } glass_act;          /* GLASSTest.py:143 passes ELU */
typedef enum { GLASS_POOL_SUM = 0, GLASS_POOL_MEAN = 1, GLASS_POOL_MAX = 2, GLASS_POOL_SIZE = 3 } glass_pool; /* impl/models.py:295-319 */
typedef enum { GLASS_GEMM_AUTO = 0, GLASS_GEMM_SIMT = 1, GLASS_GEMM_TCGEN05 = 2 } glass_gemm_path;

int glass_abi_version(void);
const char* glass_last_error(void);
/* Number of SMs of the current device (grid sizing); negative glass_status on failure. */
int glass_sm_count(void);

/* ------------------------------------------------------------------------------------------
 * buildAdj  (impl/models.py:83-111) -> CSR + transposed CSR
 *
 * Input: COO edge_index int64 [2, nnz] (row = edge_index[0..nnz), col = edge_index[nnz..2nnz)),
 * edge_weight fp32 [nnz].  Any order, duplicates and self loops allowed.  Semantics
 * (bit-exact vs. buildAdj(...).coalesce() for unit weights, see oracle/glass_oracle.py
 * build_csr_numpy): deg = row sums of the raw entries; deg < 0.5 -> += 1; mean: (1/deg)[r]*w;
 * sum: w; gcn: ((deg^-1/2)[r]*w)*(deg^-1/2)[c] with IEEE 1/sqrt; duplicates normalised first, then
 * summed in input order; entries sorted by (row, col).
 * Outputs (capacity nnz each): rowptr/rowptr_t int32 [n_node+1], col/col_t int32, val/val_t fp32,
 * deg fp32 [n_node] (after the +1 fix).  *nnz_out_host receives the number of distinct entries.
 * The transposed triple is the CSR of A^T (needed by backward: `mean` is not symmetric).
 * ------------------------------------------------------------------------------------------ */
size_t glass_csr_build_workspace_bytes(int64_t nnz, int64_t n_node);
int glass_csr_build(const int64_t* edge_index, const float* edge_weight, int64_t nnz, int64_t n_node,
                    int aggr, int32_t* rowptr, int32_t* col, float* val, int32_t* rowptr_t,
                    int32_t* col_t, float* val_t, float* deg, int64_t* nnz_out_host, void* workspace,
                    size_t workspace_bytes, void* stream);

/* to_undirected on the device (reference datasets.py:68-71 -> PyG to_undirected + coalesce: the step that sorts
 * and de-duplicates the edge list before it ever reaches buildAdj).  out_index int64 [2, 2*nnz] and out_w fp32
 * [2*nnz] are capacities; rows of the result are out_index[0 .. m) and out_index[2*nnz .. 2*nnz + m) with
 * m = *nnz_out_host, sorted by (row, col), duplicate weights added in input order.  *already_host = 1 when the
 * input already was undirected and duplicate-free (the reference leaves such a graph untouched).  Init path: one
 * host synchronisation. */
size_t glass_to_undirected_workspace_bytes(int64_t nnz);
int glass_to_undirected(const int64_t* edge_index, const float* edge_weight, int64_t nnz, int64_t n_node,
                        int64_t* out_index, float* out_w, int64_t* nnz_out_host, int* already_host,
                        void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * adj @ x  (impl/models.py:164; backward = the same kernel on the transposed CSR)
 * y[r, :] = sum_{e in row r} val[e] * x[col[e], :]  (deterministic: fixed summation order per shape).
 * Graphs that fill the GPU (n_rows * 32 lanes > about two waves): one fp32 FMA chain per output in CSR
 * order, bit-equal to a sequential CPU loop.  Smaller graphs: 32/G neighbour slots per row (G = h/4
 * feature lanes), slot s adds every (32/G)-th entry, the slot sums are added by a fixed butterfly.
 * n_rows = rows of the CSR block; n_cols = rows of x (every column index must be < n_cols).
 * Fused GraphNorm statistics (impl/models.py:164-165): with stats_partial != NULL the epilogue also emits, per
 * CTA, the fp64 column sums of y and y^2 -- stats_partial[(which*h + c)*stats_ld + blk], blk < *stats_nblk_host
 * (written on the host before the call returns; stats_ld >= glass_spmm_stats_ld() always suffices) -- which
 * glass_graphnorm_stats turns into the normalisation constants without another pass over y.
 * ------------------------------------------------------------------------------------------ */
int glass_spmm_stats_ld(void);
/* Run-time tuning knobs for benchmark scripts ("spmm_variant", "spmm_waves"); the defaults are the measured best. */
int glass_tune(const char* name, int value);
int glass_spmm_csr(const int32_t* rowptr, const int32_t* col, const float* val, const float* x,
                   int64_t ldx, float* y, int64_t ldy, int64_t n_rows, int64_t n_cols, int h,
                   double* stats_partial, int stats_ld, int* stats_nblk_host, void* stream);

/* Skewed (power-law) graphs: rows longer than max_len entries are split into several work items whose
 * partial sums go to a scratch matrix [n_slots, h] and are added up in chunk order (deterministic).
 * Init path (reads rowptr back to the host): glass_spmm_plan_size returns the array sizes,
 * glass_spmm_plan_build fills item_begin/item_end/item_dst [n_items] (heavy items first; item_dst >= 0 is
 * a row of y, < 0 is scratch row -1-item_dst; ordinary rows follow longest first) and long_row/long_slot/long_cnt [n_long].
 * glass_spmm_csr_planned is glass_spmm_csr driven by that plan (capturable; scratch from the caller). */
int glass_spmm_plan_size(const int32_t* rowptr, int64_t n_rows, int max_len, int64_t* n_items_host,
                         int64_t* n_long_host, int64_t* n_slots_host, void* stream);
int glass_spmm_plan_build(const int32_t* rowptr, int64_t n_rows, int max_len, int32_t* item_begin,
                          int32_t* item_end, int32_t* item_dst, int32_t* long_row, int32_t* long_slot,
                          int32_t* long_cnt, void* stream);
int glass_spmm_csr_planned(const int32_t* col, const float* val, const float* x, int64_t ldx, float* y,
                           int64_t ldy, int64_t n_rows, int64_t n_cols, int h, const int32_t* item_begin,
                           const int32_t* item_end, const int32_t* item_dst, int64_t n_items,
                           const int32_t* long_row, const int32_t* long_slot, const int32_t* long_cnt,
                           int64_t n_long, float* scratch, double* stats_partial, int stats_ld,
                           int* stats_nblk_host, void* stream);

/* y (+)= A x for plain (item_begin == NULL) or planned CSR: the column-partitioned phases of one product
 * (glass_b200/partition.py multiplies a peer's columns as soon as that peer's feature shard has arrived). */
int glass_spmm_csr_acc(const int32_t* rowptr, const int32_t* col, const float* val, const float* x, int64_t ldx,
                       float* y, int64_t ldy, int64_t n_rows, int64_t n_cols, int h, const int32_t* item_begin,
                       const int32_t* item_end, const int32_t* item_dst, int64_t n_items, const int32_t* long_row,
                       const int32_t* long_slot, const int32_t* long_cnt, int64_t n_long, float* scratch,
                       int accumulate, void* stream);

/* Measurement aid (bench.py `roofline.l2_gather`): `gathers` pseudo-random 4*h-byte rows of x[n_rows, ldx] are read
 * by lane groups exactly as glass_spmm_csr gathers neighbour rows (float4 per lane, 8 loads in flight) and summed
 * into sink[grid * 256]; no index or value stream, no dependent FMA chain per row -- the L2 -> SM gather rate the
 * SpMM could reach at best on this device.  h in {32, 64, 128}. */
int glass_l2_gather_probe(const float* x, int64_t ldx, int64_t n_rows, int h, int64_t gathers, float* sink,
                          int64_t sink_elems, void* stream);

/* Measurement aid: the same gathers from a 24-bit two-plane copy of x (hi: top 16 bits, lo: next 8 mantissa bits of every
 * fp32; convert != 0 fills the planes from x first).  What a reduced gather format would buy the L2-bound SpMM. */
int glass_l2_gather_probe24(const float* x, int64_t n_rows, int h, int64_t gathers, uint16_t* hi, uint8_t* lo,
                            float* sink, int64_t sink_elems, int convert, void* stream);

/* Measurement aid: the fp32 gather probe with 256-bit loads (h / 8 lanes per row instead of h / 4). */
int glass_l2_gather_probe256(const float* x, int64_t ldx, int64_t n_rows, int h, int64_t gathers, float* sink,
                             int64_t sink_elems, int unroll, void* stream);

/* Sparse label correction for multi-label-batch evaluation (reference
 * impl/train.py:20-34 evaluates every label batch with a full adj @ x).  For fixed weights the mixed features of two
 * label batches differ only on the labelled rows (impl/models.py:161-162): x_b = U + [mask] * delta, hence
 *   y[i,:] = base[i,:] + sum_{e in row i, mask[col[e]] != 0} val[e] * delta[col[e], :]     with base = adj @ U
 * computed ONCE per evaluation epoch.  Streams 4 bytes per stored entry (the column index) and gathers only for
 * labelled neighbours.  Either rowptr (item_begin == NULL) or a row-split plan as for glass_spmm_csr_planned;
 * h % 4 == 0, h <= 128.  Optional statistics epilogue as in glass_spmm_csr. */
int glass_spmm_delta(const int32_t* rowptr, const int32_t* col, const float* val, const uint8_t* mask,
                     const float* delta, int64_t ldd, const float* base, int64_t ldb, float* y, int64_t ldy,
                     int64_t n_rows, int h, const int32_t* item_begin, const int32_t* item_end,
                     const int32_t* item_dst, int64_t n_items, const int32_t* long_row, const int32_t* long_slot,
                     const int32_t* long_cnt, int64_t n_long, float* scratch, double* stats_partial, int stats_ld,
                     int* stats_nblk_host, void* stream);

/* ------------------------------------------------------------------------------------------
 * Label-mixed pair of Linear layers  (impl/models.py:158-162 with activation, :169-173 without)
 *   p0 = act([a1|a2] W0^T + b0), p1 = act([a1|a2] W1^T + b1)
 *   out = mask ? z*p1 + (1-z)*p0 : z*p0 + (1-z)*p1
 * a1 [n,k1], a2 [n,k2] (a2 may be NULL with k2 = 0: the virtual concat of models.py:167),
 * w0,w1 [h, k1+k2] row-major (nn.Linear layout), b0,b1 [h], mask uint8 [n] (1 = labelled node).
 * acts [n, 2h] (ld 2h) receives the post-activation p0|p1 when non-NULL (saved for backward;
 * pass NULL for act == NONE or inference).  path selects SIMT fp32 or tcgen05 (3xTF32).
 * ------------------------------------------------------------------------------------------ */
int glass_pair_linear_mix_fwd(const float* a1, int64_t lda1, int k1, const float* a2, int64_t lda2, int k2,
                              const float* w0, const float* b0, const float* w1, const float* b1,
                              const uint8_t* mask, float z_ratio, int act, float* out, int64_t ldo,
                              float* acts, int64_t n, int h, int path, void* stream);

/* The same with operands that are NORMALISED WHILE THEY ARE LOADED (north star: "fused with the label mix,
 * GraphNorm and the concat", impl/models.py:165-167 and :249-251): for an operand with a statistics table,
 *   a[r, c] <- keep/(1-p) * act(scale[c]*(a[r, c] - am[c]) + bias[c])
 * with rows 0, 1, 4 of the stats [6, k] table of glass_graphnorm_* and the packed keep bits of that call, so the
 * GraphNorm output / dropout / concat never exist in memory.  n1 / n2 may be NULL (plain operand).  Needs the
 * tcgen05 kernels: h in {64, 128}, k1 and k2 multiples of 32, k1+k2 <= 128 (glass_pair_norm_operand_supported). */
typedef struct {
    const float* stats;     /* device [6, k] (NULL: operand used as is) */
    const uint32_t* bits;   /* device packed keep bits, bit r*k + c (NULL: no dropout) */
    float drop_p;           /* probability the bits were drawn with (0: none) */
    int act;                /* glass_act applied between the norm and the dropout */
} glass_norm_operand;
int glass_pair_norm_operand_supported(int k1, int k2, int h);
int glass_pair_linear_mix_fwd_ex(const float* a1, int64_t lda1, int k1, const float* a2, int64_t lda2, int k2,
                                 const float* w0, const float* b0, const float* w1, const float* b1,
                                 const uint8_t* mask, float z_ratio, int act, float* out, int64_t ldo,
                                 float* acts, int64_t n, int h, int path, const glass_norm_operand* n1,
                                 const glass_norm_operand* n2, void* stream);

/* Backward of the above.  dout [n,h]; acts as saved by fwd (NULL iff act == NONE).
 * da1 [n,k1], da2 [n,k2] (either may be NULL to skip), dw0,dw1 [h,k1+k2], db0,db1 [h].
 * workspace: glass_pair_linear_mix_bwd_workspace_bytes(n, h, k1+k2) bytes (split-N partials,
 * reduced in a fixed order => deterministic). */
size_t glass_pair_linear_mix_bwd_workspace_bytes(int64_t n, int h, int k);
int glass_pair_linear_mix_bwd(const float* dout, int64_t lddo, const float* acts, const float* a1,
                              int64_t lda1, int k1, const float* a2, int64_t lda2, int k2,
                              const float* w0, const float* w1, const uint8_t* mask, float z_ratio, int act,
                              float* da1, int64_t ldda1, float* da2, int64_t ldda2, float* dw0, float* db0,
                              float* dw1, float* db1, int64_t n, int h, void* workspace,
                              size_t workspace_bytes, int path, void* stream);
/* ..._ex: n1 / n2 as in fwd_ex (the dW kernel re-creates the normalised operand on load; da1 / da2 are the
 * gradients with respect to the NORMALISED operands); accumulate_da1 / accumulate_da2 != 0 add to da1 / da2
 * instead of overwriting them (an operand that fed two GEMMs, e.g. x_ of impl/models.py:158 and :167). */
int glass_pair_linear_mix_bwd_ex(const float* dout, int64_t lddo, const float* acts, const float* a1,
                                 int64_t lda1, int k1, const float* a2, int64_t lda2, int k2,
                                 const float* w0, const float* w1, const uint8_t* mask, float z_ratio, int act,
                                 float* da1, int64_t ldda1, float* da2, int64_t ldda2, float* dw0, float* db0,
                                 float* dw1, float* db1, int64_t n, int h, void* workspace,
                                 size_t workspace_bytes, int path, const glass_norm_operand* n1,
                                 const glass_norm_operand* n2, int accumulate_da1, int accumulate_da2,
                                 void* stream);

/* ------------------------------------------------------------------------------------------
 * GraphNorm over the whole graph (PyG GraphNorm with batch=None; call sites impl/models.py:165,
 * 249, 257, 266), optionally followed by the activation and dropout that the reference applies
 * right after it (impl/models.py:166, 251, 258-259).
 *   mu = mean_rows(x); o = x - mean_scale*mu; var = mean_rows(o^2); rstd = 1/sqrt(var+eps)
 *   out = keep/(1-p) * act(weight*o*rstd + bias)
 * stats [6, c]: rows = scale (weight*rstd), am (mean_scale*mu), mu, rstd, bias, dropout call id -- written
 * by fwd, consumed by bwd and by every fused consumer.
 * Dropout (drop_p > 0), three sources of the keep decision:
 *   keep uint8 [n,c] (ld c)  explicit mask (tests inject one to compare train-mode passes with the oracle);
 *   rng + bits               rng = {seed, call counter, ticket} (3 x uint64, device); the finalize kernel draws
 *                            this call's keep bits ONCE (Philox4x32-10, counter = element index / 4 and the call
 *                            id) into `bits` (glass_dropout_bits_bytes(n, c) bytes, bit r*c+col of a row-major
 *                            bit string, 1 = keep) and advances the counter; apply, backward and the operand
 *                            loaders of the pair GEMM read one bit per element;
 *   rng alone                the same bits regenerated inside every kernel (no buffer; the one-launch cluster
 *                            kernel for matrices of at most 48 K elements always works this way).
 * Column sums are accumulated in fp64 per block and reduced in block order (deterministic).
 * workspace: glass_graphnorm_workspace_bytes(n, c).
 * Matrices of at most 48 K elements (c <= 256) run as ONE launch on a thread-block cluster (partials
 * exchanged through distributed shared memory); larger ones as three (sums, finalise + bits, element-wise).
 * glass_graphnorm_launches(n, c) tells which (1 or 3; pure host arithmetic, for launch accounting).
 * ------------------------------------------------------------------------------------------ */
size_t glass_graphnorm_workspace_bytes(int64_t n, int c);
size_t glass_dropout_bits_bytes(int64_t n, int c);
int glass_graphnorm_launches(int64_t n, int c);
int glass_graphnorm_fwd(const float* x, int64_t ldx, const float* weight, const float* bias,
                        const float* mean_scale, float eps, int act, const uint8_t* keep, float drop_p,
                        unsigned long long* rng, uint32_t* bits, float* out, int64_t ldo, float* stats,
                        int64_t n, int c, void* workspace, size_t workspace_bytes, void* stream);
/* dx [n,c]; dweight, dbias, dmean_scale [c] are OVERWRITTEN (not accumulated). */
int glass_graphnorm_bwd(const float* dout, int64_t lddo, const float* x, int64_t ldx, const float* weight,
                        const float* mean_scale, const float* stats, int act, const uint8_t* keep,
                        float drop_p, const unsigned long long* rng, const uint32_t* bits, float* dx,
                        int64_t lddx, float* dweight, float* dbias, float* dmean_scale, int64_t n, int c,
                        void* workspace, size_t workspace_bytes, void* stream);

/* Fused forms (north star: "adj@x fused with the label mix, GraphNorm and the concat", impl/models.py:164-167).
 * The column sums come out of the kernel that PRODUCES x (glass_spmm_csr_stats, the *_ex pair GEMM) as fp64
 * partials partial[(which*c + col)*ldp + blk], which = 0: sum x, 1: sum x^2, blk < nblk.
 *   glass_graphnorm_stats      partials -> stats (+ this call's dropout bits); ONE small launch
 *   glass_graphnorm_apply      out = keep/(1-p) * act(scale*(x-am)+bias) from stats (when x must be materialised)
 *   glass_graphnorm_bwd_from_sums  backward when the producer of the gradient already wrote
 *                              u = dout*keep/(1-p)*act'(pre) and the partials of S1 = sum u, S2 = sum u*yhat
 *                              (dX epilogue of glass_pair_linear_mix_bwd_ex): finalise + dx = a*u + b*yhat + g.
 *                              u and dx may alias.  workspace >= 3*c floats (256-byte aligned size). */
int glass_graphnorm_stats(const double* partial, int nblk, int ldp, const float* weight, const float* bias,
                          const float* mean_scale, float eps, const uint8_t* keep, float drop_p,
                          unsigned long long* rng, uint32_t* bits, float* stats, int64_t n, int c, void* stream);
int glass_graphnorm_apply(const float* x, int64_t ldx, const float* stats, int act, const uint8_t* keep,
                          float drop_p, const uint32_t* bits, float* out, int64_t ldo, int64_t n, int c,
                          void* stream);
int glass_graphnorm_bwd_from_sums(const double* partial, int nblk, int ldp, const float* u, int64_t ldu,
                                  const float* x, int64_t ldx, const float* weight, const float* mean_scale,
                                  const float* stats, float* dx, int64_t lddx, float* dweight, float* dbias,
                                  float* dmean_scale, int64_t n, int c, void* workspace, size_t workspace_bytes,
                                  void* stream);

/* Two-phase forms for ROW-PARTITIONED graphs (stress config): a rank computes the partial
 * column sums of its own rows, the caller adds them across ranks (one 2c-value fp64 all-reduce) and passes the
 * totals back as a one-block table together with the GLOBAL row count n_total.
 *   forward : glass_graphnorm_partials     -> all-reduce -> glass_graphnorm_stats(n = n_total) + glass_graphnorm_apply
 *   backward: glass_graphnorm_bwd_partials -> all-reduce -> glass_graphnorm_bwd_finish
 * ldp >= glass_graphnorm_partials_ld(); *nblk_host receives the number of blocks written (host, synchronously).
 * With generator dropout the bits drawn by glass_graphnorm_stats cover the rank's own n rows. */
int glass_graphnorm_partials_ld(void);
int glass_graphnorm_partials(const float* x, int64_t ldx, int64_t n, int c, double* partial, int ldp,
                             int* nblk_host, void* stream);
int glass_graphnorm_bwd_partials(const float* dout, int64_t lddo, const float* x, int64_t ldx, const float* stats,
                                 int act, const uint8_t* keep, float drop_p, const uint32_t* bits, int64_t n, int c,
                                 double* partial, int ldp, int* nblk_host, void* stream);
int glass_graphnorm_bwd_finish(const double* partial, int nblk, int ldp, int64_t n_total, const float* dout,
                               int64_t lddo, const float* x, int64_t ldx, const float* weight,
                               const float* mean_scale, const float* stats, int act, const uint8_t* keep,
                               float drop_p, const uint32_t* bits, float* dx, int64_t lddx, float* dweight,
                               float* dbias, float* dmean_scale, int64_t n, int c, void* workspace,
                               size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * nn.Embedding lookup (impl/models.py:248) and its dense gradient.
 * ids int64 [n]; table [rows, h]; out [n, h].  bwd ACCUMULATES into dtable (caller zero-fills).
 * ------------------------------------------------------------------------------------------ */
int glass_embedding_fwd(const float* table, const int64_t* ids, float* out, int64_t ldo, int64_t n,
                        int64_t rows, int h, void* stream);
int glass_embedding_bwd(const float* dout, int64_t lddo, const int64_t* ids, float* dtable, int64_t n,
                        int64_t rows, int h, void* stream);
/* Deterministic variant (bit-reproducible run to run): rows are added in the order of a stable sort by id.  Plan
 * (built once per id tensor by the host layer): perm int64 [n] = stable argsort of ids; runs = stretches of at most
 * 128 sorted positions holding ONE id (run_begin / run_end int32 [n_runs]); for every distinct id uid[u] its first
 * run and run count.  run_sum [n_runs, h] is scratch.  Rows of dtable whose id does not occur are NOT written
 * (caller zero-fills). */
int glass_embedding_bwd_ordered(const float* dout, int64_t lddo, const int64_t* perm, const int32_t* run_begin,
                                const int32_t* run_end, int64_t n_runs, const int64_t* uid,
                                const int32_t* uid_first_run, const int32_t* uid_runs, int64_t n_uid,
                                float* run_sum, float* dtable, int64_t rows, int h, void* stream);

/* The model's LAST GraphNorm feeds nothing but the pooling (impl/models.py:266 / :272 -> GLASS.Pool :346-350), so the two
 * run as one operator on the norm's INPUT x and the statistics table of glass_graphnorm_stats: forward normalises only
 * the gathered rows (the [n_node, d] output of the norm is never written) and saves ysum[b, :] = sum_v (x_v - am) rstd;
 * backward writes every row of dx = alpha u + beta yhat + gamma (u = pooled gradient scattered in subgraph order, the
 * column sums S1 / S2 are formed from the B x d pooled gradients and ysum) and the norm's parameter gradients.
 * sum / mean / size pooling, d <= 256; scratch as for glass_segment_pool_bwd. */
int glass_norm_pool_fwd(const float* x, int64_t ldx, const float* stats, const int64_t* pos, int64_t b, int64_t lmax,
                        int mode, float* out, int64_t ldo, float* cnt, float* ysum, int d, int64_t n_node, void* stream);
int glass_norm_pool_bwd(const float* dout, int64_t lddo, const int64_t* pos, int64_t b, int64_t lmax, int mode,
                        const float* cnt, const float* ysum, const float* x, int64_t ldx, const float* stats,
                        const float* weight, const float* mean_scale, float* dx, int64_t lddx, float* dweight,
                        float* dbias, float* dmean_scale, int d, int64_t n_node, void* scratch, size_t scratch_bytes,
                        void* stream);

/* ------------------------------------------------------------------------------------------
 * Padded-subgraph pooling (GLASS.Pool impl/models.py:346-350 = pad2batch + emb[pos] + pool_fn;
 * pools impl/models.py:295-319).  pos int64 [b, lmax], -1 = padding.  One CTA per subgraph reads
 * the padded row directly; row lanes accumulate strided rows in pad order and are combined in a
 * fixed order (deterministic).  out [b, d]; cnt fp32 [b] (valid nodes per row); argmax int32 [b, d] (MAX only,
 * else may be NULL): node id of the first maximum, -1 for an empty row.
 * bwd, scratch == NULL: ACCUMULATES into demb [n_node, d] (caller zero-fills) with atomics (order of the
 *   additions for a node that occurs in several subgraphs is not reproducible);
 * bwd, scratch != NULL (glass_segment_pool_bwd_scratch_bytes(b, n_node) bytes): deterministic -- WRITES every row
 *   of demb (zeros outside the batch); per-subgraph membership bitmaps are built first, then a node's contributions
 *   are added in ascending subgraph order by one thread per column.
 * ------------------------------------------------------------------------------------------ */
int glass_segment_pool_fwd(const float* emb, int64_t lde, const int64_t* pos, int64_t b, int64_t lmax,
                           int mode, float* out, int64_t ldo, float* cnt, int32_t* argmax, int d,
                           int64_t n_node, void* stream);
int glass_segment_pool_bwd(const float* dout, int64_t lddo, const int64_t* pos, int64_t b, int64_t lmax,
                           int mode, const float* cnt, const int32_t* argmax, float* demb, int64_t ldde,
                           int d, int64_t n_node, void* scratch, size_t scratch_bytes, void* stream);
size_t glass_segment_pool_bwd_scratch_bytes(int64_t b, int64_t n_node);
/* PoolModule.forward(x, batch) (impl/models.py:287-292): x [m, d] rows already gathered,
 * batch int64 [m] sorted ascending (as pad2batch produces). */
int glass_segment_pool_batch_fwd(const float* x, int64_t ldx, const int64_t* batch, int64_t m, int64_t n_seg,
                                 int mode, float* out, int64_t ldo, float* cnt, int32_t* argmax, int d,
                                 void* stream);
int glass_segment_pool_batch_bwd(const float* dout, int64_t lddo, const int64_t* batch, int64_t m,
                                 int64_t n_seg, int mode, const float* cnt, const int32_t* argmax, float* dx,
                                 int64_t lddx, int d, void* stream);

/* ------------------------------------------------------------------------------------------
 * Max-zero-one labels (impl/utils.py:32-45) and the bool mask of impl/models.py:246.
 * z int64 [n_node] is fully overwritten: 1 where the node occurs in pos, else 0.
 * mask uint8 [n_node] = (z > 0)  ( == z > 0.5 for integer z).
 * pad2batch (impl/utils.py:18-29): batch_out/pos_out capacity b*lmax; *n_valid (device int64).
 * ------------------------------------------------------------------------------------------ */
int glass_maxzoz(const int64_t* pos, int64_t n_pos, int64_t* z, uint8_t* mask_or_null, int64_t n_node,
                 void* stream);
int glass_label_mask(const int64_t* z, uint8_t* mask, int64_t n_node, void* stream);
int glass_pad2batch(const int64_t* pad, int64_t b, int64_t lmax, int64_t* batch_out, int64_t* pos_out,
                    int64_t* n_valid, void* stream);

/* ------------------------------------------------------------------------------------------
 * Multi-tensor Adam in one launch (optimizer of GLASSTest.py:213 for the captured train step; same update
 * as torch.optim.Adam without amsgrad).  table: n_tensors rows {float* p, const float* g, float* m, float* v,
 * int64 n} in device memory; the host splits every tensor into chunks of glass_adam_chunk() elements and
 * passes, per chunk, the tensor index and the first element.  lr: 1 float and state: 2 floats {step count, 0}
 * in device memory (the kernel advances the step count, so a captured graph can be replayed).
 * ------------------------------------------------------------------------------------------ */
int glass_adam_chunk(void);
int glass_adam_step(const void* table, const int32_t* chunk_tensor, const int64_t* chunk_begin, int64_t n_chunks,
                    const float* lr, float* state, float beta1, float beta2, float eps, float weight_decay,
                    void* stream);

/* ------------------------------------------------------------------------------------------
 * Label-batch data parallelism (the reference has no distributed code) without a library
 * collective in the step: reduce-scatter + Adam + all-gather of the embedding table in ONE launch over NVLink peer
 * memory.  Every rank owns a symmetric (peer-mapped) block [table | table gradient | small gradients | flags];
 * peer_base is a DEVICE array [world] with this process' mapping of every rank's block, the off_* are byte offsets
 * into a block (16-byte aligned; the flag area holds glass_dp_flags_bytes() zero-initialised bytes).
 * The launch waits until every peer's gradients are complete, averages the table gradient of the element range
 * [own_begin, own_end) in rank order straight from the peers' blocks, applies Adam (same rule / lr / step count as
 * glass_adam_step, which must run AFTER it in the same stream for the remaining parameters) and stores the new
 * values into every rank's table; averages the small-gradient blocks into small_out; returns (kernel end) only when
 * every peer has finished writing this rank's table.  *error is set to 1 if a peer does not answer within ~2 s.
 * ------------------------------------------------------------------------------------------ */
int glass_dp_flags_bytes(void);
int glass_dp_adam_step(int world, int rank, const unsigned long long* peer_base, unsigned long long off_table,
                       unsigned long long off_grad, unsigned long long off_small, unsigned long long off_flags,
                       int64_t table_elems, int64_t own_begin, int64_t own_end, int64_t small_elems, float* m,
                       float* v, float* small_out, const float* lr, const float* state, float beta1, float beta2,
                       float eps, float weight_decay, unsigned long long* epoch, unsigned* ticket, int* error,
                       void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GLASS_B200_H_ */
