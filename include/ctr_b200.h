/* ctr_b200.h -- C-ABI of libctr_b200.so: sm_100a kernels for the torchctr embedding /
 * feature-interaction hot path.
 *
 * Every entry point is `extern "C"`, takes plain device pointers and sizes (no torch
 * types), runs on the CUDA stream passed as the last argument (a `cudaStream_t`, spelled
 * `void*` so that C callers need no CUDA headers), never allocates user-visible memory
 * (scratch comes from a caller-provided workspace) and returns 0 or a negative
 * `CTR_E_*` code; `ctr_last_error_string()` describes the last failure of the calling
 * thread.  No C++ exception crosses the boundary.
 *
 * What each entry point replaces in the reference (paths relative to the torchctr repo):
 *
 *   ctr_emb_pool_fwd        torchctr/models/dnn.py:53-59  (mask, nn.Embedding gather,
 *                           mask-multiply, sum(dim=1)) and the concat of dnn.py:61-67;
 *                           with index_kind HASH also torchctr/utils.py:103-119 applied
 *                           per element by torchctr/transformer.py:487-490; with
 *                           index_kind REMAP also transformer.py:492-498.
 *   ctr_hash_bucket_i64     torchctr/utils.py:103-119 (hash_bucket) on decimal ids.
 *   ctr_emb_bwd_plan /      autograd of dnn.py:57-58 (aten::embedding_dense_backward) and
 *   ctr_emb_bwd_apply       optimizer.step() at torchctr/trainer.py:303, restricted to the
 *                           rows the batch touched; no dense [V, D] gradient exists.
 *   ctr_rows_gather /       nn.Embedding.forward as used by DynamicEmbedding,
 *   ctr_normal_fill_rows    torchctr/nn/embedding.py:69-87 (growth: N(0, std) new rows).
 *   ctr_vocab_*             vocabulary fit / transform, transformer.py:451-498.
 *   ctr_fm_fwd / ctr_fm_bwd FM second-order term (absent from the reference).
 *   ctr_cross_*             DCN-v2 cross layer (absent from the reference).
 */
#ifndef CTR_B200_H
#define CTR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CTR_B200_ABI_VERSION 4
#define CTR_MAX_FEATURES 40 /* features per launch group (kernel-parameter budget) */

/* status codes */
#define CTR_OK 0
#define CTR_E_BADARG (-1)    /* null pointer, bad size, unsupported dim ...            */
#define CTR_E_WORKSPACE (-2) /* workspace too small                                    */
#define CTR_E_CUDA (-3)      /* a CUDA runtime call failed (see ctr_last_error_string) */
#define CTR_E_UNSUPPORTED (-4)

/* bits of the device-side status word (ctr_group_t.status, optional) */
#define CTR_STATUS_INDEX_OOB 1u   /* a table index was >= num_rows (IndexError upstream) */
#define CTR_STATUS_MAP_FULL 2u    /* vocabulary map ran out of slots                      */
#define CTR_STATUS_NO_RUNS 4u     /* unique-row outputs asked of a plan built with CTR_PLAN_NO_RUNS */

/* ctr_feature_t.index_kind */
#define CTR_INDEX_DIRECT 0 /* ids are table rows                                          */
#define CTR_INDEX_HASH 1   /* row = murmur3_32(decimal(id), seed) % num_rows               */
#define CTR_INDEX_REMAP 2  /* row = vocab[id], unknown -> 0 (OOV)                          */
#define CTR_INDEX_WINDOW 3 /* row = id - first for ids in [first, first + num_rows), every other id is padding;
                              `first` travels in hash_seed.  One table cut into a head [0, K) and a tail [K, V) that
                              live in different places (hybrid placement: hot rows replicated, the rest row-sharded):
                              the two parts are two features over the SAME ids and the same output columns -- in a
                              single-id group exactly one of them writes the bag's slice (the head also writes the
                              zeros of a negative id) */

/* ctr_feature_t.pooling */
#define CTR_POOL_SUM 0
#define CTR_POOL_MEAN 1 /* sum / max(#valid ids, 1): extension, not in the reference     */

/* optimizer kinds for ctr_emb_bwd_apply */
#define CTR_OPT_NONE 0            /* only emit unique rows + summed gradients               */
#define CTR_OPT_SGD 1             /* torch.optim.SGD(lr)                                    */
#define CTR_OPT_ADAGRAD 2         /* torch.optim.Adagrad(lr, eps) element-wise              */
#define CTR_OPT_ROWWISE_ADAGRAD 3 /* one accumulator per row, += mean(g*g)                  */
#define CTR_OPT_ADAM 4            /* torch.optim.SparseAdam (lazy Adam)                     */
#define CTR_OPT_GRAD_OUT 5        /* no update: state0[row, :] = summed gradient of the row (FM part included),
                                     twin_state0[row] = gradient of the twin.  For tables REPLICATED on every rank of a
                                     multi-GPU job: the caller zero-fills the buffers, all-reduces them and applies a
                                     dense update (ctr_dense_adagrad), so every replica moves identically.          */

/* Device-resident open-addressing hash map: raw key (int64) -> table row (int32). */
typedef struct ctr_vocab_map {
    int64_t *keys;      /* [capacity], CTR_VOCAB_EMPTY where unused                        */
    int32_t *rows;      /* [capacity]                                                      */
    int64_t capacity;   /* power of two                                                    */
} ctr_vocab_map_t;
#define CTR_VOCAB_EMPTY INT64_MIN

/* One sparse feature (= one table) of a launch group.  All pointers are device pointers. */
typedef struct ctr_feature {
    const int64_t *ids;      /* [B, L] row-major; negative = padding (dataset.py:54-57)     */
    const float *id_weight;  /* [B, L] per-id multiplier or NULL                             */
    float *table;            /* [num_rows, D] fp32, row-major                                */
    float *state0;           /* optimizer state: Adagrad sum [V,D] | row-wise [V] | Adam m   */
    float *state1;           /* Adam v [V,D] or NULL                                         */
    float *bag_scale;        /* [B]: written by fwd, read by bwd when pooling == MEAN        */
    const ctr_vocab_map_t *map; /* host pointer to a map descriptor, index_kind == REMAP     */
    int64_t num_rows;        /* V (< 2^31)                                                   */
    int32_t L;               /* id slots per bag                                             */
    int32_t D;               /* embedding dim: multiple of 4 up to 128, or any D <= 32       */
    int32_t out_col;         /* first column of this feature in the pooled / grad matrix     */
    int32_t pooling;         /* CTR_POOL_*                                                   */
    int32_t index_kind;      /* CTR_INDEX_*                                                  */
    uint32_t hash_seed;
    /* optional twin: a second, ONE-column table indexed by the same ids -- DeepFM's first-order weights
     * w1_f[id] (oracle/models.py OracleDeepFM.linear_embeddings).  Its pooled value is added to
     * group->extra[bag]; the fused update moves it with the main row (same optimizer, same hyper-parameters). */
    float *twin_table;       /* [num_rows] fp32 or NULL                                      */
    float *twin_state0;      /* optimizer state of the twin, same kinds as state0, [num_rows] */
    float *twin_state1;      /* Adam v of the twin [num_rows] or NULL                        */
} ctr_feature_t;

/* A launch group: up to CTR_MAX_FEATURES features pooled for the same B bags into one
 * row-major matrix out[B, out_stride] (the tower input of dnn.py:67), plus an optional
 * dense block copied into columns [dense_col, dense_col + dense_width). */
typedef struct ctr_group {
    const ctr_feature_t *features; /* host array                                            */
    int32_t num_features;
    int32_t B;
    float *out;              /* fwd: pooled output; bwd: gradient w.r.t. it                  */
    int64_t out_stride;      /* floats per row of `out`                                      */
    const float *dense;      /* [B, dense_width] or NULL                                     */
    int32_t dense_width;
    int32_t dense_col;
    int32_t zero_from;       /* fwd: columns [zero_from, out_stride) are zero-filled; <0 = none */
    uint32_t *status;        /* device status word (CTR_STATUS_*), may be NULL               */
    /* optional per-bag scalar fused into the lookup / update (single-id groups of one width only, i.e. Criteo-shaped
     * DeepFM):  extra[b] = sum_f twin_f[id_bf]  +  (fm ? 0.5 * sum_d[(sum_f v_bfd)^2 - sum_f v_bfd^2] : 0)
     * -- the first-order and FM second-order logit terms.  fwd writes extra and, when fm is set,
     * fm_sum[b, :] = sum_f v_bf (f32 [B, D]); bwd reads extra as dL/d extra[b] and fm_sum, adds
     * extra[b] * (fm_sum[b, :] - v_bf) to the row gradients and extra[b] to the twin gradients.  NULL = off. */
    float *extra;            /* [B]                                                          */
    float *fm_sum;           /* [B, D], needed when fm != 0                                  */
    int32_t fm;
    /* backward only (ctr_emb_bwd_apply): `out` holds the gradient COLUMN-BLOCKED instead of row-major -- the columns
     * [out_col, out_col + D) of feature i are the contiguous matrix out + out_col * B, f32 [B, D] -- which is what
     * ctr_linear_fwd_blocked(block_cols = D, block_stride = B * D) writes.  A table's gradient slices are then one dense
     * 4 D B-byte block that stays in L2 while its rows are swept, instead of D-float pieces strided over [B, out_stride].
     * Single-id, sum-pooled features of one width (D % 4 == 0), out_col multiples of D. */
    int32_t grad_blocked;
} ctr_group_t;

/* Per-step scalars of the row update as the kernels consume them (fp32). */
typedef struct ctr_hyper {
    float lr, eps, one_minus_beta1, one_minus_beta2, adam_step_size;
} ctr_hyper_t;

typedef struct ctr_opt {
    int32_t kind;            /* CTR_OPT_*                                                    */
    int32_t step;            /* Adam bias-correction step (1-based)                          */
    double lr;               /* already decayed by the host if a schedule is used            */
    double eps;
    double beta1, beta2;     /* Adam (doubles: torch forms 1 - beta in double precision)     */
    const ctr_hyper_t *device_hyper; /* optional DEVICE copy of ctr_opt_hyper()'s output, read by the kernels
                                        at run time instead of the values above: lets a captured CUDA graph
                                        follow lr schedules / Adam steps without re-capture               */
} ctr_opt_t;

/* host helper: the fp32 scalars the kernels use for `opt` (what device_hyper must hold) */
void ctr_opt_hyper(const ctr_opt_t *opt, ctr_hyper_t *out);

const char *ctr_last_error_string(void);
int ctr_abi_version(void);
/* number of kernels this library has launched since it was loaded (bench.py reports it) */
int64_t ctr_kernel_launches(void);

/* ---- lookup + pool ---------------------------------------------------------------- */
int ctr_emb_pool_fwd(const ctr_group_t *group, void *stream);

/* ids i64 [n] -> out i32 [n] = murmur3_32(decimal ASCII of id, seed) % buckets */
int ctr_hash_bucket_i64(const int64_t *ids, int64_t n, uint32_t buckets, uint32_t seed, int32_t *out,
                        void *stream);

/* the same for string categories: string i = data[offsets[i], offsets[i + 1]) (utf-8 bytes of the canonical form of
 * torchctr/transformer.py:367-401), offsets i64 [n + 1]; out i32 [n] = murmur3_32(bytes, seed) % buckets */
int ctr_hash_bucket_bytes(const uint8_t *data, const int64_t *offsets, int64_t n, uint32_t buckets, uint32_t seed,
                          int32_t *out, void *stream);

/* plain row gather out[i, :] = table[ids[i], :] (ids >= 0), used by DynamicEmbedding */
int ctr_rows_gather(const int64_t *ids, int64_t n, const float *table, int64_t num_rows, int32_t D,
                    float *out, uint32_t *status, void *stream);

/* table[row0 .. row0 + n) ~ N(mean, std) from a counter-based generator (seed, row, col) */
int ctr_normal_fill_rows(float *table, int64_t row0, int64_t n, int32_t D, float mean, float std,
                         uint64_t seed, void *stream);

/* the same generator for a SHARD: local rows [row0, row0 + n) of `table` stand for the global rows global_first,
 * global_first + global_stride, ...; they get exactly the values ctr_normal_fill_rows gives those rows of the whole
 * table (shard-native initialisation: no rank ever holds a full table) */
int ctr_normal_fill_rows_strided(float *table, int64_t row0, int64_t n, int32_t D, float mean, float std, uint64_t seed,
                                 int64_t global_first, int64_t global_stride, void *stream);

/* min and max of an id tensor, written to out[0], out[1] (device) */
int ctr_ids_minmax(const int64_t *ids, int64_t n, int64_t *out, void *stream);

/* ---- backward: dedup/sort plan, then fused gradient scatter + optimizer ------------- */
/* Bytes of workspace ctr_emb_bwd_plan needs for this group. */
int64_t ctr_emb_bwd_workspace_bytes(const ctr_group_t *group);

/* Sorts the (row, slot) pairs of the group and finds the runs of equal rows.  The plan
 * lives inside `workspace` and stays valid until the workspace is reused.  It depends on
 * the ids only, so it can be applied to several tables that share those ids. */
int ctr_emb_bwd_plan(const ctr_group_t *group, void *workspace, int64_t workspace_bytes, void *stream);
/* Same with flags.  CTR_PLAN_NO_RUNS: sort only, skip the list of runs -- enough for a fused update that asks for
 * no unique-row outputs (uniq_row == NULL), which is what a training step does. */
#define CTR_PLAN_NO_RUNS 1u
int ctr_emb_bwd_plan_ex(const ctr_group_t *group, void *workspace, int64_t workspace_bytes, uint32_t flags,
                        void *stream);

/* For every unique row touched: g = sum over its slots of coef * grad_out[bag, cols];
 * then the optimizer update of that row, in place.  `group->out` is grad_out here and the
 * features' `table/state/D/out_col` may differ from the ones the plan was built with
 * (same ids, index mapping and num_rows).  Optional outputs (may be NULL):
 *   uniq_feature i32 [U], uniq_row i32 [U], row_grad f32 [U, row_grad_stride], num_unique i64 [1]. */
int ctr_emb_bwd_apply(const ctr_group_t *group, void *workspace, const ctr_opt_t *opt,
                      int32_t *uniq_feature, int32_t *uniq_row, float *row_grad, int64_t row_grad_stride,
                      int64_t *num_unique, void *stream);

/* Dense counterpart of the fused row update for tables REPLICATED on every rank (hybrid placement, see
 * ctr_emb_pool_fwd_sharded): params / grads / state0 f32 [n] (n % 4 == 0, 16-byte aligned), grads = the buffer the
 * CTR_OPT_GRAD_OUT sweep filled and the ranks all-reduced.  opt->kind: CTR_OPT_SGD or CTR_OPT_ADAGRAD, same arithmetic as
 * ctr_emb_bwd_apply (optimizer.step() of torchctr/trainer.py:303 on the touched rows); zero gradients are skipped, and with
 * clear_grads the buffer is zero again afterwards. */
int ctr_rows_dense_apply(const ctr_opt_t *opt, float *params, float *grads, float *state0, int64_t n, int32_t clear_grads,
                         void *stream);

/* ---- vocabulary (dynamic growth) ---------------------------------------------------- */
int64_t ctr_vocab_fit_workspace_bytes(int64_t n);
/* keys i64 [n] (negative = padding): admit keys with batch count >= min_freq that are not in
 * the map, give them rows next_row[0], next_row[0]+1, ... in order of first occurrence, and
 * advance next_row[0] (device int64).  counts (i64 [>= final next_row], may be NULL) gets the
 * batch count added for every key that is in the map after the call. */
int ctr_vocab_fit(const ctr_vocab_map_t *map, const int64_t *keys, int64_t n, int32_t min_freq,
                  int64_t *next_row, int64_t *counts, uint32_t *status, void *workspace,
                  int64_t workspace_bytes, void *stream);
/* keys -> rows (i32), unknown or negative -> oov_row */
int ctr_vocab_transform(const ctr_vocab_map_t *map, const int64_t *keys, int64_t n, int32_t oov_row,
                        int32_t *rows, void *stream);
/* insert n distinct (key, row) pairs that are not in the map (rebuild after a capacity change) */
int ctr_vocab_insert(const ctr_vocab_map_t *map, const int64_t *keys, const int32_t *rows, int64_t n,
                     uint32_t *status, void *stream);
int ctr_vocab_clear(const ctr_vocab_map_t *map, void *stream);

/* ---- feature interaction ------------------------------------------------------------- */
/* x f32 [B, x_stride] holding F fields of D columns starting at column 0;
 * out[b * out_stride] (+)= 0.5 * sum_d[(sum_f v)^2 - sum_f v^2] (+ sum of `first` [B, nfirst]). */
int ctr_fm_fwd(const float *x, int64_t x_stride, int32_t B, int32_t F, int32_t D, const float *first,
               int32_t nfirst, int64_t first_stride, float *out, int64_t out_stride, int32_t accumulate,
               void *stream);
/* gx[b, f*D+d] (+)= gout[b] * (sum_f' v[b,f',d] - v[b,f,d]);  gfirst[b, j] = gout[b] */
int ctr_fm_bwd(const float *x, int64_t x_stride, int32_t B, int32_t F, int32_t D, const float *gout,
               int64_t gout_stride, float *gx, int64_t gx_stride, int32_t accumulate, float *gfirst,
               int32_t nfirst, int64_t gfirst_stride, void *stream);

/* Target-attention pooling over a candidate sequence, torchctr/nn/functional.py:46-74 (target_attention):
 * out[b] = sum_n softmax_n(cand[b, n] . target[b]) cand[b, n];  target f32 [B, E], cand f32 [B, N, E] (contiguous), mask f32
 * [B, N] or NULL, E a power of two in [4, 128].  honor_mask = 0 reproduces the reference, whose masked_fill is not in
 * place (functional.py:63) so the mask has no effect; honor_mask = 1 gives masked candidates weight 0 (the intended
 * semantics; a row with every candidate masked yields zeros).  scores [B, N], row_max [B], row_sum [B] are kept for bwd.
 * bwd: gtarget [B, E], gcand [B, N, E] from gout [B, E]. */
int ctr_target_attention_fwd(const float *target, const float *cand, const float *mask, int32_t B, int32_t N, int32_t E,
                             int32_t honor_mask, float *out, float *scores, float *row_max, float *row_sum, void *stream);
int ctr_target_attention_bwd(const float *target, const float *cand, const float *mask, int32_t B, int32_t N, int32_t E,
                             int32_t honor_mask, const float *out, const float *scores, const float *row_max,
                             const float *row_sum, const float *gout, float *gtarget, float *gcand, void *stream);

/* DCN-v2 cross epilogue: y = x0 * (u + bias) + x   (u = x W^T from the GEMM), all [B, d] */
int ctr_cross_combine_fwd(const float *x0, const float *x, const float *u, const float *bias, int32_t B,
                          int32_t d, int64_t stride, float *y, void *stream);
/* given gy: gu = gy * x0 (its column sums are the bias gradient);  gx0 (+)= gy * (u + bias) */
int ctr_cross_combine_bwd(const float *x0, const float *u, const float *bias, const float *gy, int32_t B,
                          int32_t d, int64_t stride, float *gu, float *gx0, int32_t accumulate_gx0, void *stream);

/* ---- dense layers on tensor cores (tcgen05.mma kind::tf32, fp32 accumulate in TMEM) ----------------
 * C[M, N] = act(A[M, K] . W[N, K]^T + bias[N]); act 0 = none, 1 = relu.  Row-major, leading dimensions in
 * floats; lda and ldw multiples of 4, A and W 16-byte aligned; bias may be NULL.  Replaces the Linear
 * layers of torchctr/models/dnn.py:35-46 and the x.W^T of the DCN-v2 cross layer. */
int ctr_linear_fwd(const float *A, int64_t lda, const float *W, int64_t ldw, const float *bias, float *C,
                   int64_t ldc, int32_t M, int32_t N, int32_t K, int32_t act, void *stream);
/* the same with C stored column-blocked: columns [j * block_cols, (j + 1) * block_cols) are the contiguous matrix
 * C + j * block_stride, f32 [M, block_cols] row-major (block_cols a multiple of 4 that divides N; M > 128).  Used for the
 * input gradient of the tower's first Linear (autograd of torchctr/models/dnn.py:68), which the embedding update consumes
 * feature by feature (ctr_group_t.grad_blocked). */
int ctr_linear_fwd_blocked(const float *A, int64_t lda, const float *W, int64_t ldw, const float *bias, float *C,
                           int32_t block_cols, int64_t block_stride, int32_t M, int32_t N, int32_t K, int32_t act, void *stream);

/* The same GEMM (no activation) that also leaves, per block of 32 output rows, the column sums of C and of C^2 in
 * stats f32 [ctr_linear_stats_blocks(M)][2][N]: the batch statistics of the BatchNorm1d that follows the Linear in
 * torchctr/models/dnn.py:39-41 come out of the GEMM epilogue instead of a second pass over C
 * (-> ctr_bn_stats_from_partials).  N % 4 == 0; ctr_linear_stats_blocks returns 0 when M is too small for this path. */
int32_t ctr_linear_stats_blocks(int32_t M);
int ctr_linear_fwd_stats(const float *A, int64_t lda, const float *W, int64_t ldw, const float *bias, float *C, int64_t ldc,
                         int32_t M, int32_t N, int32_t K, float *stats, int64_t stats_floats, void *stream);

/* Weight gradient of the same layers: dW[N, K] = sum_b G[b, N] . X[b, K] (G = dL/dz, X = the layer input), TF32 in,
 * fp32 accumulate, the batch split over the SMs and the slices added in order.  ldg, ldx multiples of 4 floats,
 * G, X, workspace 16-byte aligned; workspace: ctr_linear_wgrad_workspace_bytes(B, N, K). */
int64_t ctr_linear_wgrad_workspace_bytes(int32_t B, int32_t N, int32_t K);
int ctr_linear_wgrad(const float *G, int64_t ldg, const float *X, int64_t ldx, int32_t B, int32_t N, int32_t K, float *dW,
                     int64_t lddw, void *workspace, int64_t workspace_bytes, void *stream);

/* Error-compensated TF32 ("3xTF32"): the exact-fp32 mode of the two GEMMs above (parity bound 1e-5 of the tower,
 * torchctr/models/dnn.py:35-46).  Splits x f32 [rows, cols] into TF32-representable parts hi = rna(x), lo = rna(x - hi)
 * and writes three segments along the reduction axis of the GEMM that will consume them:
 *   role 0 (left operand) lo | hi | hi,   role 1 (right operand) hi | lo | hi   (small terms first, see split_tf32.cu);
 *   axis 1: out [rows, >= 3 * seg], segment k in columns [k * seg, k * seg + cols), seg = cols rounded up to 4 (tail = 0)
 *           -> ctr_linear_fwd(A3, W3, K = 3 * seg);
 *   axis 0: out [3 * rows, >= seg], segment k in rows [k * rows, (k + 1) * rows)
 *           -> ctr_linear_wgrad(G3, X3, B = 3 * rows).
 * A3 . W3^T = lo.hi + hi.lo + hi.hi with fp32 accumulation: error ~2^-21 relative per product. */
int ctr_split_tf32(const float *x, int64_t ldx, int32_t rows, int32_t cols, float *out, int64_t ldo, int32_t axis,
                   int32_t role, void *stream);

/* ---- tower block: BatchNorm1d (training) + ReLU + Dropout around a Linear, fused --------------------------
 * Replaces the BatchNorm1d / ReLU / Dropout modules of torchctr/models/dnn.py:39-45 in training mode and the
 * bias-gradient reduction of the Linear in front of them.  z = x W^T + b is f32 [B, N] (N multiple of 4, <= 1024),
 * row pitches in floats, multiples of 4; every pointer 16-byte aligned.  workspace: ctr_tower_workspace_bytes(N). */
int64_t ctr_tower_workspace_bytes(int32_t N);
/* batch statistics of z: mean[N], rstd[N] = 1 / sqrt(biased var + eps); running_mean / running_var (may be NULL)
 * and num_batches_tracked (may be NULL) are updated as torch.nn.BatchNorm1d does with `momentum`. */
int ctr_bn_stats(const float *z, int64_t ldz, int32_t B, int32_t N, float eps, float momentum, float *mean, float *rstd,
                 float *running_mean, float *running_var, int64_t *num_batches_tracked, void *workspace, void *stream);
/* ctr_bn_stats on partial sums produced by ctr_linear_fwd_stats (blocks = ctr_linear_stats_blocks(B)) */
int ctr_bn_stats_from_partials(const float *partial, int32_t blocks, int32_t B, int32_t N, float eps, float momentum, float *mean,
                               float *rstd, float *running_mean, float *running_var, int64_t *num_batches_tracked, void *stream);
/* y = dropout_p(relu((z - mean) * rstd * gamma + beta)).  The keep mask is a function of (*seed_dev, seed_offset,
 * element index) and is not stored; the backward call must get the same three values. */
int ctr_bn_relu_dropout_fwd(const float *z, int64_t ldz, int32_t B, int32_t N, const float *mean, const float *rstd,
                            const float *gamma, const float *beta, float p_drop, const uint64_t *seed_dev,
                            uint64_t seed_offset, float *y, int64_t ldy, void *stream);
/* given gy = dL/dy: gz = dL/dz (through dropout, ReLU and the batch statistics), dgamma[N], dbeta[N], and
 * dbias[N] = column sums of gz (may be NULL). */
int ctr_bn_relu_dropout_bwd(const float *gy, int64_t ldgy, const float *z, int64_t ldz, int32_t B, int32_t N,
                            const float *mean, const float *rstd, const float *gamma, const float *beta, float p_drop,
                            const uint64_t *seed_dev, uint64_t seed_offset, float *gz, int64_t ldgz, float *dgamma,
                            float *dbeta, float *dbias, void *workspace, void *stream);
/* dbias[N] of a ctr_bn_relu_dropout_bwd(..., dbias = NULL, workspace) call, from the per-block partials that call left in
 * `workspace` (same B, N).  Nothing on the way to dL/dx needs the bias gradient (torchctr/trainer.py:301-303: it is read by
 * optimizer.step() only), so it may be issued on another stream, ordered after that call and before the next use of the
 * workspace. */
int ctr_bn_bias_grad_from_partials(const void *workspace, int32_t B, int32_t N, float *dbias, void *stream);

/* ---- logit head: the last Linear(H -> 1) of the tower (dnn.py:46) + extra logit terms + the mean
 * binary_cross_entropy_with_logits of dnn.py:75, forward and backward.  h f32 [B, H] (H a power of two in [4, 128],
 * row pitch multiple of 4), w [H], bias [1] or NULL, extra [B] (stride in floats) or NULL, labels [B] (stride).
 * fwd: logits[B] (may be NULL), dz[B] = (sigmoid(z) - y) / B, loss[1] = mean loss.  bwd, given the DEVICE scalar
 * gscale = dL/dloss: gh[B, H] = gscale dz w (may be NULL), gw[H], gb[1] (may be NULL), gextra[B] (may be NULL).
 * workspace: ctr_tower_workspace_bytes(H). */
int ctr_logit_bce_fwd(const float *h, int64_t ldh, int32_t B, int32_t H, const float *w, const float *bias,
                      const float *extra, int64_t extra_stride, const float *labels, int64_t label_stride, float *logits,
                      float *dz, float *loss, void *workspace, void *stream);
int ctr_logit_bce_bwd(const float *h, int64_t ldh, int32_t B, int32_t H, const float *w, const float *dz,
                      const float *gscale, float *gh, int64_t ldgh, float *gw, float *gb, float *gextra,
                      int64_t gextra_stride, void *workspace, void *stream);

/* The same with a second linear term folded in: z += xe[b, :ne] . we + be (ne <= 32) -- DeepFM's Linear(Nd, 1) over the dense
 * features -- and its weight gradient gwe[ne] (its bias gradient equals gb). */
int ctr_logit_bce_fwd_ex(const float *h, int64_t ldh, int32_t B, int32_t H, const float *w, const float *bias, const float *extra,
                         int64_t extra_stride, const float *xe, int64_t ldxe, int32_t ne, const float *we, const float *be,
                         const float *labels, int64_t label_stride, float *logits, float *dz, float *loss, void *workspace,
                         void *stream);
int ctr_logit_bce_bwd_ex(const float *h, int64_t ldh, int32_t B, int32_t H, const float *w, const float *dz, const float *gscale,
                         float *gh, int64_t ldgh, float *gw, float *gb, float *gextra, int64_t gextra_stride, const float *xe,
                         int64_t ldxe, int32_t ne, float *gwe, void *workspace, void *stream);
/* ctr_logit_bce_bwd / _bwd_ex accept gw = NULL: only gh / gextra are produced and the per-block partials stay in `workspace`;
 * this call turns them into gw[H], gb[1] (may be NULL) and gwe[ne] (ne = 0: none) -- the parameter gradients of the last
 * Linear (dnn.py:46), needed by optimizer.step() only, so it may run on another stream (ordered after the bwd call, before the
 * next use of the workspace). */
int ctr_logit_bce_bwd_finalize(int32_t B, int32_t H, const float *gscale, float *gw, float *gb, int32_t ne, float *gwe,
                               const void *workspace, void *stream);

/* torch.optim.Adagrad (lr_decay = 0, weight_decay = 0) over `count` <= 48 dense fp32 tensors in ONE launch:
 * sum += g*g; p -= lr * g / (sqrt(sum) + eps).  params / grads / sums / sizes are HOST arrays of device pointers / sizes. */
int ctr_dense_adagrad(int32_t count, float *const *params, const float *const *grads, float *const *sums,
                      const int64_t *sizes, float lr, float eps, void *stream);

/* ---- row-sharded tables: pack / unpack around the NCCL all-to-all ------------------------------
 * (the reference is replicas-only, torchctr/trainer.py:128-130).  owner(row) = row mod world; on the owner
 * all tables of the group live in one fused shard, row base[owner * num_features + table] + row / world. */
int64_t ctr_route_workspace_bytes(const ctr_group_t *group, int32_t world);
/* For every id slot of the group (features in order, slot = bag * L + l): bucket by owner (stable),
 *   counts    i64 [world + 1]  slots per owner (last = padding / invalid),
 *   send_rows i64 [S]          virtual rows in exchange order (first sum(counts[:world]) entries valid),
 *   inv       i64 [S]          slot -> position in exchange order, -1 for padding (feature-major, ready to be
 *                              used as the ids of a pooled lookup over the received vectors). */
int ctr_route_build(const ctr_group_t *group, int32_t world, const int64_t *base, int64_t *counts,
                    int64_t *send_rows, int64_t *inv, void *workspace, int64_t workspace_bytes, void *stream);
/* g_send[k, :] = grad_out[bag(slot_k), columns of feature(slot_k)] for the first n exchange positions
 * (group->out is grad_out; every feature of the group has dim D). */
int ctr_route_grad_gather(const ctr_group_t *group, int32_t world, const void *workspace, int64_t n, int32_t D,
                          float *g_send, void *stream);

/* ---- row-sharded tables over peer memory (NVLink P2P inside the kernels, no all-to-all) -----------
 * One process per GPU.  Every rank allocates its table shard, gradient buffer and routing lists with
 * ctr_peer_alloc, exports them (ctr_peer_export -> 64-byte handle, exchanged by the host, e.g. with
 * torch.distributed.all_gather_object) and maps the other ranks' buffers with ctr_peer_open.  The kernels
 * below then read peer pointers directly; a host-side barrier (any collective on the stream) separates the
 * phases.  Table f's row r lives on rank (r + f) mod world, at row  adj[owner * F + f] + (r + f) / world  of
 * that rank's fused shard (all tables of the group back to back, same D). */
#define CTR_MAX_WORLD 16
#define CTR_PEER_HANDLE_BYTES 64
int ctr_peer_alloc(int64_t bytes, void **ptr);          /* zero-filled device memory other processes can map */
int ctr_peer_free(void *ptr);
int ctr_peer_export(void *ptr, void *handle_out);       /* handle_out: CTR_PEER_HANDLE_BYTES bytes             */
int ctr_peer_open(const void *handle, void **ptr);      /* in ANOTHER process than the one that exported       */
int ctr_peer_close(void *ptr);

typedef struct ctr_shard {
    int32_t world, rank;
    const int64_t *adj;        /* DEVICE i64 [world * num_features], see above                                  */
} ctr_shard_t;

/* ctr_emb_pool_fwd with every row read from its owner: tables[o] = fused shard of rank o (device pointers, the
 * local one included).  `num_rows` is the size of the FULL table.  A feature whose `table` is NULL is row-sharded; a
 * feature that carries a `table` (and optionally a `twin_table`) is REPLICATED -- this rank holds all of it and reads it
 * locally (hybrid placement: small tables replicated, large ones sharded; the reference replicates everything,
 * torchctr/trainer.py:128-130).  shard->adj is indexed with the feature's position in THIS group. */
int ctr_emb_pool_fwd_sharded(const ctr_group_t *group, const ctr_shard_t *shard, const float *const *tables,
                             void *stream);
/* The same with the fused per-bag scalar (group->extra / fm_sum / fm, single-id groups of one width): twin_tables[o] =
 * the one-column twin shard of rank o (same row geometry as tables[o]), or NULL when the sharded features have no twins. */
int ctr_emb_pool_fwd_sharded_ex(const ctr_group_t *group, const ctr_shard_t *shard, const float *const *tables,
                                const float *const *twin_tables, void *stream);

/* Buckets this rank's id slots by owner (stable).  Outputs, to be placed in peer-visible memory:
 *   counts u32 [CTR_MAX_WORLD + 1]  slots per owner (entry `world` = padding / invalid),
 *   keys   u32 [S]  virtual row inside the owner's fused shard, owner-major, slot order inside an owner,
 *   slots  u32 [S]  feature-relative slot (bag * L + l) of the same position. */
int64_t ctr_route_p2p_workspace_bytes(const ctr_group_t *group);
int ctr_route_p2p_build(const ctr_group_t *group, const ctr_shard_t *shard, uint32_t *counts, uint32_t *keys,
                        uint32_t *slots, void *workspace, int64_t workspace_bytes, void *stream);

/* Owner side of the backward.  `group` describes THIS rank's shard: feature f = the local rows of table f
 * (table / state pointers into the fused shard, num_rows = local row count, L / out_col / D as on the
 * requesters), B = per-rank batch.  The plan pulls the (key, slot) lists addressed to this rank from every
 * peer's routing buffers (peer_counts / peer_keys / peer_slots: host arrays [world] of device pointers) and sorts
 * them; the number of pairs stays on the device.  capacity_slots = world * S bounds it. */
int64_t ctr_emb_bwd_p2p_workspace_bytes(const ctr_group_t *group, int32_t world);
int ctr_emb_bwd_plan_p2p(const ctr_group_t *group, const ctr_shard_t *shard, const uint32_t *const *peer_counts,
                         const uint32_t *const *peer_keys, const uint32_t *const *peer_slots, void *workspace,
                         int64_t workspace_bytes, void *stream);
/* Fused reduce + row update of the local shard; the gradient of a slot is read from the grad_out matrix of the
 * rank that sent it: peer_grads[r] f32 [B, group->out_stride] (group->out is ignored). */
int ctr_emb_bwd_apply_p2p(const ctr_group_t *group, const ctr_shard_t *shard, void *workspace, const ctr_opt_t *opt,
                          const float *const *peer_grads, int64_t *num_unique, void *stream);
/* The same with the fused DeepFM terms (see ctr_group_t.extra): peer_extra[r] f32 [B] = dL/d extra of rank r's bags,
 * peer_fm_sum[r] f32 [B, D] = the fm_sum rank r's lookup wrote (NULL array = no FM term); the owner's features carry
 * their twin shard slices (twin_table / twin_state0 / twin_state1).  group->extra must be non-NULL (it is not read). */
int ctr_emb_bwd_apply_p2p_ex(const ctr_group_t *group, const ctr_shard_t *shard, void *workspace, const ctr_opt_t *opt,
                             const float *const *peer_grads, const float *const *peer_extra,
                             const float *const *peer_fm_sum, int64_t *num_unique, void *stream);
/* Requester side of the same: packed[b, j * D .. (j + 1) * D) = gx[b, cols[j] .. cols[j] + D) + extra_grad[b] * fm_sum[b, :]
 * for the J sharded features of a bag (cols: HOST array of their first columns in gx; fm_sum NULL = plain copy).  With the
 * gradients packed like this the owners call ctr_emb_bwd_apply_p2p_ex with peer_grads = the packed matrices (out_col = j * D,
 * out_stride = J * D), peer_fm_sum = NULL and group->fm = 1: one D-float piece per slot crosses NVLink instead of the
 * gradient slice plus the bag's fm_sum row; only the "- row * sum(extra_grad)" part of the FM gradient is left to the owner. */
int ctr_fm_pack_grads(const float *gx, int64_t gx_stride, const float *extra_grad, const float *fm_sum, int32_t B, int32_t D,
                      const int32_t *cols, int32_t J, float *packed, void *stream);

/* ---- de-duplicated exchange (every distinct row of a rank's batch crosses NVLink once per direction) -----------
 * Requester, forward: plan the group with ctr_emb_bwd_plan (runs listed), then ctr_unique_fetch copies the row of every
 * run from its owner into staging[run, :] (f32 [S, D]) and, when uidx != NULL, writes uidx[slot] = run (i64 [S],
 * feature-major slots, -1 = padding): a pooled lookup over `staging` with `uidx` as ids gives the pooled output.
 * Requester, backward: ctr_emb_bwd_apply with CTR_OPT_NONE on the same plan leaves uniq_feature / uniq_row /
 * row_grad [U, D] / num_unique in peer-visible buffers.
 * Owner: ctr_emb_bwd_plan_p2p_unique turns every rank's (feature, row) list into (row key | invalid, rank:index) pairs
 * and sorts them (count on the device, capacity = world * S pairs); ctr_emb_bwd_apply_p2p_unique reduces and updates,
 * reading gradient row `index` of peer_row_grads[rank] (f32 [S, group->out_stride]).  The owner group has L = 1,
 * out_col = 0, tables = this rank's shard slices. */
int ctr_unique_fetch(const ctr_group_t *group, const ctr_shard_t *shard, const float *const *tables, void *plan_workspace,
                     float *staging, int64_t *uidx, void *stream);
int64_t ctr_emb_bwd_p2p_unique_workspace_bytes(const ctr_group_t *group, int64_t capacity);
int ctr_emb_bwd_plan_p2p_unique(const ctr_group_t *group, const ctr_shard_t *shard, const int64_t *const *peer_num_unique,
                                const int32_t *const *peer_uniq_feature, const int32_t *const *peer_uniq_row, int64_t capacity,
                                void *workspace, int64_t workspace_bytes, void *stream);
int ctr_emb_bwd_apply_p2p_unique(const ctr_group_t *group, const ctr_shard_t *shard, void *workspace, const ctr_opt_t *opt,
                                 const float *const *peer_row_grads, int64_t capacity, int64_t *num_unique, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* CTR_B200_H */
