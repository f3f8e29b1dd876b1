/* arx_b200.h — C ABI of the B200-native A-RecSys hot path (libarx_b200.so).
 *
 * The reference (skywaLKer518/A-Recsys) is pure Python over TensorFlow-1 ops; it has
 * no FFI of its own.  Each entry point below replaces the TF-op call sites of one
 * row of SURVEY.md section 8(a) (file:line relative to the reference root).  The
 * reference-side binding a maintainer would add is the ctypes stub shown in
 * INTEGRATION.md; the in-repo host side is a-recsys_b200/_lib.py.
 *
 * Conventions: all pointers are DEVICE pointers unless the name says host; indices
 * are int32, data fp32 row-major; every call enqueues on `stream` (a cudaStream_t
 * passed as void*) and returns immediately; return value 0 = ok, <0 = ARX_E_*;
 * no global state, caller owns all buffers; nothing throws.
 */
#ifndef ARX_B200_H_
#define ARX_B200_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ARX_OK              0
#define ARX_E_BADARG       -1
#define ARX_E_LAUNCH       -2   /* cudaGetLastError() != cudaSuccess after a launch */
#define ARX_E_UNSUPPORTED  -3
#define ARX_E_CAPACITY     -4

#define ARX_ABI_VERSION     4
#define ARX_MAX_ATTRS      64

/* One attribute (= one embedding table) of one entity side.  Mirrors the per-attribute
 * arrays of attributes/attribute.py:7-47 as uploaded by embed_attribute.py:308-318,
 * plus the variables of embed_attribute.py:265-306. */
typedef struct arx_attr_desc {
  float*         table;      /* E_f  [vocab, dim]                                   */
  float*         table_acc;  /* Adagrad accumulator [vocab, dim] or NULL            */
  float*         bias;       /* beta_f [vocab] or NULL (user side has none)         */
  float*         bias_acc;   /* [vocab] or NULL                                     */
  const int32_t* values;     /* kind 0: features_cat[f]   [N+1]
                                kind 1: features_mulhot[f] flat token ids [nnz+1]   */
  const int32_t* starts;     /* kind 1: mulhot_starts[f]  [N+2]; kind 0: NULL       */
  const int32_t* lengths;    /* kind 1: mulhot_lengths[f] [N+1]; kind 0: NULL       */
  int32_t*       touch;      /* [vocab] zero-initialised scratch owned by the table;
                                left all-zero again by arx_pool_bwd_plan            */
  int64_t        vocab;
  int32_t        kind;       /* 0 = categorical, 1 = multi-hot                      */
  int32_t        reserved;   /* row sharding: (G << 16) | rank; 0 = whole table here.  Row t of the
                                full table lives on rank t % G at local row t / G.     */
  const int32_t* lengths_full; /* NULL, or (kind 1, sharded) the CSR above is PRE-PARTITIONED: values
                                holds only the rows this rank owns, already as local row numbers,
                                starts/lengths describe those local bags (possibly empty), and
                                lengths_full[e] is the bag's full length = the mean divisor.  */
} arx_attr_desc;

#define ARX_POOL_MEAN    0   /* out[n, dim]      = (1/F) sum_f pooled_f   (reduce_mean over attrs,
                                embed_attribute.py:219,:235)                                        */
#define ARX_POOL_CONCAT  1   /* out[n, f*dim ..] = pooled_f               (concat, embed_attribute.py:415;
                                also the per-attribute list when read with a stride)               */

/* K1+K2 — replaces EmbeddingAttribute._get_embedded (attributes/embed_attribute.py:350-417)
 * and the batch_slice2/batch_segids2 index builders (attributes/mulhot_index.py:48-67):
 * gather + segment mean of every attribute of `n` entities, fused mean/concat over
 * attributes and pooled bias.  attrs: device array of n_attr descriptors. */
int arx_pool_fwd(const arx_attr_desc* attrs, int n_attr, int dim,
                 const int32_t* ent_ids, int64_t n,
                 float* out, int64_t out_stride, int mode,
                 float* bias_out /* [n] or NULL */,
                 int max_rows_per_entity /* upper bound of sum_f len_f over the store (static), or 0 =
                                            unknown: enables the flat row-list kernel when <= 1024 */,
                 void* stream);

/* Several independent K1+K2 lookups of one step (users, target items, sampled pool: embed_attribute.py:222-254,
 * :208-220, :479-508 are three separate sub-graphs in the reference) in ONE launch; mean mode, dim 128 / 256.
 * ARX_E_UNSUPPORTED when a request is outside the flat kernel's limits: issue one arx_pool_fwd per lookup then. */
typedef struct arx_pool_req {
  const arx_attr_desc* attrs;      /* first descriptor of the attribute range */
  const int32_t* ent_ids;          /* [n] entity indices */
  float* out;                      /* [n, out_stride] pooled vectors (mean over attributes) */
  float* bias_out;                 /* [n] pooled bias or NULL */
  int64_t n;
  int64_t out_stride;
  int32_t n_attr;
  int32_t max_rows_per_entity;     /* sum over the attributes of the longest bag */
} arx_pool_req;
int arx_pool_fwd_many(const arx_pool_req* reqs, int n_req, int dim, void* stream);

/* K11 — row-sharded tables over NVLink peer memory (the reference is single-GPU, nearest analogue: the
 * tower-wise tf.device placement at lstm/run.py:221-229).  One process per GPU; every rank allocates its receive
 * buffers with arx_peer_alloc, exports them (CUDA IPC, 64-byte handle) and maps the other ranks' (arx_peer_open).
 * arx_pool_fwd_many_push is arx_pool_fwd_many with the PARTIAL pooled vectors of request i ADDED (red.global.add.v4.f32,
 * system scope) into the owner rank's receive block instead of stored locally: entity e belongs to rank
 * e / rows_per_rank, row e % rows_per_rank of peer_out[owner] (row pitch `stride` floats; the pooled bias goes to
 * peer_bias[owner][row]) — lookup + reduce-scatter in one kernel; rows_per_rank = 0 adds every entity into ALL ranks
 * (lookup + all-reduce); the receive blocks must be zero before the step.
 * arx_peer_push_rows copies (mode 0, rank `skip` left out) or adds (mode 1) rows into rows [row0, row0 + rows) of
 * every rank's block (all-gather / all-reduce by push).  arx_peer_barrier: device-side barrier over the G ranks
 * (flags[g] = rank g's block of >= G uint32, epoch = this rank's device counter), stream-ordered and graph-capturable;
 * a rank that waits longer than timeout_ns stores 1 + the missing rank into *err instead of spinning forever. */
typedef struct arx_pool_push {
  float* const* peer_out;          /* device array [n_ranks] of receive-block bases, or NULL: store to arx_pool_req.out */
  float* const* peer_bias;         /* device array [n_ranks] of receive vectors for the pooled bias, or NULL           */
  int64_t rows_per_rank;           /* entities per owner rank; 0: add every entity into ALL ranks (all-reduce by push)  */
  int64_t stride;                  /* row pitch of the receive blocks, floats                                           */
  int32_t n_ranks;
  int32_t reserved;
} arx_pool_push;
/* up to 8 row blocks pushed in one launch (arx_peer_push_many): mode 0 = store to every rank, 1 = add */
typedef struct arx_peer_seg {
  const float* src;                /* [rows, width], row pitch src_stride (floats, % 4) */
  float* const* dst;               /* device array [G] of destination bases             */
  int64_t rows, width, src_stride, dst_stride, row0;
  int32_t mode, reserved;
} arx_peer_seg;
int arx_peer_push_many(const arx_peer_seg* segs, int n_segs, int G, void* stream);
int arx_pool_fwd_many_push(const arx_pool_req* reqs, const arx_pool_push* push, int n_req, int dim, void* stream);
int arx_peer_alloc(int64_t bytes, void** ptr);
int arx_peer_free(void* ptr);
int arx_peer_export(void* ptr, unsigned char* handle64);
int arx_peer_open(const unsigned char* handle64, void** ptr);
int arx_peer_close(void* ptr);
int arx_peer_barrier(uint32_t* const* flags, int rank, int G, uint32_t* epoch, int64_t timeout_ns, int32_t* err, void* stream);
int arx_peer_push_rows(const float* src, int64_t rows, int64_t width, int64_t src_stride, float* const* dst,
                       int64_t dst_stride, int64_t row0, int G, int skip, int mode, void* stream);

/* Integer part of K2 only (mulhot_index.py:48-67): flat token index and segment id
 * vectors for one multi-hot attribute; bit-exact parity target.  offsets[n+1] is an
 * exclusive scan of the bag lengths computed by the caller (device). */
int arx_mulhot_flat_index(const arx_attr_desc* attrs, int attr, const int32_t* ent_ids,
                          int64_t n, const int64_t* offsets, int32_t* flat_idx,
                          int32_t* seg_ids, void* stream);

/* Backward plan: the de-duplicated (table,row) list of a batch of entities and, per
 * row, the bucket of (source position, weight) pairs that contribute to it.  Depends
 * only on ent_ids, so it is cached for the catalog and for the sampled pool. */
typedef struct arx_bwd_plan {
  int32_t* counters;   /* [8]: 0 n_unique, 1 cursor, 2 overflow, 3 n_occ, 4 n_chunks   */
  int32_t* uniq_tok;   /* [cap_rows] token id                                          */
  int32_t* uniq_attr;  /* [cap_rows] index into attrs                                  */
  int32_t* row_base;   /* [cap_rows] first bucket slot                                 */
  int32_t* row_cnt;    /* [cap_rows] bucket size                                       */
  int32_t* bucket_src; /* [cap_occ]  row of the gradient arena that contributes         */
  float*   bucket_w;   /* [cap_occ]  1/(F*len) (mean) or 1/len (concat)                */
  int32_t* chunk_row;  /* [cap_chunks] row of each 64-entry chunk of a hot row         */
  int32_t* row_chunk0; /* [cap_rows]  first chunk of a hot row                         */
  int32_t* row_done;   /* [cap_rows]  zero-initialised arrival counters (self-resetting)*/
  float*   partials;   /* [cap_chunks*(dim+1)] chunk partial sums (+ bias partials)    */
  int64_t  cap_rows;
  int64_t  cap_occ;
  int64_t  cap_chunks; /* >= cap_occ/32 + 1                                            */
} arx_bwd_plan;

/* K2b part 1 — replaces the IndexedSlices bookkeeping of tf.gradients through
 * embedding_lookup/unsorted_segment_sum (hmf/hmf_model.py:149) and the duplicate-index
 * summation of the sparse optimizer apply.  Phased so that several lookups of the same
 * tables in one step (target items + sampled pool; the T input steps of the LSTM) are
 * de-duplicated together:  begin, count*, alloc, fill*, end.  Each count/fill pair
 * names a contiguous attribute range [attr_begin, attr_begin+n_attr) of `attrs`
 * (no_id = skip attribute 0, no_attribute = attribute 0 only; embed_attribute.py:356-373)
 * and the first row (row_base) its gradients occupy in the row arena handed to
 * arx_pool_bwd_apply: MEAN mode -> row_base + i, weight 1/(n_attr*len);
 * CONCAT mode -> row_base + i*n_attr + f, weight 1/len. */
int arx_bwd_plan_begin(arx_bwd_plan plan, void* stream);
int arx_bwd_plan_count(const arx_attr_desc* attrs, int attr_begin, int n_attr,
                       const int32_t* ent_ids, int64_t n, arx_bwd_plan plan, void* stream);
int arx_bwd_plan_alloc(const arx_attr_desc* attrs, arx_bwd_plan plan, void* stream);
int arx_bwd_plan_fill(const arx_attr_desc* attrs, int attr_begin, int n_attr,
                      const int32_t* ent_ids, int64_t n, int mode, int64_t row_base,
                      arx_bwd_plan plan, void* stream);
int arx_bwd_plan_end(const arx_attr_desc* attrs, arx_bwd_plan plan, void* stream);
/* one-shot form for a single lookup over all n_attr attributes (row_base 0) */
int arx_pool_bwd_plan(const arx_attr_desc* attrs, int n_attr, const int32_t* ent_ids,
                      int64_t n, int mode, arx_bwd_plan plan, void* stream);

#define ARX_OPT_ADAGRAD 0   /* tf.train.AdagradOptimizer: acc += g^2; w -= lr*g/sqrt(acc) (acc0 = 0.1) */
#define ARX_OPT_SGD     1   /* tf.train.GradientDescentOptimizer (lstm/seqModel.py:176)              */
#define ARX_OPT_NONE    2   /* write summed row gradients to rows_out instead (oracle check)         */

/* K2b part 2 — segment-reduce the row arena dout [R, dout_stride] (first dim columns of
 * each row) over every bucket and apply the optimizer to the touched rows only
 * (hmf/hmf_model.py:146-151).  dbias: [R] aligned with the arena rows, or NULL;
 * grad_scale_dev (device scalar or NULL) multiplies the summed gradient
 * (clip_by_global_norm, lstm/seqModel.py:180); rows_out/bias_rows_out
 * ([n_unique, dim] / [n_unique]) only for ARX_OPT_NONE. */
int arx_pool_bwd_apply(const arx_attr_desc* attrs, int n_attr, int dim,
                       arx_bwd_plan plan, const float* dout, int64_t dout_stride,
                       const float* dbias, float lr, const float* grad_scale_dev,
                       int opt, float* rows_out, float* bias_rows_out, void* stream);

/* The same optimizer step for up to two table sets in ONE launch (the user and the item tables of a training step:
 * hmf/hmf_model.py:146-151 applies every gradient in one session.run).  ARX_OPT_ADAGRAD / ARX_OPT_SGD, dim <= 128,
 * dim % 4 == 0; ARX_E_UNSUPPORTED otherwise. */
typedef struct arx_apply_set {
  const arx_attr_desc* attrs;      /* all descriptors of the table set */
  const float* dout;               /* gradient arena [R, dout_stride] */
  const float* dbias;              /* Column-slab variant for the gradient of the WHOLE pooled catalog (loss = ce / full-catalog WMRB: every item
 * contributes, ~100 contributions per table row, a 512 MB arena at C2).  Adagrad is element-wise, so the step is done
 * 16 columns at a time: dslab = columns [c0, c0 + 16) of the arena laid out contiguously ([rows][16], slab-major copy made
 * by the caller: L2-resident), every pass gathers from L2 and touches 64 B of each table / accumulator row.  The plan is
 * allocated with an explicit chunk size (arx_bwd_plan_alloc_h instead of arx_bwd_plan_alloc) that the apply calls repeat.
 * dbias is consumed by the pass with c0 == 0.  Same semantics as arx_pool_bwd_apply (hmf/hmf_model.py:146-151). */
int arx_bwd_plan_alloc_h(const arx_attr_desc* attrs, arx_bwd_plan plan, int heavy, void* stream);
int arx_pool_bwd_apply_slab(const arx_attr_desc* attrs, int n_attr, int dim, arx_bwd_plan plan, const float* dslab,
                            int c0, const float* dbias, float lr, const float* grad_scale_dev, int opt, int heavy,
                            void* stream);

/* [R] bias-gradient arena or NULL */
  int64_t dout_stride;
  arx_bwd_plan plan;
  int32_t n_attr;
  int32_t reserved;
} arx_apply_set;
int arx_pool_bwd_apply_many(const arx_apply_set* sets, int n_sets, int dim, float lr, const float* grad_scale_dev,
                            int opt, void* stream);

/* Squared-norm term of tf.clip_by_global_norm (lstm/seqModel.py:180), accumulated into *sumsq
 * (device).  merged = 0: IndexedSlices semantics, sum over OCCURRENCES of ||w * dOut[row]||^2
 * (tables reached only through lookups; TF does not merge duplicate slices for the norm).
 * merged = 1 (dense-gradient semantics: tables that also feed the scoring matmul get ONE dense gradient in
 * TF, duplicates summed before the norm) returns ARX_E_UNSUPPORTED here: write the merged rows with
 * arx_pool_bwd_apply(ARX_OPT_NONE, rows_out, bias_rows_out) and reduce them with arx_rows_sumsq. */
int arx_pool_bwd_sumsq(const arx_attr_desc* attrs, int dim, arx_bwd_plan plan, const float* dout,
                       int64_t dout_stride, const float* dbias, float* sumsq, int merged, void* stream);
/* *sumsq += sum of squares of rows[0 .. n_unique) x dim (+ bias_rows[0 .. n_unique) unless NULL), n_unique read
 * from the plan on the device. */
int arx_rows_sumsq(const float* rows, const float* bias_rows, arx_bwd_plan plan, int dim, float* sumsq,
                   void* stream);

/* K3/K4 dense contraction  C[m,n] = alpha * A[m,k] * op(B) + bias  (fp32 in/out).
 * trans_b = 1: B is [n,k] (scores = U * P^T, embed_attribute.py:171,188 after the
 * pool-first rewrite);  trans_b = 0: B is [k,n];  trans_a = 1: A is [k,m] (dP = D^T U).
 * bias_n [n] (or NULL) is added along columns. */
int arx_gemm(const float* A, const float* B, float* C, int64_t m, int64_t n, int64_t k,
             int trans_a, int trans_b, const float* bias_n, float alpha, float beta,
             void* stream);

/* Same contraction on the 5th-generation tensor cores: tcgen05.mma kind::tf32 (fp32 operands
 * read as tf32, fp32 accumulation in TMEM), TMA-fed 4-stage pipeline, split-K when the output
 * grid cannot fill 148 SMs.  Returns ARX_E_UNSUPPORTED when TMA cannot describe an operand
 * (pitch or base not 16-byte aligned): the caller then uses arx_gemm. */
int arx_gemm_tc(const float* A, const float* B, float* C, int64_t m, int64_t n, int64_t k,
                int trans_a, int trans_b, const float* bias_n, float alpha, float beta,
                void* stream);

/* K3 + K5 fused: full-catalog scoring and softmax cross-entropy without materialising the [M, N]
 * logits (embed_attribute.py:148-206 get_prediction + :530 sparse_softmax_cross_entropy_with_logits,
 * and tf.gradients of both, hmf_model.py:146-151).  logits[r, c] = U[r] . P[c] + beta[c]; tf32 operands
 * (pass arx_round_tf32 copies), fp32 accumulation in TMEM.
 *   arx_ce_fwd : lse[r] = ln sum_c exp(logits[r, c]).  workspace: arx_ce_workspace_floats() floats.
 *                The row loss is lse[r] - logits[r, target[r]] (arx_rowdot_fwd on the gathered target rows).
 *   arx_ce_bwd : with D[r, c] = g[r] * (exp(logits[r, c] - lse[r]) - [c == target[r]]):
 *                dU [M, d] = D P,  dP [N, d] = D^T U,  dbeta [N] = column sums of D (or NULL).
 *                UT [d, M] and PT [d, N] are the transposed operands (arx_transpose).
 * d must be 32, 64, 96 or 128 and M, N multiples of 4; otherwise ARX_E_UNSUPPORTED (the caller then
 * materialises the logits: arx_gemm_tc + arx_loss_rows). */
int arx_ce_workspace_floats(int64_t M, int64_t N, int64_t* n_floats);
int arx_ce_fwd(const float* U, const float* P, const float* beta, int64_t M, int64_t N, int64_t d,
               float* workspace, float* lse, void* stream);
/* loss[r] = lse[r] - (U[r] . P[target[r]] + beta[target[r]])  (embed_attribute.py:530 on the fused path). */
int arx_ce_rowloss(const float* U, const float* P, const float* beta, const int32_t* target, const float* lse,
                   int64_t M, int64_t N, int64_t d, float* loss, void* stream);
int arx_ce_bwd(const float* U, const float* P, const float* UT, const float* PT, const float* beta,
               const float* lse, const float* g, const int32_t* target, int64_t M, int64_t N, int64_t d,
               float* dU, float* dP, float* dbeta, void* stream);

/* dst[c, r] = src[r, c] (fp32).  Stages an MN-major operand K-major for arx_gemm_tc;
 * round_tf32_out != 0 also rounds to the nearest tf32 (see arx_round_tf32). */
int arx_transpose(const float* src, int64_t rows, int64_t cols, float* dst, int round_tf32_out,
                  void* stream);
/* dst = nearest-even tf32 of src (13 low mantissa bits cleared).  tcgen05 kind::tf32 truncates
 * fp32 operands, which is biased; pre-rounding halves the error and makes it zero-mean, keeping
 * logits within the north star's 1e-3 of the fp32 reference. */
int arx_round_tf32(const float* src, float* dst, int64_t n, void* stream);

/* out[c] = sum_r x[r, c] — the item-bias gradient: column sums of d(loss)/d(scores)
 * (the `+ i_biases` terms of embed_attribute.py:171,188). */
int arx_colsum(const float* x, int64_t rows, int64_t cols, int64_t ld, float* out, void* stream);

/* loss kinds: embed_attribute.py:525-649 */
#define ARX_LOSS_CE       0
#define ARX_LOSS_WARP     1   /* = rs + log                                   */
#define ARX_LOSS_RS       2
#define ARX_LOSS_RS_SIG   3
#define ARX_LOSS_RS_SIG2  4
#define ARX_LOSS_BBPR     5
#define ARX_LOSS_MW       6   /* sampled; target score supplied separately     */
/* loss_func transforms: embed_attribute.py:580-592 */
#define ARX_LF_LOG 0
#define ARX_LF_EXP 1
#define ARX_LF_POLY 2
#define ARX_LF_POLY2 3
#define ARX_LF_LINEAR 4
#define ARX_LF_SQUARE 5

/* K5/K6 — per-row loss over materialised scores [mb, V] and, in place, the gradient
 * d(mean loss)/d scores (scaled by row_scale[b], e.g. 1/mb or w_bt/(sum_t w)).
 * target: [mb] column index (CE/WARP/RS*) ; target_score: [mb] (MW) ;
 * positives as CSR (pos_ptr[R+1], pos_idx[L]) of masked columns — replaces the dense
 * bool mask variable of embed_attribute.py:651-672,721-747.  pos_row[mb] (or NULL =
 * identity) selects the CSR row of batch row b (e.g. the user index into a per-user
 * CSR built once by prepare_warp); negative pos_idx entries are ignored; rows must be
 * sorted ascending when V > 32768.
 * Outputs: loss[mb]; dscores (may alias scores, or NULL for forward only);
 * dtarget[mb] (MW only, or NULL); rank_out[mb] (true rank, warp_eval :635-637, or NULL). */
int arx_loss_rows(const float* scores, int64_t mb, int64_t V, int64_t ld,
                  const int32_t* target, const float* target_score,
                  const int32_t* pos_row, const int32_t* pos_ptr, const int32_t* pos_idx,
                  int loss_kind, int loss_func, float exp_p, const float* row_scale,
                  float* loss, float* dscores, float* dtarget, int64_t* rank_out,
                  void* stream);

/* K3 + K6 fused — sampled WMRB loss 'mw' over a pool of N items (embed_attribute.py:641-649 on the scores of
 * :148-206, hmf_model.py:112-130) and its gradients, logits never materialised; same tcgen05 pipeline as arx_ce_*.
 *   hinge[r, c] = max(0, 1 + U[r] . P[c] + beta[c] - tscore[r]) over the columns not excluded for row r
 *   loss[r] = log(1 + hsum[r]),  hsum[r] = sum_c hinge[r, c]
 *   D[r, c] = g[r] / (1 + hsum[r]) * [hinge[r, c] > 0];  dU = D P, dP = D^T U, dbeta = column sums, dts = -row sums
 * mask: dense bit matrix [M, mask_ld] uint32, bit (r, c) = 1 excludes column c for row r (the user's other positives,
 * the reference's tf.where(mask, ...) :646-647); arx_mw_mask_words gives mask_ld for N; arx_mw_mask_build zeroes it
 * and sets the bits from the per-user CSR (pos_row[b] = user of batch row b or NULL, pos_idx = pool position or -1).
 * workspace: arx_ce_workspace_floats(M, N) floats.  Shape limits as arx_ce_*. */
int arx_mw_mask_words(int64_t N, int64_t* words_per_row);
int arx_mw_mask_build(const int32_t* pos_row, const int32_t* pos_ptr, const int32_t* pos_idx, int64_t mb, int64_t N,
                      uint32_t* mask, int64_t mask_ld, void* stream);
int arx_mw_fwd(const float* U, const float* P, const float* beta, const float* tscore, const uint32_t* mask,
               int64_t mask_ld, int64_t M, int64_t N, int64_t d, float* workspace, float* hsum, float* loss,
               void* stream);
int arx_mw_bwd(const float* U, const float* P, const float* UT, const float* PT, const float* beta,
               const float* tscore, const uint32_t* mask, int64_t mask_ld, const float* hsum, const float* g,
               int64_t M, int64_t N, int64_t d, float* dU, float* dP, float* dbeta, float* dts, void* stream);
/* Same, with outputs_zeroed != 0 when the caller has zeroed dU, dP, dbeta and dts itself (off the dependent chain);
 * 2 = always ADD into them (row blocks of one batch accumulating into shared dP / dbeta: full-catalog WMRB). */
int arx_mw_bwd2(const float* U, const float* P, const float* UT, const float* PT, const float* beta,
                const float* tscore, const uint32_t* mask, int64_t mask_ld, const float* hsum, const float* g,
                int64_t M, int64_t N, int64_t d, float* dU, float* dP, float* dbeta, float* dts, int outputs_zeroed,
                void* stream);

/* K4 — target_score[b] = U[b].P[b] + beta[b] (embed_attribute.py:219-220) and its adjoint
 * dU[b] += dts[b] P[b]; dP[b] = dts[b] U[b]. */
int arx_rowdot_fwd(const float* U, const float* P, const float* beta, int64_t mb, int dim,
                   float* out, void* stream);
int arx_rowdot_bwd(const float* U, const float* P, const float* dts, int64_t mb, int dim,
                   float* dU_accum, float* dP, void* stream);

/* K8 — TF-1.0 LSTMCell pointwise stages (lstm/seqModel.py:99-103; gate order i, j, f, o;
 * c' = sigmoid(f + forget_bias) c + sigmoid(i) tanh(j); h' = sigmoid(o) tanh(c')).  The gate
 * pre-activations [x, h] W + b are produced by arx_gemm_tc (x-projection of all T steps as one
 * contraction, h_{t-1} W_h accumulated per step with beta = 1).
 * fwd: Z [mb,4H] pre-activations in -> activated gates out (kept for the adjoint); c_prev may be
 *      NULL (zero initial state, seqModel.py:468).
 * bwd: G [mb,4H] activated gates in -> dZ out; dh = dh_out + dh_rec (either may be NULL),
 *      dc_next may be NULL; writes dc_prev. */
int arx_lstm_gates_fwd(float* Z, const float* c_prev, float* c, float* h, int64_t mb, int H,
                       float forget_bias, void* stream);
int arx_lstm_gates_bwd(float* G, const float* c_prev, const float* c, const float* dh_out,
                       const float* dh_rec, const float* dc_next, float* dc_prev, int64_t mb, int H,
                       void* stream);
/* Same, with the tf32 rounding of the tensor-core operands fused in: fwd2 also writes h_tf32 (the A operand of
 * the next step's h W_h contraction; NULL = skip), bwd2 rounds dZ in place when round_tf32_out != 0 (dZ only
 * feeds contractions).  Saves one arx_round_tf32 launch per time step and direction. */
int arx_lstm_gates_fwd2(float* Z, const float* c_prev, float* c, float* h, float* h_tf32, int64_t mb, int H,
                        float forget_bias, void* stream);
int arx_lstm_gates_bwd2(float* G, const float* c_prev, const float* c, const float* dh_out,
                        const float* dh_rec, const float* dc_next, float* dc_prev, int64_t mb, int H,
                        int round_tf32_out, void* stream);
/* Fused glue of the sampled-WMRB ('mw') step around arx_mw_fwd / arx_mw_bwd (hmf/hmf_model.py:78,112-115;
 * embed_attribute.py:208-220,236):
 *   arx_mw_prep: u = u0 / keep * mask; U_r = tf32(u); UT [d, M] = U_r^T (NULL: skipped);
 *                tscore[r] = u[r] . Pt[r] + bt[r] (Pt NULL: skipped); P_r = tf32(Ps) [S, d]; PT [d, S] = P_r^T.
 *                The dropout mask is either given (`mask`, parity runs), or drawn in the kernel when `rng_state` is
 *                given (device uint64[2] = {seed, step counter}: Philox-4x32-10, element e uses counter (step, e / 4);
 *                tf.nn.dropout keeps with probability keep) and written to `mask_out` for the adjoint, or absent
 *                (both NULL: u = u0).
 *   arx_mw_post: du0 = (dU + dts[r] * Pt) / keep * mask;  dPt = dts[r] * u   (adjoint of the target score + dropout);
 *                advances rng_state[1] when given (the step's draw is over).
 * inv_keep = 1 / keep_prob.  ARX_E_UNSUPPORTED for shapes outside the vector path (caller uses the separate kernels). */
int arx_mw_prep(const float* u0, const float* mask, float inv_keep, const uint64_t* rng_state, float* mask_out,
                const float* Pt, const float* bt, const float* Ps, int64_t M, int64_t S, int d, float* u, float* U_r,
                float* UT, float* tscore, float* P_r, float* PT, void* stream);
int arx_mw_post(const float* dU, const float* dts, const float* Pt, const float* u, const float* mask,
                float inv_keep, int64_t M, int d, float* du0, float* dPt, uint64_t* rng_state, void* stream);

/* K8 as ONE persistent kernel per direction (lstm/seqModel.py:99-103,477: static_rnn over LSTMCell): a cluster of
 * H/32 CTAs owns 128 batch rows for all T steps, W_h resident in shared memory, h W_h on tcgen05 / TMEM, gate
 * non-linearities in the TMEM epilogue, the hidden state all-gathered between the CTAs through distributed shared
 * memory; the backward kernel reduce-scatters the partial dh tiles the same way.
 *   fwd: G [T, mb, 4H] = x-projection + bias on entry, ACTIVATED gates (i, j, f, o) on return; WhT [4H, H] = W_h^T
 *        (tf32-rounded); Hs / Cs [T+1, mb, H]: slots 1..T are written (slot 0 = the caller's zero state); h is stored
 *        tf32-rounded (it only feeds tensor-core contractions).
 *   bwd: G = activated gates on entry, dZ (tf32-rounded) on return; Wh [H, 4H] = W_h (tf32-rounded); Cs from the
 *        forward; dH [T, mb, H] = gradient w.r.t. the step outputs.
 * ARX_E_UNSUPPORTED unless H is 32, 64 or 128: the caller then runs the per-step kernels above. */
int arx_lstm_seq_fwd(float* G, const float* WhT, float* Hs, float* Cs, int64_t T, int64_t mb, int H,
                     float forget_bias, void* stream);
int arx_lstm_seq_bwd(float* G, const float* Wh, const float* Cs, const float* dH, int64_t T, int64_t mb, int H,
                     void* stream);
/* K9 — LSTM / CBOW input mixing: y[r,:] = a*x1[r,:] + b*x2[r % rep,:] (reduce_mean([user_embed,
 * item_embed], 0), lstm/seqModel.py:155; x2 broadcast over the T steps) and the adjoint of the
 * broadcast: out[r,:] = scale * sum_t x[t*rep + r,:]. */
int arx_axpby_rows(const float* x1, const float* x2_rows, float a, float b, int64_t rows,
                   int64_t rep, int dim, float* y, void* stream);
int arx_sum_over_steps(const float* x, int64_t T, int64_t rep, int dim, float scale, float* out,
                       void* stream);

/* SURVEY 8(f) row 1 — batch assembly and negative-pool sampling on the device (integer gathers; the draws that decide
 * WHICH examples go into a batch stay where the reference has them, or use the counter-based generator below).
 * rng_state: device uint64[2] = {seed, step}; every call that draws advances `step` (Philox-4x32-10).
 *   arx_gather_pairs      : LatentProductModel.get_batch / get_permuted_batch (hmf/hmf_model.py:230-260) for given
 *                           interaction indices: out_users[b] = users[idx[b]], out_items[b] = items[idx[b]].
 *   arx_cbow_window_batch : DataIterator.get_next_cbow (word2vec/data_iterator.py:108-169): slot b = the (cursor + b)-th
 *                           non-PAD stream event; ni inputs from the `window` preceding stream positions (distinct when the
 *                           user already has >= ni events in the window).  out_inputs [ni, mb].  window, ni <= 64.
 *   arx_lstm_pad_batch    : SeqModel.get_batch (lstm/seqModel.py:356-404) for given sequence indices sel[b] (-1 = empty
 *                           slot) over a CSR of sequences: time-major inputs / targets / weights [T, mb].
 *   arx_gumbel_keys       : keys[i] = log p[i] - log(-log U_i); the n largest (arx_topk_rows) are a draw without
 *                           replacement with probabilities p = np.random.choice(items, n, False, p)
 *                           (utils/prepare_train.py:7-17). */
int arx_gather_pairs(const int32_t* users, const int32_t* items, const int64_t* idx, int64_t n, int32_t* out_users,
                     int32_t* out_items, void* stream);
int arx_cbow_window_batch(const int32_t* users, const int32_t* items, const int32_t* u_seq_len, const int64_t* targets,
                          int64_t l_seq, int64_t n_targets, int64_t cursor, int mb, int ni, int window,
                          uint64_t* rng_state, int32_t* out_users, int32_t* out_inputs, int32_t* out_targets,
                          void* stream);
int arx_lstm_pad_batch(const int64_t* seq_ptr, const int32_t* seq_items, const int32_t* seq_users, const int64_t* sel,
                       int mb, int T, int start_id, int pad_id, int user_pad_id, int32_t* out_users,
                       int32_t* out_inputs, int32_t* out_targets, float* out_weights, void* stream);
int arx_gumbel_keys(const float* logp, int64_t n, uint64_t* rng_state, float* keys, void* stream);

/* K3m — non-linear attribute pooling of the catalog scores (attributes/embed_attribute.py:194-200, output_feat 2 / 3).
 * S [Vf, mb]: token scores E_f u^T (token-major as the reference's `innerp`), bias [Vf] or NULL; the catalog CSR of the
 * attribute: values (token ids in catalog order) + ptr [V+1] (NULL: one token per item, i.e. a categorical attribute).
 * mode 1 = gather / sum (:172), 2 = segment max (:195), 3 = m + log(1 + sum exp(s - m)) with m the maximum of the WHOLE
 * score matrix (:197-199; arx_score_max computes it, packed with its position).  out [V, mb] += scale * pooled.
 * The adjoint adds scale * dOut through the pooling into dS [Vf, mb] (atomics; for mode 3 also the gradient that reaches
 * the global maximum).  arx_rowsum: out[r] (+)= sum_c X[r, c]  (bias gradients of token-major score gradients). */
int arx_score_max(const float* S, const float* bias, int64_t Vf, int64_t mb, uint64_t* packed, void* stream);
int arx_token_pool_fwd(const float* S, const float* bias, int64_t mb, const int32_t* values, const int64_t* ptr, int64_t V,
                       int mode, const uint64_t* packed_max, float scale, float* out, int32_t* argmax, float* denom,
                       void* stream);
int arx_token_pool_bwd(const float* dOut, const float* S, const float* bias, int64_t mb, const int32_t* values,
                       const int64_t* ptr, int64_t V, int mode, const uint64_t* packed_max, float scale,
                       const int32_t* argmax, const float* denom, float* dS, float* dmax_scratch, void* stream);
int arx_rowsum(const float* X, int64_t rows, int64_t cols, float* out, int accumulate, void* stream);

/* K10 — tf.nn.top_k(sorted=True) over materialised scores (hmf/hmf_model.py:154):
 * descending, ties -> lower index. idx_out [mb,k] int32, val_out [mb,k] or NULL. */
int arx_topk_rows(const float* scores, int64_t mb, int64_t V, int64_t ld, int k,
                  int32_t* idx_out, float* val_out, void* stream);

/* K11 — dense Adagrad / SGD on a flat parameter (hmf_model.py:146-151). */
int arx_dense_update(float* w, float* acc, const float* g, int64_t n, float lr,
                     const float* grad_scale_dev, int opt, void* stream);

/* elementwise helpers used by the towers: dropout with an injected/generated mask
 * (tf.nn.dropout, embed_attribute.py:236): y = x * mask / keep. */
int arx_scale_mask(const float* x, const float* mask, float scale, int64_t n, float* y,
                   void* stream);

/* Launch-shape knobs for measurement sweeps (tools/bench_pool.py): "flat_epb" = 0 (auto) | 1..16 (entities per CTA
 * pass of the flat forward kernel), "apply_ctas_per_sm" = 1..4 (persistent grid of pool_bwd_apply), "plan_agg" = 0 | 1 (block-aggregated plan
 * kernels; needs tables of < 2^27 rows).  No reference
 * counterpart; results are identical for every setting. */
int arx_set_tuning(const char* key, int value);

int arx_abi_version(void);
const char* arx_build_info(void);

#ifdef __cplusplus
}
#endif
#endif /* ARX_B200_H_ */
