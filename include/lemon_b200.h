/*
 * lemon_b200.h — C ABI of liblemon_b200.so: the B200 (sm_100a) implementation of the
 * LEMoN pair-scoring hot path.
 *
 * Reference interfaces replaced (all under MLforHealth/LEMoN):
 *   normalize_vectors                         lib/utils/utils.py:39-40      -> lemon_normalize_cast
 *   dists_tr / d_1 row-wise distances         run_lemon.py:169,173,250-253  -> lemon_rowwise_dist
 *   faiss.IndexFlatIP/L2 .add/.search         run_lemon.py:167-176,235-236  -> lemon_knn_candidates (tcgen05)
 *                                                                              + lemon_rerank (fp32 exact re-rank)
 *                                                                              + lemon_knn_exact (fp32 fallback)
 *   per-sample loop                           run_lemon.py:238-307          -> lemon_score
 *   calc_scores_given_hparams_vectorized      lib/metrics/utils.py:47-82    -> lemon_score / lemon_combine_scores
 *
 * Conventions
 *   - every data pointer is a DEVICE pointer owned by the caller; the library allocates only
 *     the scratch held by its ctx;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), returns 0 on
 *     success or a negative lemon_status, never throws and never synchronises (except
 *     ctx create/destroy);
 *   - one ctx per device; a ctx is not thread-safe, different ctxs are independent;
 *   - DB row indices are int32 on the device (M < 2^31); the Python layer widens to int64
 *     where faiss does;
 *   - metric: 0 = inner product (descending, faiss IndexFlatIP), 1 = squared L2 (ascending,
 *     faiss IndexFlatL2).
 *   - "top list" = per query row `kp` entries sorted best-first under the documented total
 *     order (value best-first, then DB index ascending); missing entries are idx = -1,
 *     val = -inf (IP) / +inf (L2), as faiss pads.
 */
#ifndef LEMON_B200_H_
#define LEMON_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lemon_ctx lemon_ctx;

enum lemon_status {
  LEMON_OK = 0,
  LEMON_ERR_INVALID = -1,     /* bad argument (shape, alignment, unsupported d/k) */
  LEMON_ERR_CUDA = -2,        /* a CUDA runtime/driver call failed; see lemon_last_error */
  LEMON_ERR_UNSUPPORTED = -3  /* device is not sm_100 / feature not available */
};

enum lemon_metric { LEMON_METRIC_IP = 0, LEMON_METRIC_L2 = 1 };

#define LEMON_KPRIME 64          /* a row's candidate lists together hold its 64 best approximate values */
#define LEMON_LIST_CAP 1024      /* slots per candidate list */
#define LEMON_MAX_KP 64          /* largest k (+1 for self-exclusion) a top list can hold */
#define LEMON_MAX_D_TC 768       /* largest padded embedding dim served by the tensor-core path */
#define LEMON_MAX_D16_OPERAND (3 * LEMON_MAX_D_TC)   /* widest K1 operand: the split-precision second pass is 3 x d16 wide */

int lemon_version(void);
int lemon_ctx_create(int device, lemon_ctx** out);
int lemon_ctx_destroy(lemon_ctx* ctx);
const char* lemon_last_error(lemon_ctx* ctx);

/* Row-wise L2 normalisation + 16-bit operand copy (K0).
 *   in        [n, d]  fp32, rows in_stride floats apart (0 = d: contiguous).  A strided input lets the caller
 *                     normalise a column block of a wider matrix (the all-gathered [text | label] shard rows)
 *   out_f32   [n, d]  fp32, x / max(||x||, 1e-12) when do_normalize, else a copy (may alias `in`; may be NULL)
 *   out_f16   [n, d16] fp16, the (normalised) row rounded to nearest, zero padded to d16 (multiple of 64; may be NULL)
 *   row_stats [n, 4]  fp32 per row: {||x||, ||fp16(x)||, ||x - fp16(x)||, ||x||^2} of the OUTPUT row (may be NULL)
 *   stats_max [4]     fp32, maxima over the n rows of {||x||, ||fp16(x)||, ||x - fp16(x)||, | ||x||^2 - 1 |}
 *                     (zeroed by the call, on the stream; may be NULL)
 */
int lemon_normalize_cast(lemon_ctx* ctx, const float* in, float* out_f32, void* out_f16,
                         float* row_stats, float* stats_max, int64_t n, int d, int d16,
                         int64_t in_stride, int do_normalize, void* stream);

/* Split-precision operands for the second tensor-core pass over rows the first pass could not certify.
 *   x [n, d] fp32 (already normalised) = x_hi + x_lo + x_e with x_hi = fp16(x), x_lo = fp16(x - x_hi)
 *   out_f16 [n, 3*d16]: role 0 (queries) [hi | lo | hi], role 1 (database) [lo | hi | hi]; one lemon_knn_candidates
 *   call over these 3*d16-wide operands accumulates q_hi.b_lo + q_lo.b_hi + q_hi.b_hi in fp32 (small terms first),
 *   whose distance to the exact inner product is bounded by ||q_e|| max||b|| + ||q|| max||b_e|| + (accumulation of
 *   one d16-long product + the q_lo.b_lo term) -- ~1e-4 at d = 768 instead of the ~6.5e-4 of one fp16 word per operand.
 *   row_stats [n,4] = {||x||, ||x||, ||x_e||, ||x||^2}, stats_max [4] = their maxima (last: | ||x||^2 - 1 |), in the
 *   layout lemon_rerank reads (either may be NULL).
 */
int lemon_split_cast(lemon_ctx* ctx, const float* x, void* out_f16, float* row_stats, float* stats_max,
                     int64_t n, int d, int d16, int role, void* stream);

/* out[i] = 1 - <a_i, b_i> (metric IP / cosine)  or  sum (a_i - b_i)^2 (metric L2). run_lemon.py:169,173,250-253 */
int lemon_rowwise_dist(lemon_ctx* ctx, const float* a, const float* b, float* out,
                       int64_t n, int d, int metric, void* stream);

/* Tensor-core candidate search (K1): fp16 operands, fp32 TMEM accumulation, streaming top-k fused into the
 * epilogue; the nq x m similarity matrix never reaches HBM.
 *   q16  [nq, d16], db16 [m, d16]  fp16 row-major, d16 % 64 == 0, d16 <= LEMON_MAX_D16_OPERAND
 *   nseg  number of DB segments scanned independently (load balance for small nq); >= 1
 *   Output = LEMON_NLIST(nseg) = nseg * 2 candidate lists per query row (two epilogue warp groups per segment).
 *   nq_pad = nq rounded up to a multiple of 256 rows; the caller allocates all three arrays for nq_pad rows:
 *     cand_keys  [nq_pad, nseg*2, LEMON_LIST_CAP] uint64: key = (IEEE bit pattern of the approximate inner product, fp32) << 32
 *                | ~db_row; only the first cand_cnt entries of a list are valid, in no particular order;
 *                the array must be 8192-byte aligned (one list = 8 KB; LEMON_ERR_INVALID otherwise);
 *     cand_cnt   [nq_pad, nseg*2] int32 (0 .. LEMON_LIST_CAP);
 *     cand_theta [nq_pad, nseg*2] fp32: every DB column of that list's share of the scan that is NOT in the list
 *                has approximate inner product <= theta (-inf: the list holds everything it saw).
 *   keep (16 .. LEMON_KPRIME; 0 = LEMON_KPRIME): the two lists of a segment together hold at least `keep` columns at
 *   or above the larger of their two cand_theta values (when the segment is long enough to have a threshold at all),
 *   so together the lists of a row contain its `keep` best approximate inner products (ties: lower DB rows first).  A smaller `keep` means fewer appended keys; callers that need the top kp choose
 *   keep > kp with a margin for the fp16 rounding error (lemon_rerank certifies the result either way).
 *   cta_group: 1 or 2 (2 = cta_group::2 CTA pairs, 256 query rows per pair); 0 = library default.
 */
int lemon_knn_candidates(lemon_ctx* ctx, const void* q16, const void* db16, int64_t nq, int64_t m,
                         int d16, int nseg, int cta_group, int keep, uint64_t* cand_keys, int32_t* cand_cnt,
                         float* cand_theta, void* stream);

/* fp32 exact re-rank of the candidates + per-row certificate (K2a).
 *   q [nq, d], db [m, d] fp32;  cand_* from lemon_knn_candidates with nlist = nseg*2 lists per row (nlist <= 64, i.e. at
 *   most 32 segments).
 *   The kernel first selects the row's 64 best approximate candidates over the union of its lists, drops those
 *   that provably cannot be in the exact top-kp, gathers the rest and evaluates them exactly in fp32.
 *   q_row_stats [nq,4], db_stats_max [4]: outputs of lemon_normalize_cast for the query rows and the DB.
 *   They give the rigorous per-row bound on |fp16 tensor-core inner product - exact|:
 *     eps_row = ||q - q16|| * max||b16|| + ||q|| * max||b - b16|| + acc_eps * max(||q||,||q16||) * max(||b||,||b16||)
 *   (Cauchy-Schwarz on the two rounding-error vectors; acc_eps is the RELATIVE bound on the fp32 accumulation of
 *   the tensor-core product plus that of the fp32 re-evaluation: d16 * 2^-23 + (d/32 + 6) * 2^-24, i.e. the
 *   gamma_n * sum|q_i b_i| bound with truncating adds).  NULL = eps 0.
 *   Metric L2 ranks by -||q-b||^2 = 2<q,b> - ||q||^2 - ||b||^2; the bound then uses ||q||^2 from
 *   q_row_stats and min||b||^2 >= 1 - db_stats_max[3].
 *   top_val / top_idx [nq, kp]: exact top list.  A row is certified when its kp-th exact value beats every
 *   non-candidate's bound (max over lists of cand_theta, + eps_row); otherwise its row id is appended to
 *   uncert_rows[0 .. *n_uncert) (*n_uncert is zeroed by the call, on the stream; room for nq ids).
 *   out_rows [nq] int32 or NULL: query row r writes row out_rows[r] of top_val / top_idx and reports that id when
 *   uncertified (second-pass calls on a gathered subset of the rows).
 */
int lemon_rerank(lemon_ctx* ctx, const float* q, const float* db, const uint64_t* cand_keys,
                 const int32_t* cand_cnt, const float* cand_theta, const float* q_row_stats,
                 const float* db_stats_max, float acc_eps, int64_t nq, int64_t m, int d, int nlist, int kp,
                 int metric, const int32_t* out_rows, float* top_val, int32_t* top_idx, int32_t* uncert_rows,
                 int32_t* n_uncert, void* stream);

/* fp32 brute-force exact kNN on CUDA cores (GPU fallback for uncertified rows, and the
 * general path for shapes the tensor-core kernel does not take).
 *   rows: NULL = all nq rows; else a device list of row ids with its length in *n_rows (device).
 *   max_rows: upper bound of *n_rows (sizes the grid; nq when rows == NULL).
 *   Writes the top list of each processed row into top_val/top_idx [nq, kp].
 */
int lemon_knn_exact(lemon_ctx* ctx, const float* q, const float* db, const int32_t* rows,
                    const int32_t* n_rows, int64_t max_rows, int64_t nq, int64_t m, int d, int kp,
                    int metric, float* top_val, int32_t* top_idx, void* stream);

/* Per-sample records + score (K2b): restates run_lemon.py:250-307 and utils.py:63-77 for given top lists.
 *   xq,yq [nq,d]  image/text query rows;  xdb,ydb [m,d] image/text DB rows;  dists_tr [m]
 *   topn_* / topm_* [nq, kp]: image-kNN / text-kNN top lists, kp = k + 1 when query_in_db != NULL else k
 *   query_in_db [nq] int64 or NULL: train-split rule (>=0: drop rank 0, -1: drop the last)
 *   label_q [nq], label_db [m] int32 or NULL: discrete text metric (--use_discrete_for_text)
 *   class_emb [n_class, d], noisy_label [nq] int32 or NULL: --normalize_d1 (run_lemon.py:244-248):
 *     d_1 = softmax_c(1 - <x_i, T_c>)[noisy_label_i]  (cosine)  /  softmax_c(||x_i - T_c||^2)[noisy_label_i]  (euclidean)
 *   hp[6] = {beta, gamma, tau_1_n, tau_2_n, tau_1_m, tau_2_m}: HOST pointer read at call time, or NULL
 *   (then sn/sm/score are not written)
 *   outputs (any may be NULL): d1 [nq]; Dn,dists_n,dists_tr_n,Dm,dists_m,dists_tr_m [nq,k] fp32;
 *   In, Im [nq,k] int64 (index_bits = 64, what faiss returns) or int32 (index_bits = 32: half the bytes for
 *   callers that copy the records to the host);  sn, sm, score [nq] float64.
 *   sides: 3 = everything in one call; 1 = only the image-neighbour side (D_n, dists_n, dists_tr_n, I_n, s_n), 2 = only
 *   the text-neighbour side plus d_1 and the score, which then reads s_n written by an earlier sides = 1 call on the
 *   same stream (lets a caller ship the image-side records to the host while the text-side search is still running).
 */
int lemon_score(lemon_ctx* ctx, const float* xq, const float* yq, const float* xdb, const float* ydb,
                const float* dists_tr, const float* topn_val, const int32_t* topn_idx,
                const float* topm_val, const int32_t* topm_idx, const int64_t* query_in_db,
                const int32_t* label_q, const int32_t* label_db, const float* class_emb,
                const int32_t* noisy_label, int n_class, int64_t nq, int64_t m, int d, int k,
                int kp, int metric, const double* hp, float* d1, float* Dn, float* dists_n,
                float* dists_tr_n, float* Dm, float* dists_m, float* dists_tr_m, void* In,
                void* Im, int index_bits, int sides, double* sn, double* sm, double* score, void* stream);

/* Score combination only (lib/metrics/utils.py:63-77) on stacked [n,k] fp32 columns. */
int lemon_combine_scores(lemon_ctx* ctx, const float* Dn, const float* dists_tr_n, const float* dists_n,
                         const float* Dm, const float* dists_tr_m, const float* dists_m,
                         const double* d1, int64_t n, int k, const double* hp /* HOST [6] */, double* sn,
                         double* sm, double* score, void* stream);

/* Exact-duplicate DB rows are searched once (classification datasets: C distinct text embeddings; caption noise
 * duplicates captions: run_lemon.py:117-119,140-143, lib/datasets/noise_captioning.py:44-53).
 * lemon_dedup_build groups the bit-identical rows of x [n,d] entirely on the device (hash -> radix sort -> run
 * scan -> bit-wise verification -> renumbering), without a host round trip:
 *   workspace   lemon_dedup_workspace_bytes(n) bytes, 256 B aligned (caller-owned scratch)
 *   counters[2] {n_unique, collision}: collision != 0 means two different rows shared a 63-bit hash; the
 *               outputs are then unusable and the caller must not de-duplicate
 *   rep_rows [n]   first n_unique entries: lowest row index of each group, ascending (group u = u-th unique row)
 *   offsets [n+1]  first n_unique+1 entries: group u owns members[offsets[u] .. offsets[u+1])
 *   members [n]    row indices grouped by unique row, ascending inside a group
 * lemon_gather_rows: dst[r] = src[idx[r]] for r < min(*n_idx, max_idx) (n_idx device pointer or NULL = max_idx);
 *   rows of row_bytes (multiple of 16) bytes.  Builds the unique-row operands from rep_rows.
 * lemon_expand_groups: turns top lists over the unique rows (uval/uidx [nq,kp]) into top lists over the original
 *   rows; entries keep the group's value; members of consecutive unique rows with EQUAL values are merged by
 *   ascending row index (the documented total order).
 * lemon_hash_rows: the 63-bit row hash alone (diagnostics / tests).
 */
/* lemon_dedup_count: counters[0] = number of rows whose 63-bit hash equals an earlier row's (n - counters[0] =
 * number of unique rows, barring hash collisions): two kernels, enough to decide whether lemon_dedup_build is
 * worth running.  workspace: lemon_dedup_count_workspace_bytes(n) bytes, 256 B aligned. */
int64_t lemon_dedup_count_workspace_bytes(int64_t n);
int lemon_dedup_count(lemon_ctx* ctx, const float* x, int64_t n, int d, void* workspace, int32_t* counters,
                      void* stream);
int64_t lemon_dedup_workspace_bytes(int64_t n);
int lemon_dedup_build(lemon_ctx* ctx, const float* x, int64_t n, int d, void* workspace, int32_t* rep_rows,
                      int32_t* members, int64_t* offsets, int32_t* counters, void* stream);
int lemon_gather_rows(lemon_ctx* ctx, const void* src, const int32_t* idx, const int32_t* n_idx, int64_t max_idx,
                      int64_t row_bytes, void* dst, void* stream);
int lemon_hash_rows(lemon_ctx* ctx, const float* x, int64_t n, int d, int64_t* out, void* stream);
int lemon_expand_groups(lemon_ctx* ctx, const float* uval, const int32_t* uidx, const int64_t* offsets,
                        const int32_t* members, int64_t nq, int kp, int metric, float* top_val,
                        int32_t* top_idx, void* stream);

/* CC3M filtering consumer of the scores (train_clip_from_scratch.py:110-113: sort ascending by score, keep the first
 * cc3m_filtering_n rows).  out_idx [n_keep] int64 = row ids of the n_keep LOWEST scores in ascending score order
 * (ties: lower row id first; pandas' default quicksort leaves tie order unspecified), out_score [n_keep] their
 * scores (may be NULL).  workspace: lemon_keep_lowest_workspace_bytes(n) bytes, 256 B aligned.  NaN scores sort last.
 */
int64_t lemon_keep_lowest_workspace_bytes(int64_t n);
int lemon_keep_lowest(lemon_ctx* ctx, const double* score, int64_t n, int64_t n_keep, void* workspace,
                      int64_t* out_idx, double* out_score, void* stream);

/* Hyper-parameter grid stage (lib/metrics/utils.py:117-121,167-186,286-296): for each of n_points grid points
 *   score_i = d1[i] + beta[g] * sn[tidx[g], i] + gamma[g] * sm[tidx[g], i]      (all float64, utils.py:77)
 * and Brent's bounded minimiser (scipy.optimize.fminbound, xtol = xatol, maxfun) of t -> -F1(y, score >= t) on
 * [min score, max score]; out_f1[g] = F1 at the returned threshold out_thr[g] (optimize_f1_efficient).
 * sn/sm [n_tau, n] are the neighbour terms of each (tau_1, tau_2) combination (lemon_combine_scores / lemon_score);
 * with sn == sm == NULL the scores are d1 itself (single call of optimize_f1_efficient).  y [n] uint8 labels.
 * scratch: scratch_rows * n float64; one thread block per grid point, at most scratch_rows blocks resident.
 */
int lemon_f1_grid(lemon_ctx* ctx, const double* d1, const double* sn, const double* sm, const uint8_t* y,
                  int64_t n, const double* beta, const double* gamma, const int32_t* tidx, int64_t n_points,
                  double xatol, int maxfun, double* out_f1, double* out_thr, double* scratch,
                  int64_t scratch_rows, void* stream);

/* Discrepancy / diversity baseline scores (lib/baselines/discrepancy_baseline.py:213-230) from kNN lists.
 *   emb [m,d]: the modality matrix the score reads (text for *_y, image for *_x); qemb [nq,d]: the query's embedding
 *   in that modality (mode 0 only); nn [nq,kk]: text-kNN list of every query (kk = k, or k+1 for train queries);
 *   cache [m,kc]: text-kNN list of every DB row, kc = k+1, the row itself is skipped (mode 0 only).
 *   mode 0 (dis_x / dis_y): out = sum(1 - <emb[l], qemb>) / L over the second-order neighbours l;
 *   mode 1 (div_x / div_y): out = sum_ab (1 - <emb[a], emb[b]>) / k^2 over the first-order neighbours.
 */
int lemon_discrepancy(lemon_ctx* ctx, const float* emb, const float* qemb, const int32_t* nn,
                      const int32_t* cache, int64_t nq, int64_t m, int d, int kk, int kc, int k, int mode,
                      float* out, void* stream);

/* Number of kernels this library has launched through `ctx` since creation (bench "gpu_launches"). */
int64_t lemon_launch_count(lemon_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* LEMON_B200_H_ */
