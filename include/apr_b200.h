/*
 * apr_b200.h -- C ABI of the B200-native APR / BPR-MF hot path (libapr_b200.so).
 *
 * The reference (feay1234/Adversarial-Collaborative-Filtering) is pure Python on TensorFlow-1.x and has
 * NO native boundary; its seam for this path is `sess.run(fetches, feed_dict)` on attributes of class MF
 * (APR.py:85-202) called from utils.py:106-267.  Each entry point below replaces one such fetch group and
 * cites it.  INTEGRATION.md shows the ctypes stub a reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer borrowed from the caller (torch tensors on the Python side) unless the
 *     parameter name ends in `_host`; nothing is retained after the call returns;
 *   - tables are row-major float32 [rows, d], d % 4 == 0, 4 <= d <= 512, base pointers 16-byte aligned;
 *   - ids are int32; CSR row pointers are int64;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, no host synchronisation unless
 *     stated; the only allocations are the caller-provided workspaces;
 *   - return value: 0 = ok, otherwise an APR_E_* code (apr_status_string() describes it).  There is no CPU
 *     fallback: without a CUDA device every compute entry point returns APR_E_CUDA.
 */
#ifndef APR_B200_H_
#define APR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define APR_ABI_VERSION 2   /* round 2: apr_train_unique_counts takes the workspace size, new entry points */

enum {
  APR_OK = 0,
  APR_E_ARG = 1,       /* bad shape / null pointer / unsupported d */
  APR_E_ALIGN = 2,     /* pointer not 16-byte aligned */
  APR_E_WORKSPACE = 3, /* workspace too small */
  APR_E_CUDA = 4,      /* CUDA runtime error (see apr_last_cuda_error) */
  APR_E_UNSUPPORTED = 5
};

typedef void* apr_stream_t;

int apr_abi_version(void);
const char* apr_status_string(int status);
const char* apr_last_cuda_error(void);
/* sm count / compute capability of the current device. */
int apr_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor);

/* Per-device host state (side streams, fork/join events, the cache of executable CUDA graphs of the training step,
 * the cross-rank barrier epoch, the evaluation timing events) lives in one context per CUDA device, created on the
 * first call made while that device is current; entry points that use it are serialised per device (the reference
 * drives its session from one host thread).  apr_context_create builds it ahead of time (so that no
 * stream / event creation lands inside a timed call), apr_context_destroy releases it (after the device is idle).
 * apr_context_stats (current device): out4_host = {graph instantiations, in-place graph updates, graph launches,
 * cached executable graphs} since the context was created -- bench.py reports the first two across its timed region. */
int apr_context_create(int32_t device);
int apr_context_destroy(int32_t device);
int apr_context_stats(int64_t* out4_host);

/* ---- A1 / K12: MF._create_variables, APR.py:105-119 (tf.truncated_normal(mean 0, stddev)) and the
 *      `--adv random` noise of APR.py:172-177.  Philox4x32-10, counter (e_lo, e_hi, attempt, table_id),
 *      key (seed, stream_tag); see oracle/apr_oracle.py:truncated_normal. */
int apr_init_truncated_normal(float* W, int64_t rows, int32_t d, float stddev, uint32_t seed, uint32_t table_id,
                              uint32_t stream_tag, apr_stream_t stream);
int apr_fill_f32(float* x, int64_t n, float value, apr_stream_t stream);

/* ---- A8 / K6: shuffle + _get_train_batch, APR.py:39-81.  Epoch permutation (Feistel bijection keyed by
 *      (seed, epoch)) of the n_pairs (u,i) pairs, tail batch dropped, and `dns` uniform negatives per positive in
 *      [0,num_items) rejected against the sorted CSR of trainList (csr_rows rows).  Outputs: out_u/out_i
 *      [S*B], out_udns/out_j [S*B*dns] with S = n_pairs / batch.  *err_flag (device int32) is set non-zero if a
 *      draw did not terminate within 65536 attempts.  Bit-exact with oracle.sample_epoch. */
int apr_sample_epoch(const int32_t* pairs_u, const int32_t* pairs_i, int64_t n_pairs, int32_t batch, int32_t num_items,
                     const int64_t* csr_ptr, const int32_t* csr_idx, int32_t csr_rows, uint32_t seed, uint32_t epoch,
                     int32_t dns, int32_t* out_u, int32_t* out_i, int32_t* out_udns, int32_t* out_j, int32_t* err_flag,
                     apr_stream_t stream);

/* The same epoch, partitioned over the ranks of a data-parallel run (sampler row): this call draws only the
 * triples [batch_lo, batch_lo + batch_local) of every batch -- outputs out_u/out_i [S*batch_local], out_udns/out_j
 * [S*batch_local*dns].  Every random number is keyed by the triple's index in the WHOLE epoch, so the shards are
 * bit-identical slices of apr_sample_epoch's output and no collective is needed (rank r of G passes
 * batch_lo = r * batch / G, batch_local = batch / G).
 * legacy_fork_workers > 0 reproduces, statistically, the reference's fork-duplicated negative streams (SURVEY B.3:
 * Pool(cpu_count()) forked after the shuffle, every worker starting from the parent's RNG state): the W chunks of one
 * pool.map round share one counter range (oracle.fork_counter).  0 = independent draws (the default everywhere). */
int apr_sample_epoch_shard(const int32_t* pairs_u, const int32_t* pairs_i, int64_t n_pairs, int32_t batch,
                           int32_t num_items, const int64_t* csr_ptr, const int32_t* csr_idx, int32_t csr_rows,
                           uint32_t seed, uint32_t epoch, int32_t dns, int32_t batch_lo, int32_t batch_local,
                           int32_t legacy_fork_workers, int32_t* out_u, int32_t* out_i, int32_t* out_udns, int32_t* out_j,
                           int32_t* err_flag, apr_stream_t stream);

/* ---- A7 (dns > 1 branch), utils.py:121-139: for each positive keep the best-scored of its dns negatives
 *      (first maximum wins).  u_dns/j_dns are [n_pos*dns]; out_j is [n_pos]. */
int apr_select_dns(const float* P, const float* Q, int32_t d, const int32_t* u_dns, const int32_t* j_dns, int64_t n_pos,
                   int32_t dns, int32_t* out_j, apr_stream_t stream);

/* ---- A2-A7 / K1-K5: training_batch, utils.py:106-119: for every batch s of n_steps,
 *        if adver: sess.run([update_P, update_Q])   (APR.py:180-191)   Delta = eps * G / ||G|| per touched row
 *        sess.run(optimizer)                        (APR.py:143-165,193-195)  Adagrad on
 *                                                   L + reg_adv * L_adv + reg-terms, duplicates summed.
 *      u,i,j are [n_steps * batch].  accP/accQ are the Adagrad accumulators (same shape as P/Q, init 0.1).
 *      Delta never exists as a table: it lives in a per-step workspace of touched rows.
 *      The workspace must hold apr_train_workspace_bytes(n_steps, batch, d) bytes and must have been zeroed once
 *      with apr_train_workspace_init before first use (the kernels restore the zero invariant themselves).  Its internal
 *      layout is a function of (workspace_bytes, batch, d) only -- capacity = the largest step count that fits -- so a
 *      workspace sized for N steps serves calls of any n_steps <= N with the same array addresses (which is what lets
 *      the executable CUDA graphs of the step be replayed for calls of any length); pass the same workspace_bytes to
 *      every entry point that takes the workspace.
 *      stats (nullable) receives per step {sum softplus(-r) of the PLAIN forward, count(x > 0)} as float[2].
 *      mode: 0 = one kernel launch per phase (3 per APR step, 2 per BPR step);
 *            1 = one persistent cooperative kernel for all n_steps (grid barriers between phases);
 *            2 = the same persistent kernel launched as ONE thread-block cluster of <= 16 CTAs (cluster barriers between
 *                phases): for small batches whose steps take a few microseconds (the reference's default B = 512). */
int64_t apr_train_workspace_bytes(int32_t n_steps, int32_t batch, int32_t d);
int apr_train_workspace_init(void* workspace, int64_t workspace_bytes, apr_stream_t stream);
int apr_train_steps(float* P, float* Q, float* accP, float* accQ, int64_t rows_p, int64_t rows_q, int32_t d,
                    const int32_t* u, const int32_t* i, const int32_t* j, int32_t n_steps, int32_t batch, float lr,
                    float reg, float reg_adv, float eps, int32_t adver, int32_t mode, void* workspace,
                    int64_t workspace_bytes, float* stats, apr_stream_t stream);
/* adver (all training entry points): 0 = BPR step; 1 = APR, Delta from the batch gradient (--adv grad); 3 = the
 * reference's dns > 1 branch on an adversarial graph (utils.py:121-139: the optimizer of the adversarial loss with
 * Delta == 0, i.e. data term x (1 + reg_adv), regulariser counted twice); 2 is rejected here -- `--adv random` is
 * apr_train_steps_random:
 *
 * ---- A5' `--adv random`, APR.py:170-177 (shape-consistent form evaluation_adv.py:182-189): before every step
 *      Delta_P = eps * l2_normalize(truncated_normal([rows, d], 0, 0.01)), Delta_Q likewise, fresh for ALL rows.  Only
 *      touched rows matter, so each row's noise is generated where the row is used (Philox: element = row * d + col,
 *      table id = 2 * global_step + {0: P, 1: Q}, key = (noise_seed, stream ADV); oracle.random_delta) -- never a
 *      [rows, d] table.  Step s of the call uses global step noise_first_step + s. */
int apr_train_steps_random(float* P, float* Q, float* accP, float* accQ, int64_t rows_p, int64_t rows_q, int32_t d,
                           const int32_t* u, const int32_t* i, const int32_t* j, int32_t n_steps, int32_t batch, float lr,
                           float reg, float reg_adv, float eps, uint32_t noise_seed, uint32_t noise_first_step,
                           void* workspace, int64_t workspace_bytes, float* stats, apr_stream_t stream);
/* Synchronises and returns the workspace's sticky status word: bit 0 = an id outside its table reached the index
 * preparation since apr_train_workspace_init (that triple trained row 0 instead of faulting; the reference's
 * embedding_lookup raises InvalidArgument).  apr_b200.APR.Session raises on it at the end of every training_batch. */
int apr_train_status(const void* workspace, int32_t* flags_host, apr_stream_t stream);
/* The two halves of apr_train_steps, exposed so the index preparation (hash de-duplication of the rows each batch
 * touches) can be timed and profiled separately from the embedding kernels. */
int apr_train_prepare(const int32_t* u, const int32_t* i, const int32_t* j, int32_t n_steps, int32_t batch, int32_t d,
                      int64_t rows_p, int64_t rows_q, void* workspace, int64_t workspace_bytes, apr_stream_t stream);
int apr_train_run(float* P, float* Q, float* accP, float* accQ, int64_t rows_p, int64_t rows_q, int32_t d,
                  const int32_t* u, const int32_t* i, const int32_t* j, int32_t n_steps, int32_t batch, float lr,
                  float reg, float reg_adv, float eps, int32_t adver, int32_t mode, void* workspace,
                  int64_t workspace_bytes, float* stats, apr_stream_t stream);
/* Per-step unique-row counts produced by apr_train_prepare (for the algorithmic-bytes model of the roofline):
 * copies n_steps int32 pairs {unique users, unique items} to the HOST buffer; synchronises the stream.
 * Returns APR_E_ARG if any id seen by apr_train_prepare since apr_train_workspace_init was outside its table
 * (such ids are clamped to row 0 on the device instead of faulting). */
int apr_train_unique_counts(const void* workspace, int64_t workspace_bytes, int32_t n_steps, int32_t batch, int32_t d,
                            int32_t* counts_host, apr_stream_t stream);

/* ---- Row-sharded training over peer-mapped tables (multi-GPU drop-in for the same sess.run pair).
 *      Row r of a table lives on rank r % nranks at local row r / nranks (nranks in {1,2,4,8}); Pb/Qb/accPb/accQb are
 *      HOST arrays of nranks device pointers (each rank's shard base, peer-mapped over NVLink), GQb/HQb the shards of
 *      the shared-item workspace (>= batch/nranks + 1 rows each, zeroed once).  `batch` is the GLOBAL batch; every rank
 *      holds identical index arrays in `workspace` (apr_train_prepare_range on one rank + a broadcast of the regions
 *      apr_train_layout reports) and processes every nranks-th segment.  One call launches ONE stage of ONE step:
 *      stage 0,1,2 = general path, 3 = fast kernel, 4 = pair work units; the caller puts a cross-rank barrier after
 *      stages 0, 1 and after {2,3,4}.  apr_b200/distributed.py is that caller. */
/* apr_train_layout: byte offsets of the index arrays inside a workspace of EXACTLY apr_train_workspace_bytes(n_steps,
 * batch, d) bytes (the layout follows the workspace size, see apr_train_steps). */
int apr_train_layout(int32_t n_steps, int32_t batch, int32_t d, int64_t* out13);
int apr_train_prepare_range(const int32_t* u, const int32_t* i, const int32_t* j, int32_t n_steps, int32_t batch,
                            int32_t d, int64_t rows_p, int64_t rows_q, void* workspace, int64_t workspace_bytes,
                            int32_t first_step, int32_t count, int32_t clear_counters, apr_stream_t stream);
int apr_train_stage_sharded(float* const* Pb, float* const* Qb, float* const* accPb, float* const* accQb,
                            float* const* GQb, float* const* HQb, int32_t nranks, int32_t rank, int32_t d,
                            int32_t n_steps, int32_t batch, float lr, float reg, float reg_adv, float eps, int32_t adver,
                            void* workspace, int64_t workspace_bytes, float* stats, int32_t step, int32_t stage,
                            apr_stream_t stream);

/* The whole per-step sequence of the sharded path launched from the library (no host work between steps): fast kernel
 * on the internal second stream, general stages on `stream`, and a cross-rank barrier (peer-mapped signal words,
 * bounded spin) after the plain stage, after the adversarial stage and at the end of every step.  sig_host = HOST array
 * of nranks device pointers to >= nranks int32 each (every rank's signal words, zeroed once); err = this rank's device
 * int32, set non-zero if a barrier timed out.  Every rank must issue the same sequence of calls. */
int apr_train_steps_sharded(float* const* Pb, float* const* Qb, float* const* accPb, float* const* accQb,
                            float* const* GQb, float* const* HQb, int32_t* const* sig_host, int32_t nranks, int32_t rank,
                            int32_t d, int32_t n_steps, int32_t batch, float lr, float reg, float reg_adv, float eps,
                            int32_t adver, void* workspace, int64_t workspace_bytes, float* stats, int32_t first_step,
                            int32_t count, int32_t* err, apr_stream_t stream);

/* ---- A9 / K7: training_loss_acc, utils.py:159-175 (output_adv = 0): per batch s,
 *      out[2s] = sum_b softplus(-clip(x_b)), out[2s+1] = count(x_b > 0), as float64. */
int apr_loss_acc(const float* P, const float* Q, int32_t d, const int32_t* u, const int32_t* i, const int32_t* j,
                 int32_t n_steps, int32_t batch, double* out, apr_stream_t stream);

/* ---- A2 / Recommender.rank (MF.py:37-39, utils.py:246-251): scores[k] = <P[users[k]], Q[items[k]]> with the
 *      pinned order acc = fmaf(p[t], q[t], acc), t ascending (oracle.score_pairs). */
int apr_score_pairs(const float* P, const float* Q, int32_t d, const int32_t* users, const int32_t* items, int64_t n,
                    float* scores, apr_stream_t stream);

/* ---- N4: the all-item scorer `all_rating = u . Q^T` behind `rank` of IRGAN.py:36-39 / APL.py:205-211 /
 *      SASRec.py:424-436: scores[k, c - item_lo] = <P[users[k]], Q[c]> for c in [item_lo, item_hi), pinned fma order.
 *      Dense [n_users, item_hi - item_lo] output: for a few users at a time (n_users <= 65535); full evaluation uses
 *      apr_eval_fullrank*, which never writes the score matrix. */
int apr_score_all_items(const float* P, const float* Q, int32_t d, const int32_t* users, int32_t n_users, int32_t item_lo,
                        int32_t item_hi, float* scores, apr_stream_t stream);

/* ---- A10 / K8: _eval_by_user on explicit candidate lists, utils.py:244-254 and evaluation.py:114-128.
 *      User k has candidates cand_idx[cand_ptr[k] .. cand_ptr[k+1]) with the held-out item LAST;
 *      position[k] = #(score(neg) >= score(held-out)).  scores (nullable) receives every candidate score. */
int apr_eval_candidates(const float* P, const float* Q, int32_t d, const int32_t* users, const int64_t* cand_ptr,
                        const int32_t* cand_idx, int32_t n_users, int32_t* position, float* scores,
                        apr_stream_t stream);

/* ---- A10 / K9: _eval_by_user with eval_mode == "all", utils.py:210-215,244-261.  For each listed user:
 *      candidates = [item_lo, item_hi) minus excl(u) (sorted CSR; = trainList[u] united with the held-out item),
 *      position[k] += #(candidates c : score(u,c) >= score(u, test_item[k])), scores in the pinned fma order;
 *      topk (nullable, k_top > 0): the k_top best candidates by (score desc, item id asc) of this item range,
 *      topk_ids/topk_scores [n_users, k_top], padded with id -1 / -inf.  position must be zeroed by the caller
 *      (item-sharded callers sum the shards' counts).  `exact` = 1 forces the fp32 CUDA-core kernel; with 0 the call
 *      goes through apr_eval_fullrank_tc_topk (tcgen05 bf16x3 filter + exact re-scoring, same results) whenever d is
 *      supported there and the workspace is 1024-byte aligned and holds apr_eval_tc_topk_workspace_bytes(n_users,
 *      item_hi - item_lo, d, k_top) bytes; otherwise the fp32 kernel runs.
 *      The workspace must hold at least apr_eval_workspace_bytes(n_users, k_top, d) bytes. */
int64_t apr_eval_workspace_bytes(int32_t n_users, int32_t k_top, int32_t d);
int apr_eval_fullrank(const float* P, const float* Q, int32_t d, const int32_t* users, const int32_t* test_item,
                      int32_t n_users, int32_t item_lo, int32_t item_hi, const int64_t* excl_ptr,
                      const int32_t* excl_idx, int32_t k_top, int32_t* position, int32_t* topk_ids, float* topk_scores,
                      int32_t exact, void* workspace, int64_t workspace_bytes, apr_stream_t stream);

/* ---- A10 / K9 on the tensor cores: the same positions as apr_eval_fullrank, computed by a tcgen05 bf16x3 GEMM with an
 *      error-bounded count and exact fp32 re-scoring of the ambiguous candidates (csrc/eval_tc.cu).
 *      d % 8 == 0, d <= 256.  workspace: apr_eval_tc_workspace_bytes(n_users, item_hi - item_lo, d) bytes, 1024-byte
 *      aligned.  *err_flag (device int32) is set if a pipeline wait timed out. */
int64_t apr_eval_tc_workspace_bytes(int32_t n_users, int32_t n_items, int32_t d);
int apr_eval_fullrank_tc(const float* P, const float* Q, int32_t d, const int32_t* users, const int32_t* test_item,
                         int32_t n_users, int32_t item_lo, int32_t item_hi, const int64_t* excl_ptr,
                         const int32_t* excl_idx, int32_t* position, void* workspace, int64_t workspace_bytes,
                         int32_t* err_flag, apr_stream_t stream);

/* ---- K9 with the top-k fused onto the same tensor-core pipeline (evaluation.py:54-76: heapq.nlargest order, ties to the
 *      smaller item id; utils.py:244-261).  Positions as above, plus -- k_top in [1, 128] -- topk_ids / topk_scores
 *      [n_users, k_top]: the k_top best NON-EXCLUDED items of [item_lo, item_hi) by (exact fp32 chain score desc, id asc),
 *      padded with -1 / -inf; bit-identical to apr_eval_fullrank(exact = 1).  The counting pass also tracks, per user,
 *      32 x splits maxima of tensor-core scores of distinct items; the (k_top + |excl(u)|)-th largest of them minus the
 *      error bound is a lower bound tau_u of the user's k-th best exact score; a second pass over the same operand
 *      images lists every item with s_tc >= tau_u - E, the list is re-scored exactly, excluded items dropped, and the
 *      k_top best selected per user.  Users the filter cannot serve (fewer maxima than k_top + |excl(u)|, overflowing
 *      lists, non-finite rows) are ranked by an exact per-user kernel inside the same call.
 *      q_version: 0 = build the item-operand image (bf16 hi/lo split of Q[item_lo:item_hi)) on every call; non-zero = the
 *      caller's version tag of the CONTENT of Q -- the image in `workspace` is reused by following calls with the same
 *      (Q, range, d, workspace, q_version), e.g. over the next user tiles, and rebuilt when the tag changes.
 *      spos_in (nullable): the held-out items' scores [n_users], computed elsewhere -- the item-sharded caller, whose
 *      rank holds only Q[item_lo:item_hi) (pass Q = shard base - item_lo * d floats: only rows of the range are
 *      dereferenced) and gets score(u, test_item[u]) from the rank that owns that row; test_item may then be NULL.
 *      workspace: apr_eval_tc_topk_workspace_bytes(n_users, n_items, d, k_top) bytes, 1024-byte aligned.
 *      apr_eval_tc_ambiguous (synchronises): count_host[0] = (user, item) pairs the counting pass re-scored exactly,
 *      [1] = capacity of that list (0 = a segment overflowed: positions invalid, use exact = 1), [2] = top-k candidates
 *      re-scored, [3] = users ranked by the exact per-user kernel. */
int64_t apr_eval_tc_topk_workspace_bytes(int32_t n_users, int32_t n_items, int32_t d, int32_t k_top);
int apr_eval_fullrank_tc_topk(const float* P, const float* Q, int32_t d, const int32_t* users, const int32_t* test_item,
                              int32_t n_users, int32_t item_lo, int32_t item_hi, const int64_t* excl_ptr,
                              const int32_t* excl_idx, int32_t k_top, int32_t* position, int32_t* topk_ids,
                              float* topk_scores, uint64_t q_version, const float* spos_in, void* workspace,
                              int64_t workspace_bytes, int32_t* err_flag, apr_stream_t stream);
int apr_eval_tc_ambiguous(const void* workspace, int32_t n_users, int32_t n_items, int32_t d, int32_t k_top,
                          int32_t* count4_host, apr_stream_t stream);

/* ---- K10: merge of per-shard top-k lists (item-sharded evaluation): in_ids / in_scores [n_users, m] with
 *      m = shards * k <= 1024 entries per user (id < 0 = padding) -> the k best by (score desc, id asc). */
int apr_topk_merge(const int32_t* in_ids, const float* in_scores, int32_t n_users, int32_t m, int32_t k, int32_t* out_ids,
                   float* out_scores, apr_stream_t stream);

/* Measurement hook (bench.py): enable != 0 makes every following apr_eval_fullrank_tc record CUDA events around its GEMM +
 * counting kernel on the caller's stream; *gemm_ms_out (may be NULL) receives the duration of the most recent timed call
 * (-1 if none).  Not thread-safe; leave disabled in production. */
int apr_eval_tc_timing(int32_t enable, float* gemm_ms_out);

/* ---- N1: device-side data loader (csrc/loader.cu).  `text` = the bytes of a `*.rating` TSV file
 *      ("uid \t iid \t rating \t timestamp", Dataset.py:278-304, App. C) or of a He-format `.test.negative` file
 *      ("(u,i) \t n1 ... \t n99", Dataset.py:161-172) in DEVICE memory.  All functions use a caller workspace of
 *      apr_loader_workspace_bytes(n_bytes, max_lines) bytes.
 *      apr_tsv_count_lines  non-empty lines -> *n_lines_host (synchronises); leaves the per-block line bases in the
 *                           workspace for apr_tsv_parse (same text, same workspace).
 *      apr_tsv_parse        mode 0: columns uid / iid / rating (1.0 when the column is missing) [max_lines];
 *                           mode 1: ids per negatives line -> tok_count; mode 2: the ids -> neg_idx at neg_ptr[line].
 *                           *err_flag: bit 0 malformed field, bit 1 more lines than max_lines.
 *      apr_loader_train_rows  the reference's trainList row of every line (Dataset.py:306-325): quirk != 0 reproduces
 *                           its cursor that advances by at most one user per line as two prefix scans, valid
 *                           for a uid-sorted file (*unsorted_flag is set otherwise: use the host cursor); quirk == 0: row = uid.
 *      apr_loader_csr       sorted, de-duplicated CSR of (row, item): trainList membership structure (APR.py:77,
 *                           utils.py:211) or the evaluation exclusion lists.
 *      apr_loader_unique_pairs  keys of the dok trainMatrix in insertion order (rating > 0, duplicates collapsed onto the
 *                           first occurrence): the (u, i) list `sampling` enumerates (APR.py:30-36). */
int64_t apr_loader_workspace_bytes(int64_t n_bytes, int64_t max_lines);
int apr_tsv_count_lines(const char* text, int64_t n_bytes, void* workspace, int64_t workspace_bytes, int64_t* n_lines_host,
                        apr_stream_t stream);
int apr_tsv_parse(const char* text, int64_t n_bytes, const void* workspace, int64_t max_lines, int32_t mode, int32_t* out_u,
                  int32_t* out_i, float* out_r, int64_t* tok_count, const int64_t* neg_ptr, int32_t* neg_idx,
                  int32_t* err_flag, apr_stream_t stream);
int apr_loader_train_rows(const int32_t* u, int64_t n, int32_t quirk, int32_t* row, int32_t* unsorted_flag, void* workspace,
                          int64_t workspace_bytes, apr_stream_t stream);
int apr_loader_csr(const int32_t* row, const int32_t* item, int64_t n, int64_t rows, int64_t* ptr, int32_t* idx,
                   int64_t* n_unique_dev, void* workspace, int64_t workspace_bytes, apr_stream_t stream);
int apr_loader_unique_pairs(const int32_t* u, const int32_t* i, const float* rating, int64_t n, int32_t* out_u,
                            int32_t* out_i, int64_t* n_pairs_dev, void* workspace, int64_t workspace_bytes,
                            apr_stream_t stream);

/* ---- N3: one training batch of the Keras Recommender models -- MF.py:7-59 `MatrixFactorization`
 *      (loss_kind 0: binary_crossentropy of the RAW dot product clipped to [1e-7, 1 - 1e-7], MF.py:21-24; inputs users,
 *      items, labels y) and BPR.py:23-102 `BPR` (loss_kind 1: mean of 1 - log sigmoid(<u,i> - <u,j>), BPR.py:11-21; inputs
 *      users, positive items, negative items j) -- with Keras 2.2's Adam, which densifies the Embedding gradients: every
 *      row's moments decay and every row moves at every batch.  mP/vP/mQ/vQ: Adam moments (zero-initialised, table
 *      shapes); gP/gQ: dense gradient scratch of the table shapes, zero on entry and left zero; t: 1-based iteration;
 *      *loss_sum (device double) += sum of the batch's per-instance losses.  Parity: Keras unpinned (oracle.keras_step). */
int apr_keras_step(float* P, float* Q, float* mP, float* vP, float* mQ, float* vQ, float* gP, float* gQ, int64_t rows_p,
                   int64_t rows_q, int32_t d, const int32_t* u, const int32_t* i, const int32_t* j, const float* y,
                   int32_t n, int32_t loss_kind, float lr, float beta1, float beta2, int64_t t, double* loss_sum,
                   apr_stream_t stream);

/* ---- K11: np.linalg.norm(embedding_P) of utils.py:92-97: *out (device double) = sum of squares. */
int apr_sum_squares(const float* x, int64_t n, double* out, apr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* APR_B200_H_ */
