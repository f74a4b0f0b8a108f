/*
 * pxr.h — C ABI of libpxr.so: the B200 (sm_100a) full-catalogue scoring /
 * top-K ranking / ranking-metrics path of PixelRec_Multimodal.
 *
 * The reference has no FFI: this path sits behind duck-typed Python objects
 * (SURVEY.md §8(b)).  Each entry point below names the reference interface it
 * replaces (paths relative to the reference root).  The Python host
 * (pixelrec_multimodal_b200/) binds these with ctypes; device buffers are owned
 * by the caller (PyTorch) and passed as raw pointers + sizes + a cudaStream_t.
 *
 * Conventions
 *   - every call returns 0 on success or a negative pxr_status; nothing throws
 *     or aborts across the boundary; pxr_last_error() gives the text.
 *   - all device work is asynchronous on the given stream.
 *   - a handle is not thread-safe: one handle per GPU per thread.
 *   - the library allocates device memory only in pxr_load_weights (folded /
 *     bf16 weight images, < 4 MB) and frees it in pxr_destroy; everything else
 *     lives in caller-provided workspaces whose sizes the *_bytes calls report.
 *   - matrices are row-major; indices are int64 where the reference uses
 *     torch.long and int32 for item ids inside CSR / top-K lists.
 */
#ifndef PXR_H_
#define PXR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PXR_VERSION 2
#define PXR_MAX_HIDDEN 8
#define PXR_MAX_KS 8
#define PXR_METRIC_COLS 9

typedef struct pxr_handle pxr_handle;
typedef void* pxr_stream; /* cudaStream_t */

typedef enum {
  PXR_OK = 0,
  PXR_ERR_INVALID = -1,     /* bad argument / unsupported configuration */
  PXR_ERR_CUDA = -2,        /* a CUDA runtime call failed               */
  PXR_ERR_STATE = -3,       /* call order (weights / items not loaded)  */
  PXR_ERR_WORKSPACE = -4    /* workspace too small                      */
} pxr_status;

/* model.fusion_type, src/config.py:43; src/models/multimodal.py:583-593 */
typedef enum { PXR_FUSION_CONCAT = 0, PXR_FUSION_GATED = 1, PXR_FUSION_ATTENTION = 2 } pxr_fusion;
/* fusion_activation, src/models/multimodal.py:150-167 */
typedef enum { PXR_ACT_RELU = 0, PXR_ACT_GELU = 1, PXR_ACT_TANH = 2, PXR_ACT_LEAKY_RELU = 3, PXR_ACT_SILU = 4 } pxr_act;
/* final_activation, src/models/multimodal.py:381-384 */
typedef enum { PXR_FINAL_NONE = 0, PXR_FINAL_SIGMOID = 1, PXR_FINAL_TANH = 2 } pxr_final;
/* 16-bit operand format of the tcgen05 path (accumulation is always fp32).  bf16: any range, 8-bit significand;
 * fp16: 11-bit significand (about 7x lower score error), conversions saturate at +-65504. */
typedef enum { PXR_PRECISION_BF16 = 0, PXR_PRECISION_FP16 = 1 } pxr_precision;
/* which scoring kernel pxr_score_topk uses */
typedef enum { PXR_PATH_AUTO = 0, PXR_PATH_SIMT = 1, PXR_PATH_TCGEN05 = 2 } pxr_path;

/* Mirrors the constructor arguments of MultimodalRecommender that shape the
 * computation (src/models/multimodal.py:42-66). */
typedef struct {
  int32_t struct_size;          /* sizeof(pxr_config), for ABI checking            */
  int32_t fusion;               /* pxr_fusion                                      */
  int32_t embedding_dim;        /* D                                               */
  int32_t vision_dim;           /* Dv of the cached vision features, 0 = absent    */
  int32_t language_dim;         /* Dl of the cached text features, 0 = absent      */
  int32_t num_numerical;        /* F, 0 = absent                                   */
  int32_t projection_hidden;    /* projection_hidden_dim, 0 = single-layer         */
  int32_t n_hidden;             /* len(fusion_hidden_dims)                         */
  int32_t hidden[PXR_MAX_HIDDEN];
  int32_t num_heads;            /* attention fusion only                           */
  int32_t activation;           /* pxr_act                                         */
  int32_t final_activation;     /* pxr_final                                       */
  int32_t use_batch_norm;
  int32_t n_tags;
  int32_t path;                 /* pxr_path; PXR_PATH_AUTO picks tcgen05 when the  */
                                /* shape is supported, else the SIMT kernels       */
  int32_t precision;            /* pxr_precision of the tcgen05 path               */
} pxr_config;

/* fp32 DEVICE pointers named after the reference state_dict keys
 * (SURVEY.md §8(a) A1); NULL where the configuration has no such tensor.
 * The big user / item embedding tables are NOT copied: they are passed to the
 * calls that gather from them. */
typedef struct {
  int32_t struct_size;
  const float* tag_embedding;            /* tag_embedding.weight        (n_tags, D)  */
  const float* vision_w0;  const float* vision_b0;   /* vision_projection.0   (D|P, Dv)   */
  const float* vision_w1;  const float* vision_b1;   /* vision_projection.3   (D, P)      */
  const float* language_w0; const float* language_b0;/* language_projection.0             */
  const float* language_w1; const float* language_b1;/* language_projection.3             */
  const float* numerical_w0; const float* numerical_b0;
  const float* numerical_w1; const float* numerical_b1;
  const float* gate_w;  const float* gate_b;         /* fusion_layer.gating_network.0 (M, M*D) */
  const float* attn_in_w; const float* attn_in_b;    /* fusion_layer.attention.in_proj_* (3D, D) */
  const float* attn_out_w; const float* attn_out_b;  /* fusion_layer.attention.out_proj.* (D, D) */
  const float* attn_ln_w; const float* attn_ln_b;    /* fusion_layer.norm.*                      */
  const float* mlp_w[PXR_MAX_HIDDEN];    /* prediction_network.{0,4,8..}.weight     */
  const float* mlp_b[PXR_MAX_HIDDEN];
  const float* bn_w[PXR_MAX_HIDDEN];     /* prediction_network.{2,6,10..}.weight    */
  const float* bn_b[PXR_MAX_HIDDEN];
  const float* bn_mean[PXR_MAX_HIDDEN];
  const float* bn_var[PXR_MAX_HIDDEN];
  const float* out_w;  const float* out_b;           /* last Linear (1, H_L)                     */
  float bn_eps;                                      /* 1e-5 (nn.BatchNorm1d default)            */
} pxr_weights;

int pxr_version(void);
const char* pxr_last_error(const pxr_handle* h);     /* h may be NULL: last create error */

/* replaces MultimodalRecommender.__init__ (src/models/multimodal.py:42-148) */
int pxr_create(const pxr_config* cfg, pxr_handle** out);
void pxr_destroy(pxr_handle* h);

/* replaces load_state_dict of the checkpoint (scripts/evaluate.py:366-375):
 * folds eval-mode BatchNorm into the next Linear, splits the first Linear into
 * per-user / per-item halves, builds the bf16 tcgen05 operand images. */
int pxr_load_weights(pxr_handle* h, const pxr_weights* w, pxr_stream stream);

/* K1 + K2.  Replaces the per-pair item half of forward
 * (src/models/multimodal.py:554-570) and the per-user re-stacking of item
 * features (src/inference/recommender.py:162-191): gathers item / tag
 * embeddings, runs the modality projections once per item and stores the
 * per-item partial record the scoring kernels consume.
 *   item_embedding : (>= max(item_idx)+1, D) fp32 table
 *   item_idx       : (n_rows,) int64 rows of the table, NULL = item_base + r
 *   tag_idx        : (n_rows,) int64
 *   vis/txt/num    : (n_rows, Dv|Dl|F) fp32, NULL where the modality is absent
 *   item_base      : global index of row 0 (item-axis shard offset)
 * The records stay in `workspace` (pxr_items_bytes(h, n_rows) bytes, 256-byte
 * aligned), which must outlive every scoring call that uses them. */
size_t pxr_items_bytes(const pxr_handle* h, int64_t n_rows);
int pxr_precompute_items(pxr_handle* h, const float* item_embedding, const int64_t* item_idx,
                         const int64_t* tag_idx, const float* vis, const float* txt, const float* num,
                         int64_t n_rows, int64_t item_base, void* workspace, size_t workspace_bytes,
                         pxr_stream stream);

/* Items whose features could not be fetched.  The reference scores them 0.0 instead of running the model
 * (src/inference/recommender.py:199-201, 229-230: `final_scores_map.get(original_id, 0.0)`), and they take part in
 * the ranking with that score.  flags: (n_rows,) uint8 DEVICE array aligned with the precomputed rows, 1 = missing;
 * caller-owned, must outlive the scoring calls; NULL clears.  Call after pxr_precompute_items (which clears it). */
int pxr_set_missing_items(pxr_handle* h, const uint8_t* flags, int64_t n_rows);

/* K3 (+ in-kernel top-K).  Replaces Recommender.get_recommendations with
 * candidates=None for a batch of users (src/inference/recommender.py:52-110):
 * every user of the batch against every precomputed item row, seen items
 * dropped, stable descending order (ties -> lower item index), first K.
 *   user_embedding : (n_users_total, D) fp32 table;  user_idx : (n_users,) int64
 *   seen_indptr    : (n_users+1,) int64 CSR offsets for this batch, NULL = no filter
 *   seen_idx       : int32 GLOBAL item indices, ascending inside each user
 *   out_scores     : (n_users, K) fp32, descending, padded with -inf
 *   out_idx        : (n_users, K) int32 GLOBAL item indices, padded with -1 */
size_t pxr_score_topk_bytes(const pxr_handle* h, int64_t n_users, int32_t k);
int pxr_score_topk(pxr_handle* h, const float* user_embedding, const int64_t* user_idx, int64_t n_users,
                   const int64_t* seen_indptr, const int32_t* seen_idx, int32_t k,
                   float* out_scores, int32_t* out_idx, void* workspace, size_t workspace_bytes,
                   pxr_stream stream);

/* Exact mode of pxr_score_topk on the tcgen05 path (default: on).  The fused kernel ranks with 16-bit operands
 * (bf16: |score error| up to ~5e-2 at catalogue scale); in exact mode it keeps its 64 best
 * candidates per user, these are scored again with the fp32 arithmetic of pxr_score_pairs -- i.e. the reference
 * forward, src/models/multimodal.py:528-610 -- and ranked again with the reference's stable order
 * (src/inference/recommender.py:105-106, ties -> lower item index).  Returned scores then meet the fp32 tolerance
 * (1e-4) and the list equals the reference's whenever its top-K lies inside the 16-bit top-64.  The SIMT path is
 * always exact.  on = 0 returns the 16-bit scores and order as they are.
 * top_k > 64 (up to 1 024): the kernel's per-user list has 64 slots, so the fused kernel runs once per 64-slot PAGE --
 * page p admits only keys strictly below the user's last key of page p - 1 -- and the ceil(K / 64) * 64 candidates are
 * re-scored / returned as above (cost: ceil(K / 64) passes; round 1 sent such calls to the generic fp32 kernels). */
int pxr_set_rescore(pxr_handle* h, int on);
int pxr_get_rescore(const pxr_handle* h);

/* Small user batches on the fused path (gated / concat fusion, <= 8 users, K <= 64): the unit's 16 user slots become (user, item
 * sub-range) pairs, so that one Recommender.get_recommendations call (src/inference/recommender.py:52-110 scores ONE user
 * per call) is not capped at 1/16 of the tile; the per-slot lists are merged by one or two pxr_merge_topk passes.  Same
 * arithmetic per pair: the lists equal those of the plain tile shape bit for bit.
 *   mode: -1 = when the built-in cost model expects a gain (default; ~ >= 30 K items per user), 0 = never, 1 = whenever possible */
int pxr_set_small_batch(pxr_handle* h, int mode);

/* The re-score step of exact mode on its own: candidate lists (for instance the merged 16-bit top-64 lists of several item
 * shards) are scored with the fp32 arithmetic of pxr_score_pairs against the records of THIS handle and ranked (ties ->
 * lower item index); the first K come back.  Under item-axis sharding the exchange carries the raw 64-slot lists and the
 * rank that owns a user re-scores the merged list once, against fp32 records of the whole catalogue (1 280 B per item),
 * instead of every shard re-scoring its own 64 candidates of every user.
 *   cand_idx : (n_users, 64) int32 GLOBAL item indices inside [item_base, item_base + n_rows) of this handle, -1 padded
 *   out_*    : (n_users, K), K <= 64, padded with -inf / -1 ;  workspace : pxr_rescore_bytes(n_users) bytes */
size_t pxr_rescore_bytes(int64_t n_users);
int pxr_rescore_topk(pxr_handle* h, const float* user_embedding, const int64_t* user_idx, int64_t n_users,
                     const int32_t* cand_idx, int32_t k, float* out_scores, int32_t* out_idx, void* workspace,
                     size_t workspace_bytes, pxr_stream stream);
/* The same for candidate lists of list_len = 64 * pages slots (pages <= 16; what a top_k > 64 call keeps):
 *   cand_idx : (n_users, list_len) ;  out_* : (n_users, K), K <= list_len ;  workspace : pxr_rescore_lists_bytes() bytes */
size_t pxr_rescore_lists_bytes(int64_t n_users, int32_t list_len);
int pxr_rescore_lists(pxr_handle* h, const float* user_embedding, const int64_t* user_idx, int64_t n_users,
                      const int32_t* cand_idx, int32_t list_len, int32_t k, float* out_scores, int32_t* out_idx,
                      void* workspace, size_t workspace_bytes, pxr_stream stream);

/* on = 1: pxr_precompute_items of this handle keeps only the fp32 item records (1 280 B per item at D = 64) and skips the
 * fused kernel's per-item extras (up to 16.5 KB per item for attention): for a handle that only serves pxr_rescore_topk /
 * pxr_score_pairs, e.g. the whole-catalogue re-score records every rank of an item-sharded job keeps.  Call before
 * pxr_precompute_items. */
int pxr_set_records_only(pxr_handle* h, int on);

/* Scores of explicit (user, item-row) pairs against the precomputed records.
 * Replaces MultimodalRecommender.forward (src/models/multimodal.py:528-610),
 * Recommender._score_items_batch / get_item_score and the candidate-list mode
 * of get_recommendations (src/inference/recommender.py:81-82,112-236).
 *   item_row : (n,) int64 LOCAL row numbers inside the precomputed records
 *   out      : (n,) fp32 scores after the final activation and NaN/Inf guard
 *   out_logit: optional (n,) fp32 pre-activation logits (NULL to skip) */
int pxr_score_pairs(pxr_handle* h, const float* user_embedding, const int64_t* user_idx,
                    const int64_t* item_row, int64_t n, float* out, float* out_logit, pxr_stream stream);

/* K4.  Merge S per-shard top-K lists per user (the step after the NCCL
 * all-gather of SURVEY.md §8(e)); ties -> lower global item index.
 *   scores_in / idx_in : (S, n_users, K) ;  out_* : (n_users, K)
 * Each input list must be sorted the way the scoring entry points write it: score descending, ties by ascending index, -1 /
 * -inf padding at the tail.  Scores are ordered by their IEEE bit pattern, so -0.0 sorts below +0.0 (Python's float compare
 * calls them equal); the library's own kernels emit +0.0 for missing-feature items and NaN, so this only matters for caller-made lists. */
int pxr_merge_topk(const float* scores_in, const int32_t* idx_in, int32_t n_shards, int64_t n_users,
                   int32_t k, float* out_scores, int32_t* out_idx, pxr_stream stream);

/* K5.  Replaces the accuracy block of TopKRetrievalEvaluator.evaluate and
 * _calculate_ndcg (src/evaluation/tasks.py:567-635, 718-747) and the standalone
 * per-user functions of src/evaluation/metrics.py:11-133 for several cut-offs at
 * once (@10 is a prefix of @50).
 *   topk_idx   : (n_users, k_stride) int32 ranked GLOBAL item ids, -1 padded
 *   gt_indptr  : (n_users+1,) int64 ;  gt_idx : int32 relevant items (any order), the
 *                SET of positives of each user (its size is what IDCG and MAP use)
 *   recall_den : (n_users,) int32 recall denominators = len(positive_items), the raw
 *                number of test rows of the user (tasks.py:579: duplicates and items
 *                unknown to the encoder included); NULL = the set size
 *   ks         : host array of n_ks cut-offs, each <= k_stride
 *   discount   : (k_stride,) float64 DEVICE table 1/log2(i+2) (host-computed so it
 *                is bit-identical to numpy's);  ideal : (k_stride+1,) float64
 *                DEVICE prefix sums of it in Python's left-to-right order
 *   out_sums   : (n_ks, PXR_METRIC_COLS = 9) float64 DEVICE: sums over users of
 *                precision (hits / len(recs), tasks.py:577), recall, f1, hit_rate,
 *                ndcg (tasks.py variant), mrr, ndcg (metrics.py variant,
 *                src/evaluation/metrics.py:63-100), precision (hits / k,
 *                metrics.py:29-35), average precision of the first k entries
 *                (metrics.py:102-133)
 *   workspace  : pxr_metrics_bytes(n_users, n_ks) bytes */
size_t pxr_metrics_bytes(int64_t n_users, int32_t n_ks);
int pxr_metrics(const int32_t* topk_idx, int32_t k_stride, int64_t n_users, const int64_t* gt_indptr,
                const int32_t* gt_idx, const int32_t* recall_den, const int32_t* ks, int32_t n_ks,
                const double* discount, const double* ideal, double* out_sums, void* workspace,
                size_t workspace_bytes, pxr_stream stream);

/* Sampled evaluation protocol, candidate construction.  Replaces
 * TopKRetrievalEvaluator._process_user / _sample_negatives with sampling_strategy 'random'
 * (src/evaluation/tasks.py:181-224, 310-364): per user, the positives plus n_neg negatives drawn uniformly
 * without replacement from the non-positive items, in a shuffled order.  Every draw is a pure function of
 * (seed, global user index): a keyed Feistel permutation of [0, n_items) for the negatives, a 64-bit hash order
 * for the shuffle (the reference seeds with Python's per-process salted hash() and is not reproducible).
 *   user_idx   : (n_users,) int64 global user indices (NULL = 0..n_users-1), only used as sampling keys
 *   pos_indptr : (n_users+1,) int64 ;  pos_idx : int32 positives, ASCENDING inside each user
 *   stride     : row length of out_cand (<= 1024); positives beyond it are dropped, negatives fill what is left
 *   out_cand   : (n_users, stride) int32 candidate item indices, -1 padded ;  out_len : (n_users,) int32 */
int pxr_sample_candidates(const int64_t* user_idx, int64_t n_users, const int64_t* pos_indptr, const int32_t* pos_idx,
                          int64_t n_items, int32_t n_neg, uint64_t seed, int32_t stride, int32_t* out_cand,
                          int32_t* out_len, pxr_stream stream);

/* The same for sampling_strategy 'popularity' / 'popularity_inverse' (src/evaluation/tasks.py:225-308:
 * np.random.choice(pool, n, replace=False, p = normalised item weight)): item i of user u gets the key
 * log(uniform(seed, u, i)) / weights[i]; the n_neg largest keys among the non-positive items are the sample
 * (Efraimidis-Spirakis weighted sampling without replacement), ties -> lower item; shuffle order as above.
 * One pass over the catalogue per user, nothing of size users x items is stored.
 *   weights : (n_items,) float64, > 0 (test-set item counts or their reciprocals; n_items < 2^31) */
int pxr_weighted_candidates(const int64_t* user_idx, int64_t n_users, const int64_t* pos_indptr, const int32_t* pos_idx,
                            const double* weights, int64_t n_items, int32_t n_neg, uint64_t seed, int32_t stride,
                            int32_t* out_cand, int32_t* out_len, pxr_stream stream);

/* K3t on its own: exact top-K of every row of a dense score matrix; -inf entries are masked; ties -> lower
 * column.  Used to rank candidate lists (stable sort of src/inference/recommender.py:105 over the candidate order).
 *   scores : (n_rows, n_cols) fp32 ;  out_scores / out_pos : (n_rows, K), padded with -inf / -1 */
int pxr_topk_rows(pxr_handle* h, const float* scores, int64_t n_rows, int64_t n_cols, int32_t k, float* out_scores,
                  int32_t* out_pos, pxr_stream stream);

/* Beyond-accuracy metrics over the ranked lists.  Replaces the per-user loop of
 * NoveltyMetrics.calculate_metrics (src/evaluation/novelty.py:84-147, 149-226, 343-377) and
 * TopKRetrievalEvaluator._calculate_personalization (src/evaluation/tasks.py:402-427) as aggregated in
 * TopKRetrievalEvaluator.evaluate (tasks.py:637-714).
 *   topk_idx    : (n_users, k_stride <= 64) int32 ranked item ids, -1 padded
 *   self_info   : (n_items,) float64 DEVICE table -log2(max(pop_i / total, 1e-10)), NaN where the item has no
 *                 popularity entry;  iif : (n_items,) float64 log(n_hist_users / (count_i + 1e-10)), NaN likewise
 *   hist_indptr / hist_idx : interaction history CSR of these users (ascending per user), NULL = no history
 *   out6        : DEVICE float64: sum over users with a non-empty list of [mean self-information, mean IIF,
 *                 #distinct items, fraction of items outside the history, 1], then sum_i s_i^2 with
 *                 s_i = sum over lists containing item i of 1/sqrt(|list|)   (pairwise cosine of the binary
 *                 list vectors: sum_{u<v} cos = (sum_i s_i^2 - #non-empty lists) / 2)
 *   workspace   : pxr_novelty_bytes(n_users, n_items) bytes */
size_t pxr_novelty_bytes(int64_t n_users, int64_t n_items);
int pxr_novelty_metrics(const int32_t* topk_idx, int32_t k_stride, int64_t n_users, int64_t n_items,
                        const double* self_info, const double* iif, const int64_t* hist_indptr, const int32_t* hist_idx,
                        double* out6, void* workspace, size_t workspace_bytes, pxr_stream stream);

/* Gini coefficient of the per-item recommendation counts over the ranked lists.  Replaces
 * AdvancedMetrics.calculate_gini_coefficient (src/evaluation/advanced_metrics.py:72-105) applied to
 * {item: number of lists holding it}: counts sorted ascending, G = 2 sum_i i c_(i) / (n sum c) - (n + 1) / n.
 *   include_zero : 1 = every one of the n_items items is in the distribution (never-recommended items count 0, as in
 *                  the reference's own test, tests/unit/src/evaluation/test_advanced_metrics.py:70-73); 0 = only items
 *                  recommended at least once
 *   out3         : DEVICE float64 [gini, n, sum of counts] ;  workspace : pxr_gini_bytes(n_users, n_items) bytes */
size_t pxr_gini_bytes(int64_t n_users, int64_t n_items);
int pxr_gini(const int32_t* topk_idx, int32_t k_stride, int64_t n_users, int64_t n_items, int32_t include_zero,
             double* out3, void* workspace, size_t workspace_bytes, pxr_stream stream);

/* Intra-list similarity: per user the mean pairwise cosine similarity of the embeddings of its listed items; lists
 * with fewer than two embedded items give 0.0.  Replaces NoveltyMetrics.calculate_diversity
 * (src/evaluation/novelty.py:295-340) as averaged in TopKRetrievalEvaluator.evaluate (src/evaluation/tasks.py:695-701;
 * the reference's own embedding collection, tasks.py:430-507, raises NameError and yields nothing).
 *   emb       : (n_rows, dim <= 512) fp32 DEVICE item embeddings, row r = item item_base + r; NULL = the item records
 *               resident in `h` for scoring (the projected item-side modality vectors of pxr_precompute_items)
 *   out2      : DEVICE float64 [sum over users of their intra-list similarity, users counted]
 *   workspace : pxr_ils_bytes(n_users, n_rows) bytes (n_rows of the handle when emb is NULL) */
size_t pxr_ils_bytes(int64_t n_users, int64_t n_rows);
int pxr_intra_list_similarity(pxr_handle* h, const int32_t* topk_idx, int32_t k_stride, int64_t n_users, const float* emb,
                              int64_t n_rows, int32_t dim, int64_t item_base, double* out2, void* workspace,
                              size_t workspace_bytes, pxr_stream stream);

/* Live timing of the dominant kernel (the pair-scoring kernel of
 * pxr_score_topk) with CUDA events recorded on the launching stream, for the
 * roofline line of bench.py.  pxr_profile_read synchronises on the recorded
 * events, returns the summed duration and launch count since the last read and
 * resets both.  At most PXR_PROFILE_SLOTS launches are kept between reads. */
#define PXR_PROFILE_SLOTS 1024
int pxr_profile_enable(pxr_handle* h, int on);
int pxr_profile_read(pxr_handle* h, double* total_ms, int64_t* n_launches);

/* Introspection used by bench.py / tests. */
int64_t pxr_launch_count(const pxr_handle* h);       /* kernels launched through this handle */
int pxr_active_path(const pxr_handle* h);            /* pxr_path pxr_score_topk will use      */
/* why PXR_PATH_AUTO resolved to the generic fp32 SIMT kernels ("" when the fused tcgen05 kernel is active): the
 * generic path is ~100x slower, so the host logs this once per model (it is exact, and it too keeps the running
 * top-K on chip: no users x items score matrix in HBM).  A call with top_k > 1 024 on a tcgen05 handle also takes it. */
const char* pxr_path_reason(const pxr_handle* h);
int pxr_set_path(pxr_handle* h, int path);           /* force SIMT / tcgen05 (tests)          */

#ifdef __cplusplus
}
#endif
#endif /* PXR_H_ */
