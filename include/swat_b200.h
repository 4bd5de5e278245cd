/*
 * swat_b200.h -- C-ABI of the B200-native retrieval hot path (score -> per-class top-k -> T2I walk).
 *
 * The reference (tian1327/SWAT) has no FFI: the path is inline Python in
 * retrieval/sample_retrieval.py.  Every entry point below therefore cites the reference call site
 * (file:line into /root/reference/retrieval/sample_retrieval.py) whose work it replaces; the Python
 * mirror of the reference's functions (swat_b200/retrieval.py) binds these with ctypes and
 * INTEGRATION.md shows the stub a SWAT maintainer would add.
 *
 * Conventions
 *  - plain C linkage, pointers and sizes only; no C++/torch types cross the boundary;
 *  - every function returns 0 on success or a negative swat_status; swat_last_error() gives a
 *    thread-local message for the last failure on the calling thread;
 *  - "d_" pointers are device memory on the context's device, "h_" pointers are host memory;
 *    the caller owns every buffer it passes in; the library owns what *_create returns;
 *  - device work is enqueued on the caller's stream (cudaStream_t passed as void*, NULL = legacy
 *    default stream) and is asynchronous unless the comment says it synchronises;
 *  - one swat_ctx per device; a ctx and its children are not thread-safe, distinct ctxs are;
 *  - rows are D = 512 wide (OpenCLIP ViT-B/32 embedding size, utils/extras.py:97-114);
 *  - result order inside a class is the reference's walk order: score descending, ties by
 *    ascending row id (Python's stable sorted(..., reverse=True), :754, :807);
 *  - there is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef SWAT_B200_H
#define SWAT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SWAT_DIM 512
#define SWAT_VERSION 200 /* 0.2.0 */

typedef enum {
  SWAT_OK = 0,
  SWAT_ERR_INVALID = -1,     /* bad argument */
  SWAT_ERR_CUDA = -2,        /* CUDA runtime / driver error (message has the CUDA error string) */
  SWAT_ERR_NO_DEVICE = -3,   /* no CUDA device / wrong architecture (needs sm_100) */
  SWAT_ERR_OVERFLOW = -4,    /* candidate buffers overflowed and the retry budget was exhausted */
  SWAT_ERR_INCOMPLETE = -5,  /* T2I walk could not be proven exact within the escalation budget */
  SWAT_ERR_UNSUPPORTED = -6  /* valid request this build cannot serve (e.g. k too large) */
} swat_status;

typedef enum { SWAT_BF16 = 0, SWAT_F32 = 1 } swat_dtype;

/* per-class reduce over the query columns of one class:
 * NONE: one query per class, the re-normalised mean prompt (:749-750, :801-802);
 * MEAN: torch.mean over R prompt columns (:403-404, :341-342);
 * MAX / MIN: i2i_similarity_p2p modes (:377-385); MAX is the north star's "max over synonyms". */
typedef enum { SWAT_REDUCE_NONE = 0, SWAT_REDUCE_MEAN = 1, SWAT_REDUCE_MAX = 2, SWAT_REDUCE_MIN = 3 } swat_reduce;

/* which scan kernel: AUTO = the tcgen05 kernel (bf16 banks natively; fp32 banks are converted to bf16 on the fly by
 * converter warps and every candidate is re-scored exactly in fp32 afterwards); SIMT = the fp32-FMA kernel (any
 * dtype; exact dense scores of fp32 banks, the single-pass two-bank predicate of swat_job_scan, the on-device checker). */
typedef enum { SWAT_ENGINE_AUTO = 0, SWAT_ENGINE_TC = 1, SWAT_ENGINE_SIMT = 2 } swat_engine;

typedef struct swat_ctx swat_ctx;
typedef struct swat_queries swat_queries;
typedef struct swat_job swat_job;

int32_t swat_version(void);
const char* swat_last_error(void);

/* One context per device.  Fails with SWAT_ERR_NO_DEVICE when the device is not sm_100. */
int32_t swat_ctx_create(int32_t device, swat_ctx** out);
int32_t swat_ctx_destroy(swat_ctx* ctx);
/* tuning knobs (all optional): "cta_group" (1|2, before swat_queries_create), "max_ctas", "cand_cap"
 * (per-class candidates kept after the final threshold), "list_entries" (total survivor-list entries),
 * "overfetch" (first k_fetch of the T2I walk), "host_chunk_rows"; 0 = automatic.  Switches (default 1): "unit_plan",
 * "swap_pass", "zero_copy" (host pipeline reads candidates' rows from pinned banks in place), "dyn_tiles" (one query
 * block: CTA pairs claim bank tiles from a global counter instead of a fixed stride); "bootstrap_rows" (dense prefix
 * that seeds the thresholds, default 32768, 0 = off), "lock_window" (several query blocks: pairs sharing a tile range
 * stay within this many tiles of each other; 0 = off, default -1 = automatic: 4 for 4-8 query blocks while the scan
 * observes a power-capped SM clock), "f32_op_stages".  Environment: SWAT_DEBUG=1 logs
 * allocations, SWAT_SCAN_TRACE=1 prints per-launch phase stamps of the scan kernel (diagnostics: synchronises). */
int32_t swat_ctx_set_option(swat_ctx* ctx, const char* name, int64_t value);
/* counters since ctx creation: kernels launched by this library (bench.py's gpu_launches claim) */
int64_t swat_ctx_launch_count(const swat_ctx* ctx);

/* The prompt tensors of the reference: prompt_tensors[cls]['mean'] ([512], one query per class) or
 * ['all'] ([P_c,512], a group per class) (utils/features.py:39-64; consumed at :749-750, :801-802).
 * h_queries: [n_queries, 512] fp32 host, rows of one class adjacent; h_class_of_query: [n_queries]
 * dense class index 0..n_classes-1, non-decreasing (NULL = identity, needs n_queries == n_classes).
 * Uploads an fp32 copy and a bf16 (round-to-nearest-even) copy. Synchronises. */
int32_t swat_queries_create(swat_ctx* ctx, const float* h_queries, int32_t n_queries,
                            const int32_t* h_class_of_query, int32_t n_classes, int32_t reduce,
                            swat_queries** out);
int32_t swat_queries_destroy(swat_queries* q);

/* ---- feature-shard loader (torch.load(...) + .cuda(), sample_retrieval.py:1473-1476, :337, :399) -------------- */
/* Rows [row_begin, row_end) of a flat shard file (raw row-major [n_rows,512] bf16 | f32: caption.bin / image.bin of
 * swat_b200/shards.py, converted once from the reference's *_mined.pth) -> d_dst, caller-allocated
 * [row_end-row_begin, 512] device memory: a rank of a sharded run loads its own row range.  Banks are plain device
 * pointers in this ABI, so there is no separate "bank from device" call.  pread() into two pinned staging buffers of
 * chunk_rows rows (0 = 65536), the read of chunk i+1 overlapping the H2D copy of chunk i on `stream`; with SWAT_GDS=1 in
 * the environment and a libcufile that accepts the file, GPUDirect Storage instead (cuFileRead straight into d_dst,
 * *used_gds = 1).  Synchronises. */
int32_t swat_bank_load(swat_ctx* ctx, const char* path, int32_t dtype, int64_t row_begin, int64_t row_end, void* d_dst,
                       int64_t chunk_rows, int32_t* used_gds, void* stream);

/* ---- streaming job: running per-class top-k_fetch over any number of bank views ------------- */
/* Replaces, for all classes at once, the per-class  t2t_similarity -> sorted() -> walk  of
 * t2t_ranked_sampler (:752-758): state = per-class threshold, histogram and candidate buffer. */
int32_t swat_job_create(swat_ctx* ctx, const swat_queries* q, int32_t k_fetch, float t2t_threshold,
                        swat_job** out);
int32_t swat_job_reset(swat_job* job, void* stream);
/* Per-class depth: class c keeps its best h_depth[c] rows (1 <= h_depth[c] <= k_fetch) instead of
 * k_fetch; NULL restores the uniform depth.  Lets a T2I walk over-fetch deeply only for the classes
 * that need it.  Call before the first swat_job_scan after a reset. */
int32_t swat_job_set_class_depth(swat_job* job, const int32_t* h_depth, void* stream);
/* Score one bank view (rows [row_base, row_base+n_rows) of the shard) against every query and fold
 * it into the job.  d_bank: [n_rows,512] row-major, 16-byte aligned, dtype bf16|f32.
 * d_t2i_bank (nullable): same rows of the image bank; when given, the predicate
 * t2i >= t2i_threshold is evaluated in the same pass for every row on the fp32-FMA kernel (2x the bytes; the
 * whole-pipeline calls use two tensor-core passes with a per-class bitmap instead).
 * d_row_class (nullable): [n_rows] dense class index of each row, -1 = none: the reference's
 * partitioned case, a row is eligible only for its own class (transform_extracted_fea :1387-1415).
 * d_exclude (nullable): bitmap over the view's rows, bit set = never accept
 * (duplicates_dict / filtered_images_dict of add_to_split :454-456). */
int32_t swat_job_scan(swat_job* job, const void* d_bank, int32_t dtype, int64_t n_rows, int64_t row_base,
                      const void* d_t2i_bank, float t2i_threshold, const int32_t* d_row_class,
                      const uint32_t* d_exclude, int32_t engine, void* stream);
/* Sorted top-k_fetch of every class: d_scores [C,k_fetch] f32, d_rows [C,k_fetch] i64 (row_offset +
 * the shard-local row id passed via row_base; -1 padded), d_counts [C] i32, d_truncated [C] i32
 * (nullable; 1 = more than k_fetch rows were eligible, i.e. the list is a strict prefix of the walk). */
int32_t swat_job_select(swat_job* job, int64_t row_offset, float* d_scores, int64_t* d_rows, int32_t* d_counts,
                        int32_t* d_truncated, void* stream);
/* Asynchronously copies the job's overflow word (see swat_job_status) to *d_flags on `stream`, so a
 * multi-GPU caller can ship it with the candidates instead of synchronising before the exchange. */
int32_t swat_job_export_flags(swat_job* job, int32_t* d_flags, void* stream);
/* Synchronises the stream the job last ran on; *overflowed != 0 means the results are invalid:
 * bit0 = a class candidate buffer overflowed (raise "cand_cap"), bit1 = a survivor list overflowed
 * (raise "list_entries").  swat_topk / swat_topk_host retry by themselves. */
int32_t swat_job_status(swat_job* job, int32_t* overflowed);
int32_t swat_job_destroy(swat_job* job);

/* ---- exact re-score + accept walk: add_to_split (:439-482), cal_t2i_similarity (:335-353) +
 *      add_t2t_ranked_t2i_tshd_to_split (:492-540) ------------------------------------------------------------- */
/* The scan kernels rank rows by an APPROXIMATE score: tensor-core accumulation order for bf16 banks, bf16-rounded rows
 * and queries for fp32 banks.  *eps bounds |approximate - canonical| for the given bank dtype and engine
 * (SWAT_ENGINE_AUTO = what swat_job_scan would pick). */
int32_t swat_scan_eps(const swat_queries* q, int32_t dtype, int32_t engine, float* eps);
/* Candidates of each class ([C,k_fetch] as swat_job_select writes them): re-score every candidate with the canonical
 * fixed-order fp32 dot against d_t2t_bank (and d_aux_bank when given: the predicate bank, e.g. the image rows), re-sort
 * on (exact score desc, row asc) -- the reference's stable sorted(..., reverse=True) (:754, :807) -- and walk:
 * accept rows with exact >= t2t_threshold and aux >= aux_threshold, stop at k.  Bank row r is the row whose id in
 * d_cand_rows is bank_row_base + r.  A truncated list vouches only for rows scoring above (approximate score of its
 * last candidate + eps).  d_out_limit [C] (nullable): -inf = the class is proven exact, else rows scoring <= limit
 * may be missing; d_incomplete [C] (nullable): 1 = fewer than k accepted and not proven (escalate k_fetch).
 * q_aux (nullable = q): the predicate's own query set over the same classes -- the few-shot image prompts of
 * t2t_rank_i2t_tshd_sampler / t2t_rank_i2i_tshd_sampler (:869, :929). */
int32_t swat_rescore_walk(swat_ctx* ctx, const swat_queries* q, const swat_queries* q_aux,
                          const void* d_t2t_bank, const void* d_aux_bank,
                          int32_t dtype, int64_t bank_rows, int64_t bank_row_base,
                          const float* d_cand_scores, const int64_t* d_cand_rows, const int32_t* d_cand_counts,
                          const int32_t* d_truncated, int32_t k_fetch, int32_t k, float t2t_threshold,
                          float aux_threshold, float eps, float* d_out_scores, int64_t* d_out_rows, float* d_out_aux,
                          int32_t* d_out_counts, float* d_out_limit, int32_t* d_incomplete, void* stream);

/* ---- multi-GPU: merge after the single NCCL gather ------------------------------ */
/* d_scores/d_rows/d_aux: [G,C,k_in] gathered per-shard walk results (canonical scores; rows global, < 2^32),
 * d_counts [G,C], d_limit [G,C] (nullable): shard g vouches only for rows scoring above d_limit[g][c] (-inf = its
 * list is complete; what swat_rescore_walk wrote).  shard_stride_bytes == 0: the arrays are contiguous;
 * otherwise shard g of EVERY array starts g * shard_stride_bytes after shard 0 (the arrays are slices
 * of one packed per-rank buffer, as an all-gather delivers them).  Keeps, per class, the k_out best entries under
 * (score desc, row asc) among those with aux >= aux_threshold (d_aux == NULL: no predicate).
 * d_incomplete [C] (nullable): 1 = the result reaches down to some shard's limit (rows that shard never reported
 * could belong in it): re-run the shards with a larger k_fetch. */
int32_t swat_merge_topk(swat_ctx* ctx, const float* d_scores, const int64_t* d_rows, const float* d_aux,
                        const int32_t* d_counts, const float* d_limit, int32_t n_shards,
                        int64_t shard_stride_bytes, int32_t n_classes,
                        int32_t k_in, int32_t k_out, float aux_threshold, float* d_out_scores, int64_t* d_out_rows,
                        float* d_out_aux, int32_t* d_out_counts, int32_t* d_incomplete, void* stream);

/* ---- S1 compatibility: t2t_similarity / cal_t2i_similarity (:397-416, :335-353) ---------------- */
/* d_out [n_rows, n_classes] f32 class scores (after the reduce).  Tests and back-compat only: the
 * fast path never materialises this matrix. */
int32_t swat_scores_dense(swat_ctx* ctx, const swat_queries* q, const void* d_bank, int32_t dtype,
                          int64_t n_rows, float* d_out, int32_t engine, void* stream);

/* Partitioned data (one class per row, transform_extracted_fea :1387-1415): d_out[i] = canonical score of row i against
 * the queries of ITS OWN class d_row_class[i] (-inf where that is < 0) -- t2t_similarity / cal_t2i_similarity of every
 * class over its own rows in one pass (:752, :804-806).  Feeds the walk diagnostics (filtered_list.txt :463-469). */
int32_t swat_score_rows(swat_ctx* ctx, const swat_queries* q, const void* d_bank, int32_t dtype, int64_t n_rows,
                        const int32_t* d_row_class, float* d_out, void* stream);

/* ---- exclusion-set producer: zeroshot_clip_img_filter (:278-329) --------------------------------- */
/* d_pred [n_rows] i32: argmax over the class scores of every row (lowest class on ties) -- the prediction of the
 * zero-shot head `MyLinear(weights = stacked class prompts, bias = False)` (:1489-1492, :299-301).  Rows whose
 * prediction differs from their own class go into filtered_images_dict.  The scores are produced in row chunks
 * by the same scan kernels (dense mode) and never leave the device. */
int32_t swat_zeroshot_predict(swat_ctx* ctx, const swat_queries* q, const void* d_bank, int32_t dtype,
                              int64_t n_rows, int32_t* d_pred, int32_t engine, void* stream);

/* ---- exclusion-set producer: remove_near_duplicates2 (:237-275) --------------------------------- */
/* d_order [n]: bank row ids grouped by class (file order kept inside a class), d_class_start [C+1]:
 * first position of each class in d_order.  d_dup [n] (by position in d_order) is set to 1 for every
 * row that has an EARLIER row of its class with cosine > threshold (the reference uses 0.9, :257).
 * The caller zeroes d_dup. */
int32_t swat_near_duplicates(swat_ctx* ctx, const void* d_bank, int32_t dtype, int64_t n_rows, const int64_t* d_order,
                             const int32_t* d_class_start, int32_t n_classes, int32_t max_class_rows, float threshold,
                             uint8_t* d_dup, void* stream);

/* ---- whole pipeline on HBM-resident banks ------------------------------------------------------ */
/* t2t_ranked_sampler (:724-771) when d_t2i_bank == NULL, t2t_ranked_t2i_tshd_sampler (:774-825)
 * otherwise, for all classes at once.  Outputs [C,k] (d_out_t2i nullable), rows are
 * row_offset + local row (row_offset + n_rows < 2^32 - 1).  Scores are canonical (see swat_rescore_walk): identical
 * whichever engine, shard count or escalation path produced them.  Handles candidate-buffer overflow and over-fetch
 * escalation internally: deeper over-fetch for the classes that need it, then a pass over the image bank that
 * enumerates the rows able to pass T2I (classes with few of them), finally the two-pass in-pass predicate (an
 * image-bank pass writes a per-class bitmap of passing rows, the caption scan keeps only survivors whose bit is set).
 * A class that cannot be proven exact at the widest over-fetch (more than ~3500 rows tying with its k-th score)
 * fails the call with SWAT_ERR_INCOMPLETE -- never a silently wrong row.  fp32 banks: rows are assumed L2-normalised
 * (extract_mined_feature.py:121,181; utils/features.py:30-31), which is what bounds the error of their bf16-rounded
 * scan.  k <= 4096.  Synchronises. */
int32_t swat_topk(swat_ctx* ctx, const swat_queries* q, const void* d_t2t_bank, const void* d_t2i_bank,
                  int32_t dtype, int64_t n_rows, int64_t row_offset, int32_t k, float t2t_threshold,
                  float t2i_threshold, const int32_t* d_row_class, const uint32_t* d_exclude,
                  float* d_out_scores, int64_t* d_out_rows, float* d_out_t2i, int32_t* d_out_counts,
                  void* stream);

/* ---- whole pipeline on HOST banks (the reference's torch.load'ed CPU tensors, :1473-1476) ------ */
/* Streams the caption bank host->device in chunks overlapped with the scan, gathers only the
 * candidates' image rows for the T2I stage, returns results in host memory.  Pinned host memory
 * gives full PCIe rate.  Synchronises. */
int32_t swat_topk_host(swat_ctx* ctx, const swat_queries* q, const void* h_t2t_bank, const void* h_t2i_bank,
                       int32_t dtype, int64_t n_rows, int64_t row_offset, int32_t k, float t2t_threshold,
                       float t2i_threshold, const int32_t* h_row_class, const uint32_t* h_exclude,
                       float* h_out_scores, int64_t* h_out_rows, float* h_out_t2i, int32_t* h_out_counts);

/* timing of the last swat_topk / swat_topk_host on this ctx, CUDA-event milliseconds:
 * [0] scan kernels, [1] select, [2] T2I stage, [3] whole call; plus [4] scan launches,
 * [5] H2D bytes, [6] D2H bytes, [7] escalation rounds. */
int32_t swat_ctx_last_timing(const swat_ctx* ctx, double out[8]);

#ifdef __cplusplus
}
#endif
#endif /* SWAT_B200_H */
