// Host-callable launchers of the swat_b200 kernels (internal; the public boundary is include/swat_b200.h).
#pragma once
#include "common.cuh"

namespace swat {

constexpr int kTcEpiWarps = 8;   // survivor lists per CTA of the tcgen05 kernel

struct TcArgs {
  ScanArgs s;
  int32_t n_qb;           // Q blocks (each keeps <=256 padded query columns resident in shared memory)
  int32_t n_blk;          // padded columns per Q block, multiple of 16
  int32_t n_stages;       // operand ring depth (stages of 128 rows x 64 k bf16)
  int32_t n_fstages;      // fp32 banks: ring depth of the staged fp32 boxes (128 rows x 32 k)
  uint32_t smem_b_bytes;  // per-CTA bytes of the resident query block
  uint64_t bank_hint;     // L2 cache policy for bank tiles
  int32_t unit_base;      // unit plan: first unit of this launch (unit = unit_base + pair)
  int32_t n_ranges;       // unit plan: tile ranges R (0 = legacy launch)
  int32_t qb_base;        // legacy launch: covers the Q blocks [qb_base, qb_base + qb_count), block = qb_base + pair % qb_count
  int32_t qb_count;
  uint32_t* progress;     // nullable [pairs of this launch]: tiles each pair has requested so far (lockstep of the pairs sharing a tile range)
  int32_t lock_window;    // a pair requests tile i only when every pair of its range has requested tile i - lock_window
  const int32_t* blk_class;  // [n_qb+1] first class of each Q block
  const int32_t* blk_split;  // [n_qb] grouped reduces: column (multiple of 32) where the second epilogue warp set starts
  unsigned long long* trace; // nullable [grid][8]: globaltimer stamps of each CTA's phases (SWAT_SCAN_TRACE diagnostics)
  // nullable: dynamic tile scheduling (one Q block, selecting scan).  [0] = tile counter, then per pair a ring of
  // kTileRing entries ((iteration + 1) << 32 | tile): the leader CTA's producer claims tiles three iterations ahead
  // and publishes them here for the other warps of the pair.  Zeroed before every launch.
  unsigned long long* tile_sched;
  // nullable: CTA 0 writes the SM clock it observed over the launch (MHz = clock64 ticks / globaltimer time) here; the
  // host uses it to tell a power-capped GPU (automatic lockstep window, api.cu)
  unsigned long long* clock_probe;
};
constexpr int kTileRing = 16;

size_t tc_smem_bytes(int n_blk, int ctas, int n_stages, int n_fstages);
int tc_pick_stages(int n_blk, int ctas, size_t smem_limit);
int tc_pick_fstages(int n_blk, int ctas, size_t smem_limit, int* op_stages, int prefer_op);
// tm_bank / tm_q: CUtensorMap*; f32: tm_bank describes an fp32 bank (boxes of 128 rows x 32 k)
cudaError_t launch_scan_tc(const void* tm_bank, const void* tm_q, const TcArgs& p, int ctas, int reduce,
                           bool partitioned, bool dense, bool f32, int grid, cudaStream_t stream);

// SIMT fp32-FMA scan: any dtype, optional in-pass T2I predicate (bank2), exact reference arithmetic order
cudaError_t launch_scan_simt(const ScanArgs& a, const void* bank, const void* bank2, const void* queries_padded,
                             int dtype, int reduce, bool partitioned, bool dense, cudaStream_t stream);

// final per-class select of the k_fetch best candidates
// final per-class select: final thresholds from the histograms, partition of the survivor lists by
// class, radix-select + sort of the k_fetch best candidates (3 launches)
// band_k > 0: cut every list behind the first candidate scoring below (band_k-th best score - band); see select_kernel
// zero_word (nullable): one int32 the select clears on the way (the walk's eps_violation word)
cudaError_t launch_select(const JobState& st, int n_classes, int64_t row_offset, float* d_scores, int64_t* d_rows,
                          int32_t* d_counts, int32_t* d_truncated, cudaStream_t stream, uint32_t band_k = 0, float band = 0.0f,
                          int32_t* zero_word = nullptr);
constexpr int kSelectLaunches = 3;

// Exact re-score + accept walk over per-class candidate lists (select.cu).
// The scan kernels rank rows by an APPROXIMATE score (tensor-core fp32 accumulation order; bf16-rounded operands for
// fp32 banks): |approx - exact| <= eps.  Every candidate is re-scored with one fixed-order fp32 dot (the canonical
// score: identical whichever engine produced the list), the list is re-sorted on (exact desc, row asc) and walked.
// A truncated list (more eligible rows existed) only vouches for rows scoring above its frontier: the approximate
// score of its last candidate plus eps.
struct WalkArgs {
  const void* t2t_bank;         // rows the candidates were ranked on (exact T2T re-score)
  const void* aux_bank;         // nullable: predicate bank (image rows for T2T-rank-T2I-tshd, sample_retrieval.py:804-806)
  int dtype;                    // of both banks
  int64_t bank_rows;            // rows in the bank views
  int64_t bank_row_base;        // id (as stored in cand_rows) of bank row 0
  int64_t key_row_base;         // cand_rows - key_row_base fits 32 bits (the shard's row_offset): tie-break key
  const int64_t* gather_index;  // nullable [C*stride]: the banks are compact gathers, candidate e reads row gather_index[e]
  const void* queries;          // [Q,512] unpadded, same dtype as the banks
  const int32_t* class_begin;   // [C+1] first query of each class
  int reduce;
  // the predicate may have its own query set (few-shot image prompts, t2t_rank_i2t_tshd_sampler :831-890); same classes
  const void* aux_queries; const int32_t* aux_class_begin; int aux_reduce;
  const float* cand_scores; const int64_t* cand_rows; const int32_t* cand_counts; const int32_t* truncated;   // [C,stride] / [C]
  int stride; int k; int n_classes;
  float thr, aux_thr, eps;
  int lazy_t2t;                 // 1: fetch a candidate's ranking row only if its predicate row passes (rows over PCIe)
  int all_or_nothing;           // 1: a truncated list vouches for nothing (lists ranked on another metric: bank-swap pass)
  float* exact_scratch; float* aux_scratch;      // [C,stride] each
  float* out_scores; int64_t* out_rows; float* out_aux; int32_t* out_counts;   // [C,k] / [C]
  float* out_limit;             // nullable [C]: -inf = proven exact, else rows scoring <= limit may be missing
  int32_t* incomplete;          // nullable [C]: 1 = fewer than k accepted and the list cannot vouch for that
  int32_t* eps_violation;       // nullable [1]: set when some candidate's exact score is further than eps from the score
                                // the scan ranked it by -- the error bound does not hold (rows not L2-normalised?)
  // nullable [2C+2]: everything the host reads back after a step, in one block: [0] the job's overflow word
  // (*job_flags), [1..C] incomplete flags, [1+C..2C] accepted counts ([1+2C] is the eps_violation word)
  int32_t* status; const uint32_t* job_flags;
};
cudaError_t launch_rescore_walk(const WalkArgs& a, cudaStream_t stream);
constexpr int kWalkLaunches = 2;

// out[i] = canonical score of bank row i against the queries of class row_class[i] (-inf where row_class[i] < 0)
cudaError_t launch_score_rows(const void* bank, int dtype, const int32_t* row_class, int64_t n_rows, const void* queries,
                              const int32_t* class_begin, int n_classes, int reduce, float* out, cudaStream_t stream);
// result rows of n classes: dst[d_dst_cls[i]] <- src[d_src_idx[i]] ([.,k] scores / rows / aux, [.] counts)
cudaError_t launch_splice(const int32_t* d_dst_cls, const int32_t* d_src_idx, int n, int k, const float* s_scores, const int64_t* s_rows,
                          const float* s_aux, const int32_t* s_counts, float* d_scores, int64_t* d_rows, float* d_aux, int32_t* d_counts,
                          cudaStream_t stream);
// row_class of a sub-query run: out[i] = map[in[i]] (map: original class -> sub class or -1), -1 stays -1
cudaError_t launch_remap_classes(const int32_t* in, const int32_t* map, int n_map, int64_t n, int32_t* out, cudaStream_t stream);

cudaError_t launch_argmax_rows(const float* d_scores, int64_t n_rows, int n_classes, int32_t* d_pred, cudaStream_t stream);
// d_limit [G,C] nullable: rows of shard g scoring <= d_limit[g][c] may be missing from its list
cudaError_t launch_merge(const float* d_scores, const int64_t* d_rows, const float* d_aux, float aux_thr,
                         const int32_t* d_counts, const float* d_limit, int n_shards, int64_t shard_stride_bytes, int n_classes,
                         int k, int k_out, uint64_t* d_key_scratch, float* d_out_scores, int64_t* d_out_rows, float* d_out_aux,
                         int32_t* d_out_counts, int32_t* d_incomplete, cudaStream_t stream);

cudaError_t launch_job_reset(const JobState& st, int n_classes, cudaStream_t stream);
// per class: flag rows that have an earlier row of the class with cosine > threshold (select.cu)
cudaError_t launch_near_dup(const void* bank, int dtype, const int64_t* d_order, const int32_t* d_class_start, int n_classes,
                            int max_class_rows, float threshold, uint8_t* d_dup, cudaStream_t stream);
// seed a fresh job from dense prefix scores [C][n_prefix] (see select.cu)
cudaError_t launch_bootstrap(const JobState& st, int n_classes, const float* d_scores_t, uint32_t n_prefix, uint32_t row_base,
                             uint32_t first_spare_list, cudaStream_t stream);

}  // namespace swat
