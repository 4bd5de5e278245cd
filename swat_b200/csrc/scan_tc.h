// Host-callable launchers of the swat_b200 kernels (internal; the public boundary is include/swat_b200.h).
#pragma once
#include "common.cuh"

namespace swat {

constexpr int kTcEpiWarps = 8;   // survivor lists per CTA of the tcgen05 kernel

struct TcArgs {
  ScanArgs s;
  int32_t n_qb;           // Q blocks (each keeps <=256 padded query columns resident in shared memory)
  int32_t n_blk;          // padded columns per Q block, multiple of 16
  int32_t n_stages;       // bank-tile ring depth
  uint32_t smem_b_bytes;  // per-CTA bytes of the resident query block
  uint64_t bank_hint;     // L2 cache policy for bank tiles
  int32_t unit_base;      // unit plan: first unit of this launch (unit = unit_base + pair)
  int32_t n_ranges;       // unit plan: tile ranges R (0 = single launch, Q block = pair % n_qb)
  const int32_t* blk_class;  // [n_qb+1] first class of each Q block
  const int32_t* blk_split;  // [n_qb] grouped reduces: column (multiple of 32) where the second epilogue warp set starts
};

size_t tc_smem_bytes(int n_blk, int ctas, int n_stages);
int tc_pick_stages(int n_blk, int ctas, size_t smem_limit);
// tm_bank / tm_q: CUtensorMap*
cudaError_t launch_scan_tc(const void* tm_bank, const void* tm_q, const TcArgs& p, int ctas, int reduce,
                           bool partitioned, bool dense, int grid, cudaStream_t stream);

// SIMT fp32-FMA scan: any dtype, optional in-pass T2I predicate (bank2), exact reference arithmetic order
cudaError_t launch_scan_simt(const ScanArgs& a, const void* bank, const void* bank2, const void* queries_padded,
                             int dtype, int reduce, bool partitioned, bool dense, cudaStream_t stream);

// final per-class select of the k_fetch best candidates
// final per-class select: final thresholds from the histograms, partition of the survivor lists by
// class, radix-select + sort of the k_fetch best candidates (3 launches)
cudaError_t launch_select(const JobState& st, int n_classes, int64_t row_offset, float* d_scores, int64_t* d_rows,
                          int32_t* d_counts, int32_t* d_truncated, cudaStream_t stream);
constexpr int kSelectLaunches = 3;

struct T2iArgs {
  const void* img_bank; int dtype; int64_t img_rows; int64_t img_row_base; const int64_t* img_index;
  const void* queries;          // [Q,512] unpadded, same dtype as the bank
  const int32_t* class_begin;   // [C+1] first query of each class
  int reduce;
  const float* cand_scores; const int64_t* cand_rows; const int32_t* cand_counts; const int32_t* truncated;
  int k_fetch; int k; float t2i_thr; int n_classes;
  float* t2i_scratch;           // [C, k_fetch]
  float* out_scores; int64_t* out_rows; float* out_t2i; int32_t* out_counts; int32_t* incomplete;
};
cudaError_t launch_t2i_walk(const T2iArgs& a, cudaStream_t stream);

cudaError_t launch_argmax_rows(const float* d_scores, int64_t n_rows, int n_classes, int32_t* d_pred, cudaStream_t stream);
cudaError_t launch_merge(const float* d_scores, const int64_t* d_rows, const float* d_aux, float aux_thr,
                         const int32_t* d_counts, const int32_t* d_truncated, int n_shards, int64_t shard_stride_bytes, int n_classes,
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
