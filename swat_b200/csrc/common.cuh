// Shared device-side definitions for the swat_b200 scan / select kernels.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

namespace swat {

constexpr int kDim = 512;
constexpr int kHistBins = 1024;          // per-class score histogram (32 lanes x 32 bins)
constexpr int kSortCap = 4096;           // block-level bitonic sort capacity (u64 keys)
constexpr int kMaxKFetch = 4096;         // k_fetch <= kSortCap

enum : int { RED_NONE = 0, RED_MEAN = 1, RED_MAX = 2, RED_MIN = 3 };

// ---- ordered-uint encoding of fp32 (monotone: a < b  <=>  enc(a) < enc(b); NaN-free inputs) ----
__host__ __device__ __forceinline__ uint32_t f32_enc(float f) {
#ifdef __CUDA_ARCH__
  uint32_t b = __float_as_uint(f);
#else
  union { float f; uint32_t u; } x; x.f = f; uint32_t b = x.u;
#endif
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float f32_dec(uint32_t e) {
  uint32_t b = (e & 0x80000000u) ? (e & 0x7fffffffu) : ~e;
#ifdef __CUDA_ARCH__
  return __uint_as_float(b);
#else
  union { float f; uint32_t u; } x; x.u = b; return x.f;
#endif
}
// candidate key: sorting keys DESCENDING gives (score descending, row ascending) -- the reference's
// stable sorted(..., reverse=True) order (sample_retrieval.py:754, :807).
__host__ __device__ __forceinline__ uint64_t make_key(float score, uint32_t row) {
  return (static_cast<uint64_t>(f32_enc(score)) << 32) | static_cast<uint64_t>(~row);
}
__host__ __device__ __forceinline__ float key_score(uint64_t k) { return f32_dec(static_cast<uint32_t>(k >> 32)); }
__host__ __device__ __forceinline__ uint32_t key_row(uint64_t k) { return ~static_cast<uint32_t>(k); }

// ---- running top-k state of one job (device pointers) ----
// Scan kernels only ever (a) bump the class histogram (fire-and-forget RED) and (b) append
// {key, class} to a survivor list; nothing in the scan waits on an atomic's return value.  At select
// time the final class thresholds are derived from the histograms, the lists are partitioned by class
// (dropping everything below the final threshold) into `cand`, and each class is radix-selected/sorted.
struct JobState {
  uint32_t* tau_enc;     // [C]  f32_enc of the class threshold: rows scoring below can never be in the top k_fetch
  uint32_t* hist;        // [C * kHistBins] histogram of appended scores
  uint4* list;           // [n_lists * list_cap] survivor entries {key.lo, key.hi, class, 0}
  uint32_t* list_count;  // [n_lists] entries appended to each list (may exceed list_cap after an overflow)
  uint32_t* count;       // [C]  candidates per class after partition
  uint64_t* cand;        // [C * cap] candidate keys after partition
  uint32_t* flags;       // [0] bit0 = class candidate overflow (partition), bit1 = survivor list overflow (scan)
  const uint32_t* k_class;  // nullable [C]: per-class k_fetch (<= k_fetch); classes whose T2I walk needs depth get more
  uint32_t n_lists, list_cap;
  uint32_t cap;
  uint32_t k_fetch;
  float thr;             // user T2T threshold (sample_retrieval.py:1576 passes 0.0)
  float hist_lo;         // histogram covers [hist_lo, 1]
  float hist_scale;      // bins per unit score
  float hist_inv_scale;
};

// per-launch parameters common to both scan kernels
struct ScanArgs {
  JobState st;
  const int32_t* col_class;   // [n_cols_padded] global class index of each query column, -1 = padding
  const float* col_count;     // [n_cols_padded] group size R at the group's last column, 0 elsewhere
  int32_t n_cols;             // padded query columns (all Q blocks)
  int64_t n_rows;             // rows in this view
  uint32_t row_base;          // shard-local id of row 0 of the view
  const int32_t* row_class;   // nullable, [n_rows]
  const uint32_t* exclude;    // nullable, bitmap over view rows
  float t2i_thr;              // dual (in-pass T2I) mode only
  float* dense_out;           // nullable: write class scores instead of selecting
  int64_t dense_ld;           // leading dimension of dense_out
  int32_t dense_transposed;   // 0: dense_out[row*ld + class], 1: dense_out[class*ld + row] (coalesced)
  int32_t n_classes;
  // In-pass predicate on the tensor cores, two passes.  Pass A (dense mode over the predicate bank): bit (class, row)
  // of bits_out = class score >= bits_thr.  Pass B (selecting scan of the ranking bank): a survivor is kept only if
  // its bit in pass_bits is set.  Both are [n_classes][bits_words] words over shard rows (row_base + view row).
  uint32_t* bits_out;
  const uint32_t* pass_bits;
  int64_t bits_words;
  float bits_thr;
};

__device__ __forceinline__ uint32_t class_k(const JobState& st, int cls) { return st.k_class ? st.k_class[cls] : st.k_fetch; }

__device__ __forceinline__ int hist_bin(const JobState& st, float s) {
  float x = (s - st.hist_lo) * st.hist_scale;
  int b = static_cast<int>(x);
  b = b < 0 ? 0 : b;
  return b > kHistBins - 1 ? kHistBins - 1 : b;
}
// a threshold such that every score counted in bins >= b is >= it (half a bin of slack absorbs
// the fp32 rounding of hist_bin)
__device__ __forceinline__ float hist_edge(const JobState& st, int b) {
  return st.hist_lo + (static_cast<float>(b) - 0.5f) * st.hist_inv_scale;
}

__device__ __forceinline__ uint32_t ld_cg_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ uint4 ld_cg_u32x4(const uint32_t* p) {
  uint4 v;
  asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

// finite identities: the epilogue tests the sign of (value - threshold) and +-inf would turn
// inf - inf into NaN
template <int RED> __device__ __forceinline__ float red_init() {
  if (RED == RED_MAX) return -3.402823466e+38f;
  if (RED == RED_MIN) return 3.402823466e+38f;
  return 0.0f;
}
template <int RED> __device__ __forceinline__ float red_op(float a, float x) {
  if (RED == RED_MAX) return fmaxf(a, x);
  if (RED == RED_MIN) return fminf(a, x);
  if (RED == RED_MEAN) return a + x;
  return x;
}
template <int RED> __device__ __forceinline__ float red_fin(float a, float cnt) {
  if (RED == RED_MEAN) return __fdiv_rn(a, cnt);
  return a;
}

}  // namespace swat
