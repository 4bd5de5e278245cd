// Per-class selection kernels: final top-k_fetch of a job, T2I re-score + accept walk, shard merge.
#include <algorithm>
#include "common.cuh"
#include "epilogue.cuh"
#include "scan_tc.h"

namespace swat {
namespace {

constexpr int kSelThreads = 1024;

// Exact top-`K` of the non-zero u64 keys in keys[0..n) (zero = absent entry; real keys are unique),
// sorted descending into s_keys.  MSB-first 8-bit radix select narrows the candidate set until it fits
// the block sort (usually 0-2 passes), then a bitonic sort orders it.  `n_valid` = number of non-zero
// keys.  Returns the number of sorted entries `total` (every key >= the radix prefix); the caller
// takes the first min(K, total).
struct GlobalKeys {
  const uint64_t* keys;
  __device__ __forceinline__ uint64_t operator()(uint32_t i) const { return keys[i]; }
};
// keys made on the fly from a dense score column: rows below the user threshold are absent
struct ScoreKeys {
  const float* scores; float thr; uint32_t row_base;
  __device__ __forceinline__ uint64_t operator()(uint32_t i) const {
    const float s = scores[i] + 0.0f;
    return s >= thr ? make_key(s, row_base + i) : 0ull;
  }
};

// keys below a lower bound of the k-th best are absent (merge: most of G*k entries cannot be in the result)
struct FloorKeys {
  const uint64_t* keys; uint64_t floor;
  __device__ __forceinline__ uint64_t operator()(uint32_t i) const {
    const uint64_t k = keys[i];
    return k >= floor ? k : 0ull;
  }
};

// Descending bitonic sort of s_keys[0..P) by the whole block (P a power of two <= kSortCap; optional 16-bit payload
// moved with the keys).  Called after a __syncthreads(); ends with one.  A thread owns E = P / 1024 consecutive
// positions: every compare-exchange at distance j < 32 E stays inside a warp (registers and xor-shuffles, no barrier),
// only the steps at distance >= 32 E go through shared memory with a block barrier each -- 20 barriers instead of 55
// at P = 1024, 27 instead of 78 at P = 4096.
template <int E, bool IDX>
__device__ __forceinline__ void sort_warp_steps(uint64_t (&x)[E], uint32_t (&xi)[E], uint32_t pos0, uint32_t k, uint32_t j_from) {
  for (uint32_t j = j_from; j >= static_cast<uint32_t>(E); j >>= 1) {       // partner = same slot of lane ^ (j / E)
    const int lane_mask = static_cast<int>(j / E);
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const uint32_t i = pos0 + e;
      const uint64_t y = __shfl_xor_sync(0xffffffffu, x[e], lane_mask);
      const uint32_t yi = IDX ? __shfl_xor_sync(0xffffffffu, xi[e], lane_mask) : 0u;
      const bool keep_max = ((i & j) == 0) == ((i & k) == 0);                 // lower position of a descending pair, or upper of an ascending one
      const bool take = keep_max ? (y > x[e]) : (y < x[e]);
      if (take) { x[e] = y; xi[e] = yi; }
    }
  }
#pragma unroll
  for (int jj = E / 2; jj > 0; jj >>= 1) {                                    // partner inside the thread
    if (j_from >= static_cast<uint32_t>(jj)) {
#pragma unroll
      for (int e = 0; e < E; ++e) {
        if ((e & jj) == 0) {
          const bool desc = ((pos0 + e) & k) == 0;
          if ((x[e] < x[e | jj]) == desc) {
            const uint64_t t = x[e]; x[e] = x[e | jj]; x[e | jj] = t;
            const uint32_t ti = xi[e]; xi[e] = xi[e | jj]; xi[e | jj] = ti;
          }
        }
      }
    }
  }
}

template <int E, bool IDX>
__device__ void block_sort_desc_e(uint64_t* s_keys, uint16_t* s_idx, uint32_t P) {
  const uint32_t tid = threadIdx.x, pos0 = tid * E;
  const bool active = pos0 < P;
  constexpr uint32_t W = 32u * E;               // positions per warp
  uint64_t x[E];
  uint32_t xi[E];
  auto load = [&]() {
#pragma unroll
    for (int e = 0; e < E; ++e) {
      x[e] = active ? s_keys[pos0 + e] : 0ull;
      xi[e] = (IDX && active) ? s_idx[pos0 + e] : 0u;
    }
  };
  auto store = [&]() {
    if (active) {
#pragma unroll
      for (int e = 0; e < E; ++e) {
        s_keys[pos0 + e] = x[e];
        if (IDX) s_idx[pos0 + e] = static_cast<uint16_t>(xi[e]);
      }
    }
  };
  load();
  for (uint32_t k = 2; k <= min(W, P); k <<= 1) sort_warp_steps<E, IDX>(x, xi, pos0, k, k >> 1);
  store();
  __syncthreads();
  for (uint32_t k = 2 * W; k <= P; k <<= 1) {
    for (uint32_t j = k >> 1; j >= W; j >>= 1) {
      for (uint32_t i = tid; i < P; i += kSelThreads) {
        const uint32_t ixj = i ^ j;
        if (ixj > i) {
          const uint64_t a = s_keys[i], b = s_keys[ixj];
          const bool desc = (i & k) == 0;
          if ((a < b) == desc) {
            s_keys[i] = b; s_keys[ixj] = a;
            if (IDX) { const uint16_t t = s_idx[i]; s_idx[i] = s_idx[ixj]; s_idx[ixj] = t; }
          }
        }
      }
      __syncthreads();
    }
    load();
    sort_warp_steps<E, IDX>(x, xi, pos0, k, W >> 1);
    store();
    __syncthreads();
  }
}

template <bool IDX>
__device__ void block_sort_desc(uint64_t* s_keys, uint16_t* s_idx, uint32_t P) {
  if (P >= 4u * kSelThreads) block_sort_desc_e<4, IDX>(s_keys, s_idx, P);
  else if (P >= 2u * kSelThreads) block_sort_desc_e<2, IDX>(s_keys, s_idx, P);
  else block_sort_desc_e<1, IDX>(s_keys, s_idx, P);
}

template <typename KeyFn>
__device__ uint32_t select_sorted(const KeyFn keys, uint32_t n, uint32_t n_valid, uint32_t K, uint64_t* s_keys,
                                  uint32_t* s_hist, uint32_t* s_misc) {
  const int tid = threadIdx.x;
  uint64_t prefix = 0, mask = 0;
  uint32_t need = K, m = n_valid, sure = 0;
  int shift = 56;
  while (sure + m > static_cast<uint32_t>(kSortCap) && shift >= 0) {   // uniform: all values come from shared memory
    for (int i = tid; i < 256; i += kSelThreads) s_hist[i] = 0;
    __syncthreads();
    for (uint32_t i = tid; i < n; i += kSelThreads) {
      const uint64_t key = keys(i);
      if (key != 0ull && (key & mask) == prefix) atomicAdd(&s_hist[(key >> shift) & 0xFFu], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      uint32_t cum = 0;
      int d = 255;
      for (; d > 0; --d) {
        if (cum + s_hist[d] >= need) break;
        cum += s_hist[d];
      }
      s_misc[0] = cum;            // keys strictly above the chosen digit: certainly selected
      s_misc[1] = s_hist[d];      // keys matching the chosen digit: still undecided
      s_misc[2] = static_cast<uint32_t>(d);
    }
    __syncthreads();
    sure += s_misc[0];
    need -= min(need, s_misc[0]);
    m = s_misc[1];
    prefix |= static_cast<uint64_t>(s_misc[2]) << shift;
    mask |= 0xFFull << shift;
    shift -= 8;
    __syncthreads();
  }
  // gather every real key >= prefix (on the decided digits)
  if (tid == 0) s_misc[3] = 0;
  __syncthreads();
  for (uint32_t i0 = 0; i0 < n; i0 += kSelThreads) {       // warp-aggregated: one shared-memory atomic per warp and round
    const uint32_t i = i0 + tid;
    const uint64_t key = i < n ? keys(i) : 0ull;
    const bool keep = key != 0ull && (key & mask) >= prefix;
    const uint32_t ball = __ballot_sync(0xffffffffu, keep);
    uint32_t base = 0;
    if ((tid & 31) == 0 && ball) base = atomicAdd(&s_misc[3], static_cast<uint32_t>(__popc(ball)));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (keep) {
      const uint32_t pos = base + __popc(ball & ((1u << (tid & 31)) - 1u));
      if (pos < static_cast<uint32_t>(kSortCap)) s_keys[pos] = key;
    }
  }
  __syncthreads();
  const uint32_t total = min(s_misc[3], static_cast<uint32_t>(kSortCap));
  uint32_t P = 2;
  while (P < total) P <<= 1;
  for (uint32_t i = total + tid; i < P; i += kSelThreads) s_keys[i] = 0ull;   // real keys are never 0
  __syncthreads();
  block_sort_desc<false>(s_keys, nullptr, P);
  return total;
}

// band_k > 0 (walks without a predicate): the k-th best APPROXIMATE score a_k bounds what the walk can use.  The k best
// approximate rows score at least a_k - eps exactly, so a row ranked below a_k - 2 eps (= `band`, plus a rounding margin)
// can never be among the k best exact rows: the list is cut behind the first such candidate, which stays as the
// frontier marker (frontier = its score + eps < a_k - eps, so every possible top-k row is still vouched for).  The
// re-score then reads k + (rows within the band) rows per class instead of the whole over-fetch.
__global__ void __launch_bounds__(kSelThreads)
select_kernel(const JobState st, int64_t row_offset, float* __restrict__ out_scores, int64_t* __restrict__ out_rows,
              int32_t* __restrict__ out_counts, int32_t* __restrict__ out_trunc, uint32_t band_k, float band, int32_t* zero_word) {
  __shared__ uint64_t s_keys[kSortCap];
  __shared__ uint32_t s_hist[256];
  __shared__ uint32_t s_misc[4];
  const int c = blockIdx.x;
  const uint32_t appended = st.count[c];
  const uint32_t n = min(appended, st.cap);
  const uint32_t K = class_k(st, c);          // this class's depth
  const uint32_t stride = st.k_fetch;         // output row pitch = the deepest class
  const uint32_t total = select_sorted(GlobalKeys{st.cand + static_cast<size_t>(c) * st.cap}, n, n, K, s_keys, s_hist, s_misc);
  uint32_t cnt = min(K, total);
  bool cut = false;
  if (band_k > 0 && cnt > band_k) {
    const float floor_score = key_score(s_keys[band_k - 1]) - band;
    uint32_t lo = band_k, hi = cnt;           // first position whose score is below the band (scores descend)
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (key_score(s_keys[mid]) >= floor_score) lo = mid + 1; else hi = mid;
    }
    if (lo + 1 < cnt) { cnt = lo + 1; cut = true; }
  }
  for (uint32_t i = threadIdx.x; i < stride; i += kSelThreads) {
    const bool ok = i < cnt;
    const uint64_t key = ok ? s_keys[i] : 0ull;
    out_scores[static_cast<size_t>(c) * stride + i] = ok ? key_score(key) : 0.0f;
    out_rows[static_cast<size_t>(c) * stride + i] = ok ? static_cast<int64_t>(key_row(key)) + row_offset : -1;
  }
  if (threadIdx.x == 0) {
    if (zero_word && c == 0) *zero_word = 0;
    out_counts[c] = static_cast<int32_t>(cnt);
    // more eligible rows than k_fetch exist iff more than k_fetch candidates survived, or the
    // threshold ever rose above the user threshold (rows below it were dropped)
    if (out_trunc) out_trunc[c] = (cut || appended > K || st.tau_enc[c] > f32_enc(st.thr)) ? 1 : 0;
  }
}

__global__ void reset_kernel(const JobState st, int n_classes) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t nh = static_cast<size_t>(n_classes) * kHistBins;
  if (i < nh) st.hist[i] = 0;
  if (i < static_cast<size_t>(n_classes)) { st.count[i] = 0; st.tau_enc[i] = f32_enc(st.thr); }
  if (i < st.n_lists) st.list_count[i] = 0;
  if (i == 0) st.flags[0] = 0;
}

// final class thresholds from the complete histograms (one warp per class), and per-class counters
// zeroed for the partition
__global__ void __launch_bounds__(256) final_tau_kernel(const JobState st, int n_classes) {
  const int cls = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (cls >= n_classes) return;
  refresh_tau(st, cls);
  if ((threadIdx.x & 31) == 0) st.count[cls] = 0;
}

// survivor lists -> per-class candidate arrays, keeping only entries at or above the final class
// threshold (about k_fetch plus one histogram bin per class)
__global__ void __launch_bounds__(256) partition_kernel(const JobState st) {
  // direct version (label sets too large for the shared-memory counters below): one global atomic per kept entry
  const uint32_t l = blockIdx.x;
  const uint32_t n = min(st.list_count[l], st.list_cap);
  const uint4* src = st.list + static_cast<size_t>(l) * st.list_cap;
  for (uint32_t i = blockIdx.y * blockDim.x + threadIdx.x; i < n; i += gridDim.y * blockDim.x) {
    const uint4 e = src[i];
    const uint32_t cls = e.z;
    if (e.y >= st.tau_enc[cls]) {
      const uint32_t slot = atomicAdd(&st.count[cls], 1u);
      if (slot < st.cap) st.cand[static_cast<size_t>(cls) * st.cap + slot] = (static_cast<uint64_t>(e.y) << 32) | e.x;
      else atomicOr(st.flags, 1u);
    }
  }
}

// Block-aggregated version.  About k_fetch entries per class survive the final threshold, i.e. hundreds of thousands
// of atomics on a few hundred hot counters: same-address atomics serialise in L2 and that, not the 40 MB of list
// traffic, set the pace (67 us).  Each CTA owns a slice of the lists, counts its kept entries per class in shared
// memory, reserves one range per class with ONE global atomic, and scatters on a second pass over its (L2-resident) slice.
__global__ void __launch_bounds__(1024) partition_agg_kernel(const JobState st, int n_classes) {
  extern __shared__ uint32_t s_part[];
  uint32_t* s_cnt = s_part;                 // [n_classes] kept entries of this CTA, then running offsets
  uint32_t* s_base = s_part + n_classes;    // [n_classes] first slot of this CTA's range
  for (int c = threadIdx.x; c < n_classes; c += blockDim.x) s_cnt[c] = 0;
  __syncthreads();
  for (uint32_t l = blockIdx.x; l < st.n_lists; l += gridDim.x) {
    const uint32_t n = min(st.list_count[l], st.list_cap);
    const uint4* src = st.list + static_cast<size_t>(l) * st.list_cap;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
      const uint4 e = src[i];
      if (e.y >= st.tau_enc[e.z]) atomicAdd(&s_cnt[e.z], 1u);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < n_classes; c += blockDim.x) {
    const uint32_t m = s_cnt[c];
    s_base[c] = m ? atomicAdd(&st.count[c], m) : 0u;
    s_cnt[c] = 0;
  }
  __syncthreads();
  for (uint32_t l = blockIdx.x; l < st.n_lists; l += gridDim.x) {
    const uint32_t n = min(st.list_count[l], st.list_cap);
    const uint4* src = st.list + static_cast<size_t>(l) * st.list_cap;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
      const uint4 e = src[i];
      const uint32_t cls = e.z;
      if (e.y >= st.tau_enc[cls]) {
        const uint32_t slot = s_base[cls] + atomicAdd(&s_cnt[cls], 1u);
        if (slot < st.cap) st.cand[static_cast<size_t>(cls) * st.cap + slot] = (static_cast<uint64_t>(e.y) << 32) | e.x;
        else atomicOr(st.flags, 1u);
      }
    }
  }
}

// ---------------------------------------------------------------------------------- bootstrap
// Seeds a fresh job from the dense class scores of a bank prefix (scores_t[class][row], B rows).
// Pass 1 builds the class histogram of the prefix in shared memory and finds the highest bin edge
// with at least k_fetch scores at or above it; pass 2 appends exactly the rows at or above that edge
// (k_fetch plus about one bin: a superset of the prefix's top k_fetch, which is all the job needs) to
// a spare survivor list, adds those bins to the global histogram and raises the class threshold to
// the edge.  (34 us for 200 classes x 32 K rows; loads batched 16 at a time measured the same, register-cached scores
// with warp-aggregated atomics 71 us: the plain loops stay.)  The main scan then begins with selective thresholds instead of appending every
// non-negative score of its first waves.
__global__ void __launch_bounds__(kSelThreads)
bootstrap_kernel(const JobState st, const float* __restrict__ scores_t, uint32_t B, uint32_t row_base, uint32_t first_spare_list) {
  __shared__ uint32_t s_h[kHistBins];
  __shared__ uint32_t s_cut, s_total, s_base, s_fill;
  const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  const float* sc = scores_t + static_cast<size_t>(c) * B;
  for (int i = tid; i < kHistBins; i += kSelThreads) s_h[i] = 0;
  if (tid == 0) s_fill = 0;
  __syncthreads();
  for (uint32_t i = tid; i < B; i += kSelThreads) {
    const float s = sc[i] + 0.0f;
    if (s >= st.thr) atomicAdd(&s_h[hist_bin(st, s)], 1u);
  }
  __syncthreads();
  if (tid < 32) {   // warp 0: lane l owns bins 32l .. 32l+31; suffix-scan from the top
    uint32_t mine = 0;
    for (int b = 0; b < 32; ++b) mine += s_h[lane * 32 + b];
    uint32_t suf = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_down_sync(0xffffffffu, suf, d);
      if (lane + d < 32) suf += t;
    }
    const uint32_t K = class_k(st, c);
    const uint32_t ball = __ballot_sync(0xffffffffu, suf >= K);
    int cut = 0;                      // fewer than k_fetch eligible rows: keep them all
    uint32_t total = __shfl_sync(0xffffffffu, suf, 0);
    if (ball != 0u) {
      const int L = 31 - __clz(ball);
      int bin = 0;
      uint32_t run = 0, at = 0;
      if (lane == L) {
        run = suf - mine;
        for (int b = 31; b >= 0; --b) {
          run += s_h[lane * 32 + b];
          if (run >= K) { bin = lane * 32 + b; at = run; break; }
        }
      }
      cut = __shfl_sync(0xffffffffu, bin, L);
      total = __shfl_sync(0xffffffffu, at, L);
    }
    if (lane == 0) {
      s_cut = static_cast<uint32_t>(cut);
      s_total = total;
      const uint32_t list_id = first_spare_list + static_cast<uint32_t>(c) % (st.n_lists - first_spare_list);
      s_base = atomicAdd(&st.list_count[list_id], total);
      if (cut >= 1) atomicMax(&st.tau_enc[c], f32_enc(hist_edge(st, cut)));
    }
  }
  __syncthreads();
  const int cut = static_cast<int>(s_cut);
  const uint32_t list_id = first_spare_list + static_cast<uint32_t>(c) % (st.n_lists - first_spare_list);
  uint4* dst = st.list + static_cast<size_t>(list_id) * st.list_cap;
  for (uint32_t i = tid; i < B; i += kSelThreads) {
    const float s = sc[i] + 0.0f;
    if (s >= st.thr && hist_bin(st, s) >= cut) {
      const uint32_t slot = s_base + atomicAdd(&s_fill, 1u);
      const uint64_t key = make_key(s, row_base + i);
      if (slot < st.list_cap) dst[slot] = make_uint4(static_cast<uint32_t>(key), static_cast<uint32_t>(key >> 32), static_cast<uint32_t>(c), 0u);
      else atomicOr(st.flags, 2u);
    }
  }
  for (int b = cut + tid; b < kHistBins; b += kSelThreads)
    if (s_h[b]) atomicAdd(&st.hist[static_cast<size_t>(c) * kHistBins + b], s_h[b]);
}

// ---------------------------------------------------------------------------------- T2I stage
template <typename T> __device__ __forceinline__ void load16(const T* p, float (&x)[16]);
template <> __device__ __forceinline__ void load16<float>(const float* p, float (&x)[16]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 v = reinterpret_cast<const float4*>(p)[i];
    x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w;
  }
}
template <> __device__ __forceinline__ void load16<__nv_bfloat16>(const __nv_bfloat16* p, float (&x)[16]) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const uint4 v = reinterpret_cast<const uint4*>(p)[i];
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      x[8 * i + 2 * j] = __uint_as_float(w[j] << 16);
      x[8 * i + 2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
    }
  }
}

// Canonical score of one (row, class) pair: per lane 16 consecutive elements in a sequential FMA chain, then an
// xor-shuffle tree; group reduce over the class's queries in query order.  Every result the library returns is this
// value, whichever scan engine ranked the row (t2t_similarity / cal_t2i_similarity, sample_retrieval.py:397-416, :335-353).
template <typename T>
__device__ __forceinline__ float canonical_score(const float (&x)[16], const T* __restrict__ queries, int q0, int q1, int reduce, int lane) {
  float red = (reduce == RED_MAX) ? -INFINITY : (reduce == RED_MIN) ? INFINITY : 0.0f;
  for (int qi = q0; qi < q1; ++qi) {
    float q[16];
    load16<T>(queries + static_cast<size_t>(qi) * kDim + lane * 16, q);
    float d = 0.0f;
#pragma unroll
    for (int i = 0; i < 16; ++i) d = fmaf(x[i], q[i], d);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if (reduce == RED_MAX) red = fmaxf(red, d);
    else if (reduce == RED_MIN) red = fminf(red, d);
    else if (reduce == RED_MEAN) red += d;
    else red = d;
  }
  if (reduce == RED_MEAN) red = __fdiv_rn(red, static_cast<float>(q1 - q0));
  return red;
}

// 16 consecutive elements of a row as loaded (packed): unpacked to fp32 only when consumed, so that several rows can
// be in flight per lane without running out of registers
template <typename T> struct Raw16;
template <> struct Raw16<float> {
  float4 v[4];
  __device__ __forceinline__ void load(const float* p) {
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = reinterpret_cast<const float4*>(p)[i];
  }
  __device__ __forceinline__ void unpack(float (&x)[16]) const {
#pragma unroll
    for (int i = 0; i < 4; ++i) { x[4 * i] = v[i].x; x[4 * i + 1] = v[i].y; x[4 * i + 2] = v[i].z; x[4 * i + 3] = v[i].w; }
  }
};
template <> struct Raw16<__nv_bfloat16> {
  uint4 v[2];
  __device__ __forceinline__ void load(const __nv_bfloat16* p) {
    v[0] = reinterpret_cast<const uint4*>(p)[0];
    v[1] = reinterpret_cast<const uint4*>(p)[1];
  }
  __device__ __forceinline__ void unpack(float (&x)[16]) const {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        x[8 * i + 2 * j] = __uint_as_float(w[j] << 16);
        x[8 * i + 2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
      }
    }
  }
};

// Exact scores of the candidates: one warp per NC consecutive candidates of a class.  All of their rows (ranking bank
// and, with a predicate bank, the aux rows) are requested before any is consumed.  The kernel moves 2.6-3.2 TB/s of
// scattered 1-2 KB rows, which is what HBM gives this access pattern: staging the rows through shared memory
// (cp.async.bulk per row, then 16-byte cp.async, 96 KB in flight per CTA, ids and class tables prefetched a chunk
// ahead) was measured SLOWER on the B200 (193-294 us against 129 us for 200 x 1024 candidates of both banks).
// AUX = a predicate bank is read as well; without one a warp keeps twice as many ranking rows in flight.
template <typename T, int NC, bool AUX>
__global__ void __launch_bounds__(256, 2) rescore_kernel(const WalkArgs a) {
  const int lane = threadIdx.x & 31;
  const int groups = (a.stride + NC - 1) / NC;
  const int64_t w = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (w >= static_cast<int64_t>(a.n_classes) * groups) return;
  const int c = static_cast<int>(w / groups), j0 = static_cast<int>(w % groups) * NC;
  const size_t base = static_cast<size_t>(c) * a.stride;
  // everything the warp needs to know is requested at once (candidate count, row ids, approximate scores, query
  // ranges): one L2 round trip ahead of the row reads instead of a chain of three
  const int n_raw = a.cand_counts[c];
  int64_t r[NC];
  float approx[NC];
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    const bool in = j0 + i < a.stride;          // slots beyond the class's count hold stale ids: masked below
    r[i] = in ? (a.gather_index ? a.gather_index[base + j0 + i] : a.cand_rows[base + j0 + i] - a.bank_row_base) : -1;
    approx[i] = (in && a.eps_violation) ? a.cand_scores[base + j0 + i] : 0.0f;
  }
  const int q0 = a.class_begin[c], q1 = a.class_begin[c + 1];
  const int n = min(n_raw, a.stride);
  if (j0 >= n) return;
#pragma unroll
  for (int i = 0; i < NC; ++i)
    if (j0 + i >= n || r[i] >= a.bank_rows || r[i] < 0) r[i] = -1;
  Raw16<T> x[NC], y[AUX ? NC : 1];
  float aux[AUX ? NC : 1];
  if (AUX && a.lazy_t2t) {
    // rows are expensive to fetch (pinned host memory over PCIe): predicate rows first, ranking rows only for the
    // candidates that pass it -- the others can never be accepted and need no exact score
#pragma unroll
    for (int i = 0; i < NC; ++i)
      if (r[i] >= 0) y[AUX ? i : 0].load(static_cast<const T*>(a.aux_bank) + r[i] * kDim + lane * 16);
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      aux[AUX ? i : 0] = -INFINITY;
      if (r[i] >= 0) {
        float f[16];
        y[AUX ? i : 0].unpack(f);
        aux[AUX ? i : 0] = canonical_score<T>(f, static_cast<const T*>(a.aux_queries), a.aux_class_begin[c], a.aux_class_begin[c + 1], a.aux_reduce, lane);
      }
    }
#pragma unroll
    for (int i = 0; i < NC; ++i)
      if (r[i] >= 0 && aux[AUX ? i : 0] >= a.aux_thr) x[i].load(static_cast<const T*>(a.t2t_bank) + r[i] * kDim + lane * 16);
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      if (j0 + i >= n) break;                    // warp-uniform
      float t2t = -INFINITY;
      if (r[i] >= 0 && aux[AUX ? i : 0] >= a.aux_thr) {
        float f[16];
        x[i].unpack(f);
        t2t = canonical_score<T>(f, static_cast<const T*>(a.queries), q0, q1, a.reduce, lane);
      }
      if (lane == 0) {
        a.exact_scratch[base + j0 + i] = t2t;
        a.aux_scratch[base + j0 + i] = aux[AUX ? i : 0];
      }
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    if (r[i] >= 0) {
      x[i].load(static_cast<const T*>(a.t2t_bank) + r[i] * kDim + lane * 16);
      if (AUX) y[AUX ? i : 0].load(static_cast<const T*>(a.aux_bank) + r[i] * kDim + lane * 16);
    }
  }
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    if (j0 + i >= n) break;                      // warp-uniform
    float t2t = -INFINITY, ax = -INFINITY;
    if (r[i] >= 0) {
      float f[16];
      x[i].unpack(f);
      t2t = canonical_score<T>(f, static_cast<const T*>(a.queries), q0, q1, a.reduce, lane);
      if (AUX) {
        y[AUX ? i : 0].unpack(f);
        ax = canonical_score<T>(f, static_cast<const T*>(a.aux_queries), a.aux_class_begin[c], a.aux_class_begin[c + 1], a.aux_reduce, lane);
      }
    }
    if (lane == 0) {
      a.exact_scratch[base + j0 + i] = t2t;
      if (AUX) a.aux_scratch[base + j0 + i] = ax;
      // the frontier proof rests on |approximate - exact| <= eps: check it on every row we look at anyway
      if (a.eps_violation && !a.all_or_nothing && r[i] >= 0 && fabsf(t2t - approx[i]) > a.eps) *a.eps_violation = 1;
    }
  }
}

// Accept walk (add_to_split :439-482, add_t2t_ranked_t2i_tshd_to_split :507-527) over the re-scored candidates of one
// class: order by (exact score desc, row asc), accept rows with exact >= thr and aux >= aux_thr, stop at k.  A
// truncated list vouches only for rows above its frontier (approximate score of its last candidate + eps): anything
// the scan left out scores at most that.
__global__ void __launch_bounds__(kSelThreads) walk_kernel(const WalkArgs a) {
  __shared__ uint64_t s_keys[kSortCap];
  __shared__ uint16_t s_idx[kSortCap];
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_total;
  const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t n = static_cast<uint32_t>(min(a.cand_counts[c], a.stride));
  const size_t base = static_cast<size_t>(c) * a.stride;
  const bool trunc = a.truncated != nullptr && a.truncated[c] != 0;
  float frontier = -INFINITY;
  if (trunc) frontier = a.all_or_nothing ? INFINITY : (n > 0 ? a.cand_scores[base + n - 1] + a.eps : INFINITY);
  uint32_t P = 2;
  while (P < n) P <<= 1;
  for (uint32_t i = tid; i < P; i += kSelThreads) {
    uint64_t key = 0ull;
    if (i < n) {
      const float e = a.exact_scratch[base + i];
      if (e > -INFINITY) key = make_key(e + 0.0f, static_cast<uint32_t>(a.cand_rows[base + i] - a.key_row_base));
    }
    s_keys[i] = key;
    s_idx[i] = static_cast<uint16_t>(i);
  }
  __syncthreads();
  block_sort_desc<true>(s_keys, s_idx, P);
  constexpr int kPer = kSortCap / kSelThreads;   // 4 consecutive sorted positions per thread
  bool pass[kPer];
  uint32_t mine = 0;
#pragma unroll
  for (int i = 0; i < kPer; ++i) {
    const uint32_t p = tid * kPer + i;
    bool ok = false;
    if (p < n) {
      const uint64_t key = s_keys[p];
      const float e = key_score(key);
      ok = key != 0ull && e > frontier && e >= a.thr;
      if (ok && a.aux_bank) ok = a.aux_scratch[base + s_idx[p]] >= a.aux_thr;
    }
    pass[i] = ok;
    mine += ok ? 1u : 0u;
  }
  uint32_t inc = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = s_warp[lane], wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, wi, d);
      if (lane >= d) wi += t;
    }
    s_warp[lane] = wi - w;   // exclusive
  }
  __syncthreads();
  uint32_t pos = s_warp[warp] + inc - mine;
  if (tid == kSelThreads - 1) s_total = pos + mine;
  __syncthreads();
  const uint32_t total = s_total;
#pragma unroll
  for (int i = 0; i < kPer; ++i) {
    if (pass[i]) {
      if (pos < static_cast<uint32_t>(a.k)) {
        const uint32_t p = tid * kPer + i;
        const size_t src = base + s_idx[p];
        const size_t dst = static_cast<size_t>(c) * a.k + pos;
        a.out_scores[dst] = key_score(s_keys[p]);
        a.out_rows[dst] = a.cand_rows[src];
        if (a.out_aux) a.out_aux[dst] = a.aux_bank ? a.aux_scratch[src] : 0.0f;
      }
      ++pos;
    }
  }
  const uint32_t cnt = min(total, static_cast<uint32_t>(a.k));
  for (uint32_t i = cnt + tid; i < static_cast<uint32_t>(a.k); i += kSelThreads) {
    const size_t dst = static_cast<size_t>(c) * a.k + i;
    a.out_scores[dst] = 0.0f;
    a.out_rows[dst] = -1;
    if (a.out_aux) a.out_aux[dst] = 0.0f;
  }
  if (tid == 0) {
    a.out_counts[c] = static_cast<int32_t>(cnt);
    const bool complete = !trunc || total >= static_cast<uint32_t>(a.k);
    if (a.out_limit) a.out_limit[c] = complete ? -INFINITY : frontier;
    if (a.incomplete) a.incomplete[c] = complete ? 0 : 1;
    if (a.status) {
      a.status[1 + c] = complete ? 0 : 1;
      a.status[1 + a.n_classes + c] = static_cast<int32_t>(cnt);
      if (c == 0) a.status[0] = a.job_flags ? static_cast<int32_t>(*a.job_flags) : 0;
    }
  }
}

// canonical score of every row against the queries of ITS OWN class (partitioned data): one warp per row
template <typename T>
__global__ void __launch_bounds__(256) score_rows_kernel(const T* __restrict__ bank, const int32_t* __restrict__ row_class, int64_t n_rows,
                                                          const T* __restrict__ queries, const int32_t* __restrict__ class_begin,
                                                          int n_classes, int reduce, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (r >= n_rows) return;
  const int c = row_class[r];
  float s = -INFINITY;
  if (c >= 0 && c < n_classes) {
    float x[16];
    load16<T>(bank + r * kDim + lane * 16, x);
    s = canonical_score<T>(x, queries, class_begin[c], class_begin[c + 1], reduce, lane);
  }
  if (lane == 0) out[r] = s;
}

__global__ void remap_classes_kernel(const int32_t* __restrict__ in, const int32_t* __restrict__ map, int n_map, int64_t n,
                                     int32_t* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t c = in[i];
  out[i] = (c >= 0 && c < n_map) ? map[c] : -1;
}

// ---------------------------------------------------------------------------------- shard merge
// The gathered arrays are either contiguous ([G,C,k] / [G,C]) or slices of per-rank packed buffers laid end to
// end by the all-gather: shard g of every array then starts g * stride bytes after shard 0.
struct ShardView {
  int64_t big_f32, big_i64, small;   // byte strides between shards for [C,k] f32, [C,k] i64 and [C] i32 arrays
};
__host__ __device__ inline ShardView make_shard_view(int64_t stride_bytes, int C, int k) {
  ShardView v;
  v.big_f32 = stride_bytes ? stride_bytes : static_cast<int64_t>(C) * k * 4;
  v.big_i64 = stride_bytes ? stride_bytes : static_cast<int64_t>(C) * k * 8;
  v.small = stride_bytes ? stride_bytes : static_cast<int64_t>(C) * 4;
  return v;
}
template <typename T> __device__ __forceinline__ const T* shard_ptr(const T* base, int g, int64_t stride) {
  return reinterpret_cast<const T*>(reinterpret_cast<const char*>(base) + static_cast<int64_t>(g) * stride);
}

// keys laid out [C][G*k]; absent entries and entries failing the aux (T2I) predicate get key 0,
// which sorts below every real key
__global__ void merge_keys_kernel(const float* __restrict__ scores, const int64_t* __restrict__ rows,
                                  const float* __restrict__ aux, float aux_thr, const int32_t* __restrict__ counts,
                                  int G, ShardView sv, int C, int k, uint64_t* __restrict__ keys) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = static_cast<size_t>(G) * C * k;
  if (i >= total) return;
  const int j = static_cast<int>(i % k);
  const int c = static_cast<int>((i / k) % C);
  const int g = static_cast<int>(i / (static_cast<size_t>(k) * C));
  const size_t e = static_cast<size_t>(c) * k + j;
  uint64_t key = 0;
  if (j < shard_ptr(counts, g, sv.small)[c] && (aux == nullptr || shard_ptr(aux, g, sv.big_f32)[e] >= aux_thr))
    key = make_key(shard_ptr(scores, g, sv.big_f32)[e] + 0.0f, static_cast<uint32_t>(shard_ptr(rows, g, sv.big_i64)[e]));
  keys[(static_cast<size_t>(c) * G + g) * k + j] = key;
}

// Global accept walk over the shards' lists: the k_out best predicate-passing entries under (score desc, row asc).
// Shard g vouches only for rows scoring above limit[g][c] (-inf: its list is complete); the result is proven exact
// only if all of it stays above every shard's limit.
__global__ void __launch_bounds__(kSelThreads)
merge_kernel(const uint64_t* __restrict__ keys, const float* __restrict__ scores, const int64_t* __restrict__ rows,
             const float* __restrict__ aux, const int32_t* __restrict__ counts, const float* __restrict__ limit,
             int G, ShardView sv, int C, int k, int k_out, float* __restrict__ out_scores, int64_t* __restrict__ out_rows,
             float* __restrict__ out_aux, int32_t* __restrict__ out_counts, int32_t* __restrict__ incomplete) {
  __shared__ uint64_t s_keys[kSortCap];
  __shared__ uint32_t s_hist[256];
  __shared__ uint32_t s_misc[4];
  __shared__ uint32_t s_valid;
  __shared__ unsigned long long s_floor;
  const int c = blockIdx.x, tid = threadIdx.x;
  const uint32_t n = static_cast<uint32_t>(G) * k;
  const uint64_t* kc = keys + static_cast<size_t>(c) * n;
  if (tid == 0) { s_valid = 0; s_floor = ~0ull; }
  __syncthreads();
  // Lower bound of the k_out-th best key: the first ceil(k_out/G) entries of all shard lists together are at least
  // k_out keys, so the result cannot reach below the smallest of them (for sorted lists that is the weakest shard's
  // ceil(k_out/G)-th row).  An absent or predicate-failing entry among them (key 0) voids the bound.  Typically ~k_out
  // of the G*k keys stay and the block sort shrinks accordingly.
  const uint32_t per = (static_cast<uint32_t>(k_out) + G - 1) / G;
  if (per <= static_cast<uint32_t>(k)) {
    unsigned long long lo = ~0ull;
    for (uint32_t i = tid; i < static_cast<uint32_t>(G) * per; i += kSelThreads) lo = min(lo, static_cast<unsigned long long>(kc[static_cast<size_t>(i / per) * k + i % per]));
    if (lo != ~0ull) atomicMin(&s_floor, lo);
  } else if (tid == 0) {
    s_floor = 0ull;
  }
  __syncthreads();
  const FloorKeys fk{kc, s_floor != 0ull ? static_cast<uint64_t>(s_floor) : 1ull};
  uint32_t v = 0;
  for (uint32_t i = tid; i < n; i += kSelThreads) v += fk(i) != 0ull ? 1u : 0u;
  if (v) atomicAdd(&s_valid, v);
  __syncthreads();
  const uint32_t valid = s_valid;
  const uint32_t total = select_sorted(fk, n, valid, static_cast<uint32_t>(k_out), s_keys, s_hist, s_misc);
  const uint32_t cnt = min(static_cast<uint32_t>(k_out), total);
  for (uint32_t i = tid; i < static_cast<uint32_t>(k_out); i += kSelThreads) {
    const bool ok = i < cnt;
    const uint64_t key = ok ? s_keys[i] : 0ull;
    out_scores[static_cast<size_t>(c) * k_out + i] = ok ? key_score(key) : 0.0f;
    out_rows[static_cast<size_t>(c) * k_out + i] = ok ? static_cast<int64_t>(key_row(key)) : -1;
    if (out_aux && !ok) out_aux[static_cast<size_t>(c) * k_out + i] = 0.0f;
  }
  if (tid == 0) {
    out_counts[c] = static_cast<int32_t>(cnt);
    if (incomplete) {
      float lim = -INFINITY;
      if (limit)
        for (int g = 0; g < G; ++g) lim = fmaxf(lim, shard_ptr(limit, g, sv.small)[c]);
      incomplete[c] = (lim > -INFINITY && (cnt < static_cast<uint32_t>(k_out) || key_score(s_keys[cnt - 1]) <= lim)) ? 1 : 0;
    }
  }
  if (out_aux && aux) {
    // route each surviving entry's aux value (its T2I score) to its merged position: keys are
    // unique, binary-search the descending sorted list
    for (uint32_t i = tid; i < n; i += kSelThreads) {
      const uint64_t key = kc[i];
      if (key == 0ull) continue;
      uint32_t lo = 0, hi = cnt;
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (s_keys[mid] > key) lo = mid + 1; else hi = mid;
      }
      if (lo < cnt && s_keys[lo] == key) {
        const int g = static_cast<int>(i / k), j = static_cast<int>(i % k);
        out_aux[static_cast<size_t>(c) * k_out + lo] = shard_ptr(aux, g, sv.big_f32)[static_cast<size_t>(c) * k + j];
      }
    }
  }
}

// ---------------------------------------------------------------------------------- near duplicates
// remove_near_duplicates2 (sample_retrieval.py:237-275): inside a class, row j is a duplicate iff some
// EARLIER row i of the class has <x_i, x_j> > threshold (np.triu(sim, k=1) > 0.9, j indices).  Rows are
// addressed through `order` (bank rows sorted by class, file order kept inside a class).  One CTA per
// (class, block of 32 j rows): 32 x 32 threads, thread (ty, tx) owns the pair (i-block row ty, j row tx).
template <typename T>
__global__ void __launch_bounds__(1024)
near_dup_kernel(const T* __restrict__ bank, const int64_t* __restrict__ order, const int32_t* __restrict__ class_start,
                float threshold, uint8_t* __restrict__ dup) {
  __shared__ float sA[32][33], sB[32][33];
  const int c = blockIdx.y;
  const int s = class_start[c], e = class_start[c + 1];
  const int jb = blockIdx.x;
  if (s + jb * 32 >= e) return;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int j = s + jb * 32 + tx;
  const int jl = s + jb * 32 + ty;                       // the j row this thread stages
  const int64_t rj = jl < e ? order[jl] : -1;
  bool flag = false;
  for (int ib = 0; ib <= jb; ++ib) {
    const int i = s + ib * 32 + ty;
    const int64_t ri = i < e ? order[i] : -1;
    float dot = 0.0f;
    for (int k0 = 0; k0 < kDim; k0 += 32) {
      sA[ty][tx] = ri >= 0 ? static_cast<float>(bank[ri * kDim + k0 + tx]) : 0.0f;
      sB[ty][tx] = rj >= 0 ? static_cast<float>(bank[rj * kDim + k0 + tx]) : 0.0f;
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < 32; ++kk) dot = fmaf(sA[ty][kk], sB[tx][kk], dot);
      __syncthreads();
    }
    if (i < j && i < e && j < e && dot > threshold) flag = true;
  }
  if (flag) dup[j] = 1;      // every writer stores the same value
}

}  // namespace

cudaError_t launch_near_dup(const void* bank, int dtype, const int64_t* d_order, const int32_t* d_class_start, int n_classes,
                            int max_class_rows, float threshold, uint8_t* d_dup, cudaStream_t stream) {
  if (n_classes <= 0 || max_class_rows <= 0) return cudaSuccess;
  const dim3 grid((max_class_rows + 31) / 32, n_classes), block(32, 32);
  if (dtype == 0) near_dup_kernel<__nv_bfloat16><<<grid, block, 0, stream>>>(static_cast<const __nv_bfloat16*>(bank), d_order, d_class_start, threshold, d_dup);
  else near_dup_kernel<float><<<grid, block, 0, stream>>>(static_cast<const float*>(bank), d_order, d_class_start, threshold, d_dup);
  return cudaGetLastError();
}

cudaError_t launch_select(const JobState& st, int n_classes, int64_t row_offset, float* d_scores, int64_t* d_rows,
                          int32_t* d_counts, int32_t* d_truncated, cudaStream_t stream, uint32_t band_k, float band, int32_t* zero_word) {
  if (n_classes <= 0) return cudaSuccess;
  final_tau_kernel<<<(n_classes + 7) / 8, 256, 0, stream>>>(st, n_classes);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const size_t part_smem = static_cast<size_t>(n_classes) * 8;
  if (part_smem <= 160 * 1024) {
    static bool attr_set = false;
    if (!attr_set) {
      e = cudaFuncSetAttribute(partition_agg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
      if (e != cudaSuccess) return e;
      attr_set = true;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    partition_agg_kernel<<<sms, 1024, part_smem, stream>>>(st, n_classes);
  } else {
    partition_kernel<<<dim3(st.n_lists, 4), 256, 0, stream>>>(st);
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  select_kernel<<<n_classes, kSelThreads, 0, stream>>>(st, row_offset, d_scores, d_rows, d_counts, d_truncated, d_truncated ? band_k : 0u, band, zero_word);
  return cudaGetLastError();
}

cudaError_t launch_bootstrap(const JobState& st, int n_classes, const float* d_scores_t, uint32_t n_prefix, uint32_t row_base,
                             uint32_t first_spare_list, cudaStream_t stream) {
  bootstrap_kernel<<<n_classes, kSelThreads, 0, stream>>>(st, d_scores_t, n_prefix, row_base, first_spare_list);
  return cudaGetLastError();
}

cudaError_t launch_job_reset(const JobState& st, int n_classes, cudaStream_t stream) {
  const size_t n = std::max(static_cast<size_t>(n_classes) * kHistBins, static_cast<size_t>(st.n_lists));
  reset_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(st, n_classes);
  return cudaGetLastError();
}

cudaError_t launch_rescore_walk(const WalkArgs& a, cudaStream_t stream) {
  constexpr int kBf16Cands = 4, kF32Cands = 2;          // candidates per warp with a predicate bank; twice that without
  const bool aux = a.aux_bank != nullptr;
  const int per = (a.dtype == 0 ? kBf16Cands : kF32Cands) * (aux ? 1 : 2);
  const int64_t n = static_cast<int64_t>(a.n_classes) * ((a.stride + per - 1) / per);
  if (n <= 0) return cudaSuccess;
  const unsigned grid = static_cast<unsigned>((n + 7) / 8);
  if (a.dtype == 0) {
    if (aux) rescore_kernel<__nv_bfloat16, kBf16Cands, true><<<grid, 256, 0, stream>>>(a);
    else rescore_kernel<__nv_bfloat16, 2 * kBf16Cands, false><<<grid, 256, 0, stream>>>(a);
  } else {
    if (aux) rescore_kernel<float, kF32Cands, true><<<grid, 256, 0, stream>>>(a);
    else rescore_kernel<float, 2 * kF32Cands, false><<<grid, 256, 0, stream>>>(a);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  walk_kernel<<<a.n_classes, kSelThreads, 0, stream>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_score_rows(const void* bank, int dtype, const int32_t* row_class, int64_t n_rows, const void* queries,
                              const int32_t* class_begin, int n_classes, int reduce, float* out, cudaStream_t stream) {
  if (n_rows <= 0) return cudaSuccess;
  const unsigned grid = static_cast<unsigned>((n_rows + 7) / 8);
  if (dtype == 0)
    score_rows_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(bank), row_class, n_rows,
                                                              static_cast<const __nv_bfloat16*>(queries), class_begin, n_classes, reduce, out);
  else
    score_rows_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(bank), row_class, n_rows,
                                                        static_cast<const float*>(queries), class_begin, n_classes, reduce, out);
  return cudaGetLastError();
}

// out[dst_cls[i]] <- src[src_idx[i]] for the [.,k] result arrays of n classes (splicing escalated classes into the result)
__global__ void __launch_bounds__(256) splice_kernel(const int32_t* __restrict__ dst_cls, const int32_t* __restrict__ src_idx, int k,
                                                     const float* __restrict__ s_scores, const int64_t* __restrict__ s_rows,
                                                     const float* __restrict__ s_aux, const int32_t* __restrict__ s_counts,
                                                     float* __restrict__ d_scores, int64_t* __restrict__ d_rows, float* __restrict__ d_aux,
                                                     int32_t* __restrict__ d_counts) {
  const size_t d = static_cast<size_t>(dst_cls[blockIdx.x]) * k, s = static_cast<size_t>(src_idx[blockIdx.x]) * k;
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    d_scores[d + j] = s_scores[s + j];
    d_rows[d + j] = s_rows[s + j];
    if (d_aux && s_aux) d_aux[d + j] = s_aux[s + j];
  }
  if (threadIdx.x == 0) d_counts[dst_cls[blockIdx.x]] = s_counts[src_idx[blockIdx.x]];
}

cudaError_t launch_splice(const int32_t* d_dst_cls, const int32_t* d_src_idx, int n, int k, const float* s_scores, const int64_t* s_rows,
                          const float* s_aux, const int32_t* s_counts, float* d_scores, int64_t* d_rows, float* d_aux, int32_t* d_counts,
                          cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  splice_kernel<<<n, 256, 0, stream>>>(d_dst_cls, d_src_idx, k, s_scores, s_rows, s_aux, s_counts, d_scores, d_rows, d_aux, d_counts);
  return cudaGetLastError();
}

cudaError_t launch_remap_classes(const int32_t* in, const int32_t* map, int n_map, int64_t n, int32_t* out, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  remap_classes_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(in, map, n_map, n, out);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------- zero-shot prediction
// argmax over the class scores of a row, lowest class index on ties (torch.argmax on the logits of the zero-shot head,
// zeroshot_clip_img_filter sample_retrieval.py:299-301).  One warp per row.
__global__ void __launch_bounds__(256) argmax_rows_kernel(const float* __restrict__ scores, int64_t n_rows, int C, int32_t* __restrict__ pred) {
  const int lane = threadIdx.x & 31;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (r >= n_rows) return;
  const float* s = scores + r * C;
  float best = -INFINITY;
  int arg = 0x7fffffff;
  for (int c = lane; c < C; c += 32) {
    const float v = s[c];
    if (v > best || (v == best && c < arg)) { best = v; arg = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
  }
  if (lane == 0) pred[r] = arg == 0x7fffffff ? 0 : arg;
}

cudaError_t launch_argmax_rows(const float* d_scores, int64_t n_rows, int n_classes, int32_t* d_pred, cudaStream_t stream) {
  if (n_rows <= 0) return cudaSuccess;
  argmax_rows_kernel<<<static_cast<unsigned>((n_rows + 7) / 8), 256, 0, stream>>>(d_scores, n_rows, n_classes, d_pred);
  return cudaGetLastError();
}

cudaError_t launch_merge(const float* d_scores, const int64_t* d_rows, const float* d_aux, float aux_thr,
                         const int32_t* d_counts, const float* d_limit, int n_shards, int64_t shard_stride_bytes, int n_classes,
                         int k, int k_out, uint64_t* d_key_scratch, float* d_out_scores, int64_t* d_out_rows, float* d_out_aux,
                         int32_t* d_out_counts, int32_t* d_incomplete, cudaStream_t stream) {
  const size_t total = static_cast<size_t>(n_shards) * n_classes * k;
  if (total == 0) return cudaSuccess;
  const ShardView sv = make_shard_view(shard_stride_bytes, n_classes, k);
  merge_keys_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(d_scores, d_rows, d_aux, aux_thr, d_counts,
                                                                                  n_shards, sv, n_classes, k, d_key_scratch);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  merge_kernel<<<n_classes, kSelThreads, 0, stream>>>(d_key_scratch, d_scores, d_rows, d_aux, d_counts, d_limit, n_shards, sv,
                                                      n_classes, k, k_out, d_out_scores, d_out_rows, d_out_aux, d_out_counts,
                                                      d_incomplete);
  return cudaGetLastError();
}

}  // namespace swat
