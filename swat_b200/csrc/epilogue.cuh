// Fused selection epilogue shared by the tcgen05 and the SIMT scan kernels.
//
// A thread owns ONE bank row and receives that row's scores against NC consecutive query columns in
// registers.  Columns of one class are adjacent; the last column of a class carries the group size
// and the class threshold.  The common case -- no lane of the warp beats the class threshold -- costs
// a compare, a vote and an untaken branch per class.  Survivors (rare after warm-up: about
// k*ln(N/k) per class over the whole scan) take the slow path: a fire-and-forget histogram RED and a
// plain store into a survivor list.  Nothing on this path waits for an atomic to return, so the
// epilogue never stalls the TMEM pipeline behind L2 round trips.  Class thresholds are recomputed
// from the histograms by a dedicated refresher warp (tcgen05 kernel) or by each CTA on entry (SIMT
// kernel).  The N x Q score matrix never reaches HBM.
//
// Replaces, for all classes at once: t2t_similarity -> sorted() -> add_to_split of
// /root/reference/retrieval/sample_retrieval.py:752-758 (and :804-812 with the in-pass T2I predicate).
#pragma once
#include "common.cuh"

namespace swat {

struct EpiCtx {
  float* tau_col;          // smem, thresholds by block-local column (valid at group-end columns)
  const int32_t* cls_col;  // smem, global class index by block-local column (-1 = padding)
  const float* cnt_col;    // smem, group size by block-local column (at group-end columns)
  uint32_t row;            // view-local row owned by this thread
  bool row_valid;
  int my_cls;              // partitioned mode: class of this row (-1 = none)
  float acc, acc2;         // running group reduce (T2T, in-pass T2I)
  uint32_t list_pos;       // warp-private lists: next free slot (kept in a register across the kernel)
};

// Everything the slow path needs, held in registers and passed BY VALUE: the noinline slow path must
// not chase pointers into the kernel-parameter struct (each such load is a long-scoreboard stall).
struct SlowCtx {
  uint4* list_base;           // this warp's survivor list (ATOMIC_LIST: the shared list of this CTA slot)
  uint32_t* list_count;       // ATOMIC_LIST only: slot counter of that list
  uint32_t* hist;             // [C * kHistBins]
  uint32_t* flags;
  uint32_t list_cap;
  uint32_t row_base;
  float hist_lo, hist_scale;
  const uint32_t* pass_bits;  // nullable: per-class bitmap of rows that pass the predicate (see ScanArgs)
  int64_t bits_words;
};
__device__ __forceinline__ SlowCtx make_slow_ctx(const ScanArgs& a, uint32_t list_id) {
  SlowCtx s;
  s.list_base = a.st.list + static_cast<size_t>(list_id) * a.st.list_cap;
  s.list_count = a.st.list_count + list_id;
  s.hist = a.st.hist;
  s.flags = a.st.flags;
  s.list_cap = a.st.list_cap;
  s.row_base = a.row_base;
  s.hist_lo = a.st.hist_lo;
  s.hist_scale = a.st.hist_scale;
  s.pass_bits = a.pass_bits;
  s.bits_words = a.bits_words;
  return s;
}

// Recompute one class threshold from its histogram (whole warp): the largest bin edge with at least
// k_fetch appended scores at or above it.  Valid at any time: every counted score belongs to a
// distinct bank row, so at least k_fetch rows score >= the edge and no row below it can enter the
// top k_fetch.
static __device__ __noinline__ void refresh_tau(const JobState& st, int cls) {
  const int lane = threadIdx.x & 31;
  const uint32_t* h = st.hist + static_cast<size_t>(cls) * kHistBins + lane * 32;
  uint32_t v[32];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    uint4 t = ld_cg_u32x4(h + 4 * i);
    v[4 * i + 0] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
  uint32_t mine = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) mine += v[i];
  uint32_t suf = mine;  // inclusive suffix sum over lanes (higher lane = higher scores)
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_down_sync(0xffffffffu, suf, d);
    if (lane + d < 32) suf += t;
  }
  const uint32_t K = class_k(st, cls);
  const uint32_t ball = __ballot_sync(0xffffffffu, suf >= K);
  if (ball == 0) return;
  const int L = 31 - __clz(ball);
  if (lane == L) {
    int bin = -1;
    uint32_t run = suf - mine;
#pragma unroll
    for (int j = 31; j >= 0; --j) {
      if (bin < 0) {
        run += v[j];
        if (run >= K) bin = lane * 32 + j;
      }
    }
    if (bin >= 1) atomicMax(&st.tau_enc[cls], f32_enc(hist_edge(st, bin)));
  }
}

// Row-level part of the accept predicate (exclusion bitmap: near-duplicates, rows already taken by an earlier
// sampler, sample_retrieval.py:452-462).  It does not depend on the column, so it is folded into row_valid once
// per row instead of being tested per survivor.
__device__ __forceinline__ bool row_excluded(const uint32_t* exclude, uint32_t row) {
  return exclude != nullptr && ((exclude[row >> 5] >> (row & 31)) & 1u) != 0u;
}

// v[j] for a run-time j without spilling the register array: a jump table of 32 moves.
template <int NC> __device__ __forceinline__ float pick_column(const float (&v)[NC], int j) {
  static_assert(NC == 32, "chunks are 32 columns wide");
  float r = 0.0f;
  switch (j) {
#define SWAT_PICK(k) case k: r = v[k]; break;
    SWAT_PICK(0) SWAT_PICK(1) SWAT_PICK(2) SWAT_PICK(3) SWAT_PICK(4) SWAT_PICK(5) SWAT_PICK(6) SWAT_PICK(7)
    SWAT_PICK(8) SWAT_PICK(9) SWAT_PICK(10) SWAT_PICK(11) SWAT_PICK(12) SWAT_PICK(13) SWAT_PICK(14) SWAT_PICK(15)
    SWAT_PICK(16) SWAT_PICK(17) SWAT_PICK(18) SWAT_PICK(19) SWAT_PICK(20) SWAT_PICK(21) SWAT_PICK(22) SWAT_PICK(23)
    SWAT_PICK(24) SWAT_PICK(25) SWAT_PICK(26) SWAT_PICK(27) SWAT_PICK(28) SWAT_PICK(29) SWAT_PICK(30) SWAT_PICK(31)
#undef SWAT_PICK
  }
  return r;
}

// Slow path.  mask bit j: this lane's row passed the fast predicate at column col0 + j (v[j] holds the class value).
// With ~k'(1 + ln(N/n0)) survivors per class a 32 x 32 chunk holds one with probability 0.3-0.5 at the benchmark
// sizes, so this runs for every second or third chunk and has to be short: no per-column branches, no call.  Each
// round appends the lowest remaining survivor of EVERY lane (lane-parallel; one round in the usual case): slots
// from a ballot, a 16-byte store into the warp's list, a fire-and-forget histogram RED.
// ATOMIC_LIST: lists are shared between warps (SIMT kernel) and slots are reserved with an atomic;
// otherwise the list is private to this warp and the position lives in a register.
template <int NC, int RED, bool ATOMIC_LIST>
__device__ __forceinline__ void drain_survivors(const SlowCtx& sc, EpiCtx& cx, const float (&v)[NC], uint32_t mask, int col0) {
  const uint32_t lane = threadIdx.x & 31u;
  uint32_t ballot = __ballot_sync(0xffffffffu, mask != 0u);
  while (ballot != 0u) {
    // every lane with work takes its lowest remaining survivor; with a predicate bitmap the survivor must also have
    // its (class, row) bit set -- tested here, on the rare path, so the fast path never sees the bitmap
    int j = 0, cls = 0;
    bool keep = mask != 0u;
    if (keep) {
      j = __ffs(mask) - 1;
      mask &= mask - 1u;
      cls = cx.cls_col[col0 + j];
      if (sc.pass_bits != nullptr) {
        const uint32_t r = sc.row_base + cx.row;
        keep = ((sc.pass_bits[static_cast<size_t>(cls) * sc.bits_words + (r >> 5)] >> (r & 31u)) & 1u) != 0u;
      }
    }
    const uint32_t kept = __ballot_sync(0xffffffffu, keep);
    const uint32_t n = __popc(kept);
    uint32_t base = cx.list_pos;
    if (n != 0u) {
      if (ATOMIC_LIST) {
        if (lane == static_cast<uint32_t>(__ffs(kept) - 1)) base = atomicAdd(sc.list_count, n);
        base = __shfl_sync(0xffffffffu, base, __ffs(kept) - 1);
      } else {
        cx.list_pos += n;
      }
    }
    if (keep) {
      const float s = red_fin<RED>(pick_column<NC>(v, j), cx.cnt_col[col0 + j]) + 0.0f;   // -0.0 -> +0.0: Python compares them equal, the key must too
      const uint32_t slot = base + __popc(kept & ((1u << lane) - 1u));
      if (slot < sc.list_cap) {
        const uint64_t key = make_key(s, sc.row_base + cx.row);
        sc.list_base[slot] = make_uint4(static_cast<uint32_t>(key), static_cast<uint32_t>(key >> 32), static_cast<uint32_t>(cls), 0u);
      } else {
        atomicOr(sc.flags, 2u);
      }
      int b = static_cast<int>((s - sc.hist_lo) * sc.hist_scale);
      b = b < 0 ? 0 : (b > kHistBins - 1 ? kHistBins - 1 : b);
      // fire-and-forget reduction (no return value, nothing waits on it)
      asm volatile("red.global.add.u32 [%0], 1;" :: "l"(sc.hist + static_cast<size_t>(cls) * kHistBins + b) : "memory");
    }
    ballot = __ballot_sync(0xffffffffu, mask != 0u);
  }
}

// Threshold the fast path compares the *accumulated* value of a class against.  For MEAN the
// accumulator is the sum, so the class threshold is scaled by the group size and loosened by a few ulp:
// the fast filter may only ever let slightly MORE rows through (the survivor lists are supersets; the
// exact value sum/R is what gets stored and the partition step filters on it exactly).
template <int RED> __device__ __forceinline__ float fast_tau(float tau, float cnt) {
  // one-query-per-class mode tests the SIGN of (score - tau): a threshold of +0.0 becomes -0.0 so that a
  // score of -0.0 still passes, as it does in the reference (Python: -0.0 >= 0.0)
  if (RED != RED_MEAN) return tau == 0.0f ? -0.0f : tau;
  if (!(cnt > 0.0f) || !isfinite(tau)) return tau;
  const float t = tau * cnt;
  return t - fabsf(t) * 1.0e-6f - 1.0e-30f;
}

// NC columns starting at block-local column col0.  endmask bit j: column col0+j closes a class.
//
// Branch-free for every reduce mode: the running reduce is carried along the columns (reset after a
// closing column by a select on the warp-uniform endmask bit), each column's value is compared with
// the per-column threshold table -- NaN everywhere except at closing columns -- and the results
// fold into one pass mask; ONE vote decides whether anybody in the warp has a survivor in the chunk.
template <int NC, int RED, bool PART, bool DUAL, bool DENSE, bool ATOMIC_LIST>
__device__ __forceinline__ void process_chunk(const ScanArgs& a, const SlowCtx& sc, EpiCtx& cx, float (&v)[NC],
                                              float (&v2)[NC], int col0, uint32_t endmask) {
  const float* tau = cx.tau_col + col0;
  const int32_t* cls = cx.cls_col + col0;
  const float* cnt = cx.cnt_col + col0;
  if (DENSE) {
    if (RED == RED_NONE && a.bits_out == nullptr && a.dense_transposed) {
      // Threshold-bootstrap prefix (scores_t[class][row]): one query per class, so the real columns of a chunk are a
      // prefix and their classes are consecutive -- a pointer bump and a coalesced store per column.  The generic
      // loop below (class table lookups, a branch and a 64-bit multiply per column, ~27 instructions) ran mostly out
      // of a cold instruction cache in this two-tile kernel: 15 us per tile against 3 us for the selecting epilogue.
      const int n_real = __popc(endmask);
      if (n_real > 0 && cx.row_valid) {
        float* o = a.dense_out + static_cast<size_t>(cls[0]) * a.dense_ld + cx.row;
#pragma unroll
        for (int j = 0; j < NC; ++j) {
          if (j < n_real) *o = v[j];
          o += a.dense_ld;
        }
      }
      return;
    }
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      cx.acc = (RED == RED_NONE) ? v[j] : red_op<RED>(cx.acc, v[j]);
      if ((endmask >> j) & 1u) {  // warp-uniform
        if (a.bits_out != nullptr) {
          // a warp owns 32 consecutive rows starting at a multiple of 32: one word per (class, warp)
          const uint32_t word = __ballot_sync(0xffffffffu, cx.row_valid && red_fin<RED>(cx.acc, cnt[j]) >= a.bits_thr);
          const uint32_t r = a.row_base + cx.row;
          if ((threadIdx.x & 31) == 0 && static_cast<int64_t>(r >> 5) < a.bits_words)
            a.bits_out[static_cast<size_t>(cls[j]) * a.bits_words + (r >> 5)] = word;
        } else if (cx.row_valid) {
          a.dense_out[a.dense_transposed ? static_cast<size_t>(cls[j]) * a.dense_ld + cx.row : static_cast<size_t>(cx.row) * a.dense_ld + cls[j]] =
              red_fin<RED>(cx.acc, cnt[j]);
        }
        cx.acc = red_init<RED>();
      }
    }
    return;
  }
  if (!DUAL && !PART && NC == 32) {
    // Leanest form: d = value - tau, sign bits collected with a funnel shift -- two instructions per
    // column plus the running reduce.  tau is +inf at padding / non-closing columns (d = -inf: fails);
    // scores and reduce identities are finite, so no NaN can reach the sign test.
    uint32_t fail = 0;
#pragma unroll
    for (int j4 = 0; j4 < NC / 4; ++j4) {
      const float4 t = *reinterpret_cast<const float4*>(tau + 4 * j4);
      const float tt[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int j = 4 * j4 + i;
        if (RED != RED_NONE) {
          cx.acc = red_op<RED>(cx.acc, v[j]);
          v[j] = cx.acc;                                   // value of the class that closes here (if any)
          cx.acc = ((endmask >> j) & 1u) ? red_init<RED>() : cx.acc;
        }
        fail = __funnelshift_l(__float_as_uint(v[j] - tt[i]), fail, 1);
      }
    }
    const uint32_t mask = cx.row_valid ? ~__brev(fail) : 0u;
    drain_survivors<NC, RED, ATOMIC_LIST>(sc, cx, v, mask, col0);
    return;
  }
  uint32_t mask = 0;
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    if (RED != RED_NONE) {
      const bool end = (endmask >> j) & 1u;
      cx.acc = red_op<RED>(cx.acc, v[j]);
      v[j] = cx.acc;                                   // value of the class that closes here (if any)
      cx.acc = end ? red_init<RED>() : cx.acc;
      if (DUAL) {
        cx.acc2 = red_op<RED>(cx.acc2, v2[j]);
        v2[j] = cx.acc2;
        cx.acc2 = end ? red_init<RED>() : cx.acc2;
      }
    }
    bool p = cx.row_valid && (v[j] >= tau[j]);
    if (DUAL) p = p && (red_fin<RED>(v2[j], cnt[j]) >= a.t2i_thr);
    if (PART) p = p && (cx.my_cls == cls[j]);
    mask |= static_cast<uint32_t>(p) << j;
  }
  drain_survivors<NC, RED, ATOMIC_LIST>(sc, cx, v, mask, col0);
}

}  // namespace swat
