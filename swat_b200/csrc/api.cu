// C-ABI of swat_b200 (include/swat_b200.h): contexts, query sets, streaming jobs, and the two
// whole-pipeline entry points (HBM-resident banks, host banks).
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <numeric>
#include <vector>

#include "../../include/swat_b200.h"
#include "common.cuh"
#include "scan_tc.h"

using namespace swat;

namespace {

thread_local std::string g_err;

int32_t fail(int32_t code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

}  // namespace

namespace swat {
// error reporting / context access for the other translation units of the library (loader.cu)
int32_t api_fail(int32_t code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
}  // namespace swat

namespace {

#define CU_OK(expr)                                                                                     \
  do {                                                                                                  \
    cudaError_t e__ = (expr);                                                                           \
    if (e__ != cudaSuccess)                                                                             \
      return fail(SWAT_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)
#define SW_OK(expr)              \
  do {                           \
    int32_t r__ = (expr);        \
    if (r__ != SWAT_OK) return r__; \
  } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

uint16_t f32_to_bf16_rne(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return static_cast<uint16_t>((u >> 16) | 0x40u);
  return static_cast<uint16_t>((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
}

struct DevBuf {   // grow-only device workspace
  void* p = nullptr;
  size_t cap = 0;
  int32_t ensure(size_t bytes) {
    if (bytes <= cap) return SWAT_OK;
    const auto t0 = std::chrono::steady_clock::now();
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    CU_OK(cudaMalloc(&p, bytes));
    cap = bytes;
    if (getenv("SWAT_DEBUG")) fprintf(stderr, "[swat] workspace grows to %zu bytes (%.2f ms)\n", bytes,
                                      std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    return SWAT_OK;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <typename T> T* as() const { return static_cast<T*>(p); }
};

}  // namespace

struct swat_ctx {
  int device = 0;
  int sm_count = 0;
  size_t smem_optin = 0;
  EncodeTiledFn encode = nullptr;
  // options
  int cta_group = 2;
  int max_ctas = 0;
  int64_t cand_cap = 0;       // per-class candidate capacity after partition, 0 = auto
  int64_t list_entries = 0;   // total survivor-list entries, 0 = auto
  int overfetch = 0;          // 0 = auto
  int64_t host_chunk_rows = 1 << 18;
  bool unit_plan = true;            // several Q blocks: balanced (Q block x tile range) units over several launches
  int64_t bootstrap_rows = 32768;   // dense prefix used to seed thresholds of a fresh job (0 = off)
  // stats
  int64_t launches = 0;
  double timing[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  // workspaces for the whole-pipeline calls
  DevBuf w_scores, w_rows, w_counts, w_trunc, w_exact, w_aux, w_incomplete, w_keys, w_stage[3], w_rc[3], w_ex[3], w_img, w_idx;
  DevBuf w_out_scores, w_out_rows, w_out_t2i, w_out_counts, w_boot;
  DevBuf w_swap[10];                // bank-swap escalation pass: two re-score stages
  // several Q blocks: pairs sharing a tile range stay within this many tiles of each other.  0 = off, > 0 = fixed,
  // -1 (default) = automatic: 4 for 4-8 Q blocks while the GPU is power-capped (SM clock observed by the previous such
  // scan below 0.70 of the maximum; released above 0.80), else off -- held in step the pairs read the bank from HBM
  // once instead of 1.7x: neutral on one busy GPU, +20 % on a box with all eight busy
  int lock_window = -1;
  bool lock_auto_on = false;
  int clock_khz = 0;                // maximum SM clock
  unsigned long long* h_probe = nullptr;   // mapped pinned word: SM clock (MHz) observed by the last probed scan launch
  unsigned long long* d_probe = nullptr;
  DevBuf w_progress, w_bits, w_splice, w_tiles;
  bool dyn_tiles = true;            // one Q block: dynamic tile scheduling (pairs finish 3-5 % apart under a static split)
  int f32_op_stages = 3;            // fp32 banks: bf16 operand stages (the rest of the shared memory stages fp32 boxes); before swat_queries_create
  bool zero_copy = true;            // host pipeline: read candidates' rows from pinned host banks in place
  bool swap_pass = true;            // classes with fewer than k rows passing T2I: enumerate the passers from the image bank
  cudaStream_t copy_stream = nullptr, work_stream = nullptr;
  cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_copied[3] = {nullptr, nullptr, nullptr}, ev_used[3] = {nullptr, nullptr, nullptr};
  void* h_pinned = nullptr;
  size_t h_pinned_cap = 0;
  int32_t* h_status = nullptr;      // pinned: [0] job overflow word, [1..C] incomplete flags (one read-back per step)
  size_t h_status_cap = 0;
  swat_job* cached_job = nullptr;   // job buffers are reused across whole-pipeline calls
  // last sub-query set built for a targeted escalation (repeated calls hit the same classes)
  // sub-query sets of the classes being escalated, kept across calls (two slots: deeper over-fetch / bank-swap pass)
  swat_queries* esc_q[2] = {nullptr, nullptr};
  const swat_queries* esc_parent[2] = {nullptr, nullptr};
  std::vector<int> esc_classes[2];
  int esc_next = 0;
  DevBuf e_bufs[4][4];   // per escalation depth: scores, rows, t2i, counts of the sub-run
  DevBuf e_remap[4][2];  // per escalation depth, partitioned data: class map and the renumbered row_class of the sub-run
};

struct swat_queries {
  swat_ctx* ctx = nullptr;
  int Q = 0, C = 0, reduce = 0;
  std::vector<int32_t> class_begin;   // [C+1]
  // unpadded (T2I re-score) and padded-by-Q-block (scan kernels) device copies
  float* d_q_f32 = nullptr; uint16_t* d_q_bf16 = nullptr; int32_t* d_class_begin = nullptr;
  float* d_qp_f32 = nullptr; uint16_t* d_qp_bf16 = nullptr; int32_t* d_col_class = nullptr; float* d_col_count = nullptr;
  int32_t* d_blk_class = nullptr;     // [n_qb+1] first class of each Q block
  int32_t* d_blk_split = nullptr;     // [n_qb] grouped reduces: first column of the second epilogue warp set
  void* d_arena = nullptr;            // one allocation backs every device array above
  std::vector<float> h_q;             // host copy (sub-query sets for targeted escalation)
  mutable std::vector<int32_t> kclass_hint;   // per class: deepest over-fetch its T2I walk has needed so far (0 = default)
  mutable int32_t last_k_fetch = 0;   // over-fetch at which the last pipeline run completed
  mutable bool all_few_hint = false;  // every class had too few T2I passers for a T2T-ordered walk: start with the bank-swap pass
  int ctas = 2, n_qb = 1, n_blk = 16, n_cols = 16, n_stages = 0;
  int n_fstages = 0, n_opstages_f32 = 0;   // fp32 banks on the tcgen05 engine: staged fp32 boxes / bf16 operand stages
  // |score of bf16-rounded row and bf16-rounded queries - fp32 score| <= eps_conv for every L2-normalised row (see swat_queries_create)
  float eps_conv = 0.0f;
  CUtensorMap tm_q;
};

struct swat_job {
  swat_ctx* ctx = nullptr;
  const swat_queries* q = nullptr;
  JobState st;
  int n_classes_alloc = 0;
  bool fresh = true;                  // no rows folded in since the last reset
  uint32_t* d_k_class = nullptr;      // [n_classes_alloc] per-class k_fetch, used when the depths differ
  std::vector<uint32_t> h_k_class;    // what swat_job_set_class_depth last uploaded there (empty = unknown)
  cudaStream_t last_stream = nullptr;
};

namespace swat {
int api_ctx_device(const swat_ctx* ctx) { return ctx->device; }
}  // namespace swat

namespace {

// 2-D map of a row-major [rows, 512] array: 128-byte boxes (64 bf16 or 32 fp32 along k) x box_rows, SWIZZLE_128B
int32_t encode_2d(swat_ctx* ctx, CUtensorMap* tm, const void* ptr, uint64_t rows, uint32_t box_rows, int dtype) {
  const bool f32 = dtype == SWAT_F32;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(kDim), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(kDim) * (f32 ? 4 : 2)};
  cuuint32_t box[2] = {f32 ? 32u : 64u, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = ctx->encode(tm, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims,
                           strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(SWAT_ERR_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d, rows %llu, box %u, dtype %d)", (int)r,
                                     (unsigned long long)rows, box_rows, dtype);
  return SWAT_OK;
}
int32_t encode_2d_bf16(swat_ctx* ctx, CUtensorMap* tm, const void* ptr, uint64_t rows, uint32_t box_rows) {
  return encode_2d(ctx, tm, ptr, rows, box_rows, SWAT_BF16);
}

// Split the Q query columns into n_qb blocks at class boundaries, every block padded to the same
// n_blk (multiple of 16).  max_cols = what one CTA (pair) can keep resident.
bool plan_blocks(const std::vector<int32_t>& class_begin, int max_cols, int& n_qb, int& n_blk, std::vector<int>& blk_first_class) {
  const int C = static_cast<int>(class_begin.size()) - 1;
  const int Q = class_begin[C];
  for (n_qb = std::max(1, (Q + max_cols - 1) / max_cols); n_qb <= std::max(1, C); ++n_qb) {
    blk_first_class.assign(1, 0);
    int widest = 0, start_col = 0;
    bool ok = true;
    for (int b = 0, c = 0; b < n_qb; ++b) {
      const int64_t target = (static_cast<int64_t>(b + 1) * Q + n_qb - 1) / n_qb;   // cumulative column target
      int c_end = c;
      while (c_end < C && (class_begin[c_end + 1] <= target || c_end == c)) ++c_end;
      if (b == n_qb - 1) c_end = C;
      const int cols = class_begin[c_end] - start_col;
      if (cols > max_cols) { ok = false; break; }
      widest = std::max(widest, cols);
      start_col = class_begin[c_end];
      c = c_end;
      blk_first_class.push_back(c);
    }
    if (ok) {
      n_blk = std::max(16, (widest + 15) / 16 * 16);
      return true;
    }
  }
  return false;
}

// Bound on |score a scan kernel ranks a row by - the canonical fp32 score of select.cu|.  Full-precision engines
// (bf16 banks on the tensor cores: exact products, fp32 accumulation; the fp32-FMA kernel) differ from the canonical
// summation order only: 512 terms x 2^-23 x sum|x_i q_i| <= 6.1e-5 for unit vectors, observed ~2e-6.
constexpr float kEpsAccum = 1.0e-4f;

int32_t resolve_engine(const swat_queries* q, int32_t dtype, bool dual) {
  if (dual) return SWAT_ENGINE_SIMT;                 // in-pass T2I predicate
  if (dtype == SWAT_BF16) return SWAT_ENGINE_TC;
  return q->n_fstages > 0 ? SWAT_ENGINE_TC : SWAT_ENGINE_SIMT;
}
float scan_eps(const swat_queries* q, int32_t dtype, int32_t engine) {
  return (engine == SWAT_ENGINE_TC && dtype == SWAT_F32) ? q->eps_conv : kEpsAccum;
}

// the two passes of the tensor-core in-pass predicate (ScanArgs::bits_out / pass_bits)
struct ScanBits {
  uint32_t* out = nullptr;          // pass A: write the predicate bitmap (dense mode, nothing is selected)
  const uint32_t* pass = nullptr;   // pass B: keep a survivor only if its bit is set
  int64_t words = 0;                // words per class
  float thr = 0.0f;                 // pass A threshold
};

// launch_scan_tc, with the CTA phase stamps printed when SWAT_SCAN_TRACE is set (diagnostics: synchronises)
int32_t launch_scan_traced(const CUtensorMap* tm_bank, const swat_queries* q, TcArgs p, bool part, bool dense, bool f32, int grid, cudaStream_t stream) {
  static const bool tracing = getenv("SWAT_SCAN_TRACE") != nullptr;
  if (!tracing) {
    CU_OK(launch_scan_tc(tm_bank, &q->tm_q, p, q->ctas, q->reduce, part, dense, f32, grid, stream));
    return SWAT_OK;
  }
  static unsigned long long* d_trace = nullptr;
  if (!d_trace) CU_OK(cudaMalloc(&d_trace, 1024 * 8 * 8));
  CU_OK(cudaMemsetAsync(d_trace, 0, 1024 * 8 * 8, stream));
  p.trace = d_trace;
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  cudaEventCreate(&t0); cudaEventCreate(&t1); cudaEventRecord(t0, stream);
  CU_OK(launch_scan_tc(tm_bank, &q->tm_q, p, q->ctas, q->reduce, part, dense, f32, grid, stream));
  cudaEventRecord(t1, stream);
  std::vector<unsigned long long> h(static_cast<size_t>(grid) * 8);
  CU_OK(cudaMemcpyAsync(h.data(), d_trace, h.size() * 8, cudaMemcpyDeviceToHost, stream));
  CU_OK(cudaStreamSynchronize(stream));
  float ms = 0; cudaEventElapsedTime(&ms, t0, t1); cudaEventDestroy(t0); cudaEventDestroy(t1);
  unsigned long long first = ~0ull, last = 0;
  for (int b = 0; b < grid; ++b) { first = std::min(first, h[b * 8]); last = std::max(last, h[b * 8 + 7]); }
  double mx[8] = {0}, mn[8];
  for (int i = 0; i < 8; ++i) mn[i] = 1e30;
  for (int b = 0; b < grid; ++b)
    for (int i = 0; i < 8; ++i) if (h[b * 8 + i]) { const double v = (h[b * 8 + i] - first) * 1e-3; mx[i] = std::max(mx[i], v); mn[i] = std::min(mn[i], v); }
  fprintf(stderr, "[swat trace] scan rows=%lld%s event %.1f us, first entry -> last exit %.1f us | us since first entry (min/max over CTAs): "
                  "entry %.1f/%.1f prologue %.1f/%.1f queries %.1f/%.1f mma1 %.1f/%.1f mmaN %.1f/%.1f epi1 %.1f/%.1f epiN %.1f/%.1f exit %.1f/%.1f\n",
          (long long)p.s.n_rows, dense ? " dense" : "", ms * 1e3, (last - first) * 1e-3, mn[0], mx[0], mn[1], mx[1], mn[2], mx[2], mn[3], mx[3], mn[4], mx[4],
          mn[5], mx[5], mn[6], mx[6], mn[7], mx[7]);
  return SWAT_OK;
}

int32_t scan_view(swat_job* job, const void* d_bank, int32_t dtype, int64_t n_rows, int64_t row_base, const void* d_t2i_bank,
                  float t2i_threshold, const int32_t* d_row_class, const uint32_t* d_exclude, int32_t engine, float* dense_out,
                  cudaStream_t stream, const ScanBits* bits = nullptr) {
  swat_ctx* ctx = job->ctx;
  const swat_queries* q = job->q;
  if (n_rows == 0) return SWAT_OK;
  if (n_rows < 0 || row_base < 0 || row_base + n_rows > 0xFFFFFFFEll || n_rows > 0x7FFFFF00ll)
    return fail(SWAT_ERR_INVALID, "bank view out of range: n_rows=%lld row_base=%lld (shard-local row ids are 32-bit)",
                (long long)n_rows, (long long)row_base);
  if (dtype != SWAT_BF16 && dtype != SWAT_F32) return fail(SWAT_ERR_INVALID, "dtype must be SWAT_BF16 or SWAT_F32");
  if ((reinterpret_cast<uintptr_t>(d_bank) & 15) != 0) return fail(SWAT_ERR_INVALID, "bank pointer must be 16-byte aligned");
  const bool f32 = dtype == SWAT_F32;
  // dense scores are the S1 primitives (t2t_similarity :397-416) and the zero-shot logits: exact fp32 arithmetic for fp32 banks
  if (engine == SWAT_ENGINE_AUTO) engine = (dense_out != nullptr && f32) ? SWAT_ENGINE_SIMT : resolve_engine(q, dtype, d_t2i_bank != nullptr);
  const bool dense = dense_out != nullptr || (bits != nullptr && bits->out != nullptr);   // nothing is selected
  if (engine == SWAT_ENGINE_TC && d_t2i_bank != nullptr)
    return fail(SWAT_ERR_UNSUPPORTED, "the tcgen05 engine does not evaluate the in-pass T2I predicate");
  ScanArgs a;
  a.st = job->st;
  a.col_class = q->d_col_class;
  a.col_count = q->d_col_count;
  a.n_cols = q->n_cols;
  a.n_rows = n_rows;
  a.row_base = static_cast<uint32_t>(row_base);
  a.row_class = d_row_class;
  a.exclude = d_exclude;
  a.t2i_thr = t2i_threshold;
  a.dense_out = dense_out;
  a.dense_ld = q->C;
  a.dense_transposed = 0;
  a.n_classes = q->C;
  a.bits_out = bits ? bits->out : nullptr;
  a.pass_bits = bits ? bits->pass : nullptr;
  a.bits_words = bits ? bits->words : 0;
  a.bits_thr = bits ? bits->thr : 0.0f;
  job->last_stream = stream;
  if (engine == SWAT_ENGINE_TC) {
    if ((f32 ? q->n_fstages : q->n_stages) <= 0)
      return fail(SWAT_ERR_UNSUPPORTED, "query block does not fit in shared memory for the tcgen05 engine");
    TcArgs p{};
    p.n_qb = q->n_qb;
    p.n_blk = q->n_blk;
    p.n_stages = f32 ? q->n_opstages_f32 : q->n_stages;
    p.n_fstages = f32 ? q->n_fstages : 0;
    p.qb_base = 0;
    p.qb_count = q->n_qb;
    p.smem_b_bytes = static_cast<uint32_t>(8) * (q->n_blk / q->ctas) * 128;
    p.bank_hint = (q->n_qb == 1) ? 0x12F0000000000000ull /* evict_first: streamed once */ : 0x1000000000000000ull;
    p.blk_class = q->d_blk_class;
    p.blk_split = q->d_blk_split;
    int grid = ctx->sm_count;
    if (ctx->max_ctas > 0) grid = std::min(grid, ctx->max_ctas);
    grid = std::max(q->ctas, grid / q->ctas * q->ctas);
    if (!dense && static_cast<uint32_t>(grid) * kTcEpiWarps > job->st.n_lists)
      return fail(SWAT_ERR_INVALID, "grid of %d CTAs needs %d survivor lists, job has %u", grid, grid * kTcEpiWarps, job->st.n_lists);
    const char* bank = static_cast<const char*>(d_bank);
    // ---- threshold bootstrap: a fresh job would append every non-negative score of its first waves
    // (thresholds start at the user threshold).  Dense-score a small prefix with the same kernel, take
    // its exact top-k_fetch per class, and start the real scan with selective thresholds.
    int64_t B = 0;
    // (not with a predicate bitmap: the prefix's best rows need not pass the predicate, their scores are no valid bound)
    if (!dense && a.pass_bits == nullptr && job->fresh && ctx->bootstrap_rows > 0 && d_row_class == nullptr && d_exclude == nullptr) {
      B = std::min<int64_t>(ctx->bootstrap_rows, (512ll << 20) / (4ll * q->C)) / 256 * 256;
      if (B < 8192 || n_rows < 8 * B || static_cast<uint32_t>(grid) * kTcEpiWarps >= job->st.n_lists) B = 0;
    }
    if (B > 0) {
      SW_OK(ctx->w_boot.ensure(static_cast<size_t>(q->C) * B * 4));
      CUtensorMap tm_pre;
      SW_OK(encode_2d(ctx, &tm_pre, bank, static_cast<uint64_t>(B), 128, dtype));
      TcArgs pd = p;
      pd.s = a;
      pd.s.n_rows = B;
      pd.s.dense_out = ctx->w_boot.as<float>();
      pd.s.dense_ld = B;
      pd.s.dense_transposed = 1;
      pd.s.bits_out = nullptr;
      pd.bank_hint = 0x1000000000000000ull;
      for (int b0 = 0; b0 < q->n_qb; b0 += grid / q->ctas) {
        pd.qb_base = b0;
        pd.qb_count = std::min(grid / q->ctas, q->n_qb - b0);
        SW_OK(launch_scan_traced(&tm_pre, q, pd, false, true, f32, grid, stream));
      }
      CU_OK(launch_bootstrap(job->st, q->C, ctx->w_boot.as<float>(), static_cast<uint32_t>(B), a.row_base,
                             static_cast<uint32_t>(grid) * kTcEpiWarps, stream));
      ctx->launches += 2;
      bank += static_cast<size_t>(B) * kDim * (f32 ? 4 : 2);
      a.n_rows = n_rows - B;
      a.row_base += static_cast<uint32_t>(B);
    }
    CUtensorMap tm_bank;
    SW_OK(encode_2d(ctx, &tm_bank, bank, static_cast<uint64_t>(a.n_rows), 128, dtype));
    p.s = a;
    // Several Q blocks rarely divide the CTA pairs evenly (74 pairs: 4 blocks leave 2 idle, 16 leave 10, 21 leave 11) and
    // blocks served by different numbers of pairs drift apart, so the bank is re-read from HBM per block.  The unit
    // plan cuts the work into lcm(n_qb, pairs) equal (Q block x tile range) units, `pairs` per launch.
    const int pairs = grid / q->ctas;
    const int64_t tiles = (a.n_rows + 128 * q->ctas - 1) / (128 * q->ctas);
    int launches = 1;
    if (ctx->unit_plan && !dense && q->n_qb > 1 && pairs % q->n_qb != 0) {
      const int g = std::gcd(q->n_qb, pairs);
      const int per_pair = q->n_qb / g, ranges = pairs / g;        // units per pair, tile ranges
      if (per_pair <= 64 && tiles >= 16ll * ranges) {              // units of >= 16 tiles, else one launch does
        launches = per_pair;
        p.n_ranges = ranges;
      }
    }
    if (p.n_ranges == 0 && q->n_qb > pairs) launches = (q->n_qb + pairs - 1) / pairs;   // legacy plan: `pairs` Q blocks per launch
    // one Q block: pairs take tiles from a global counter instead of a fixed stride (TcArgs::tile_sched)
    if (ctx->dyn_tiles && !dense && q->n_qb == 1 && p.n_ranges == 0 && launches == 1 && tiles > 4ll * pairs) {
      const size_t bytes = (2 + static_cast<size_t>(pairs) * kTileRing) * 8;
      SW_OK(ctx->w_tiles.ensure(bytes));
      CU_OK(cudaMemsetAsync(ctx->w_tiles.p, 0, bytes, stream));
      p.tile_sched = ctx->w_tiles.as<unsigned long long>();
    }
    int lock_window = ctx->lock_window;
    if (lock_window < 0) {
      lock_window = 0;
      if (q->n_qb >= 4 && q->n_qb <= 8 && !dense && ctx->h_probe != nullptr && ctx->clock_khz > 0) {
        const unsigned long long mhz = *static_cast<volatile unsigned long long*>(ctx->h_probe);
        if (mhz > 0) {
          const double ratio = static_cast<double>(mhz) * 1000.0 / ctx->clock_khz;
          const bool was = ctx->lock_auto_on;
          if (ratio < 0.70) ctx->lock_auto_on = true;
          else if (ratio > 0.80) ctx->lock_auto_on = false;
          static const bool verbose = getenv("SWAT_DEBUG") != nullptr && atoi(getenv("SWAT_DEBUG")) > 1;
          if ((was != ctx->lock_auto_on && getenv("SWAT_DEBUG")) || verbose)
            fprintf(stderr, "[swat] SM clock %llu MHz during the last %d-block scan (%.2f of the maximum): lockstep window %s\n", mhz, q->n_qb, ratio,
                    ctx->lock_auto_on ? "on" : "off");
        }
        if (ctx->lock_auto_on) lock_window = 4;
        p.clock_probe = ctx->d_probe;
      }
    }
    const bool lockstep = lock_window > 0 && q->n_qb > 1 && !dense;
    if (lockstep) {
      SW_OK(ctx->w_progress.ensure(static_cast<size_t>(launches) * pairs * 4));
      CU_OK(cudaMemsetAsync(ctx->w_progress.p, 0, static_cast<size_t>(launches) * pairs * 4, stream));
      p.lock_window = lock_window;
    }
    for (int i = 0; i < launches; ++i) {
      p.progress = lockstep ? ctx->w_progress.as<uint32_t>() + static_cast<size_t>(i) * pairs : nullptr;
      if (p.n_ranges > 0) {
        p.unit_base = i * pairs;
      } else {
        p.qb_base = i * pairs;
        p.qb_count = std::min(pairs, q->n_qb - p.qb_base);
      }
      SW_OK(launch_scan_traced(&tm_bank, q, p, d_row_class != nullptr, dense, f32, grid, stream));
    }
    ctx->launches += launches - 1;
  } else {
    const void* qp = (dtype == SWAT_BF16) ? static_cast<const void*>(q->d_qp_bf16) : static_cast<const void*>(q->d_qp_f32);
    CU_OK(launch_scan_simt(a, d_bank, d_t2i_bank, qp, dtype, q->reduce, d_row_class != nullptr, dense, stream));
  }
  ctx->launches += 1;
  if (!dense) job->fresh = false;
  return SWAT_OK;
}

int64_t auto_cap(const swat_ctx* ctx, int k_fetch) {
  if (ctx->cand_cap > 0) return ctx->cand_cap;
  return 2ll * k_fetch + 4096;          // candidates at/above the final threshold: ~k_fetch + one histogram bin + ties
}
int64_t auto_list_entries(const swat_ctx* ctx, int n_classes, int k_fetch) {
  if (ctx->list_entries > 0) return ctx->list_entries;
  // survivors over a whole scan: ~k_fetch*(1 + ln(N/first wave)) per class plus the first-wave burst
  const int64_t per_class = 12ll * std::max(k_fetch, 512) + 16384;
  return std::max<int64_t>(4ll << 20, per_class * n_classes * 3 / 2);
}

void job_set_params(swat_job* j, const swat_queries* q, int32_t k_fetch, float thr) {
  JobState& st = j->st;
  j->q = q;
  st.k_fetch = static_cast<uint32_t>(k_fetch);
  st.k_class = nullptr;
  st.thr = thr;
  float lo = std::max(thr, -1.0f);
  float hi = std::max(1.0f, lo + 1.0f / 64.0f);
  if (lo >= 1.0f) hi = lo + 1.0f;
  st.hist_lo = lo;
  st.hist_scale = static_cast<float>(kHistBins) / (hi - lo);
  st.hist_inv_scale = (hi - lo) / static_cast<float>(kHistBins);
}

int32_t job_create_cap(swat_ctx* ctx, const swat_queries* q, int32_t k_fetch, float thr, int64_t cap, int64_t list_entries,
                       swat_job** out, int n_classes_alloc = 0) {
  if (!ctx || !q || !out) return fail(SWAT_ERR_INVALID, "null argument");
  if (k_fetch < 1 || k_fetch > kMaxKFetch) return fail(SWAT_ERR_UNSUPPORTED, "k_fetch must be in [1, %d], got %d", kMaxKFetch, k_fetch);
  if (thr != thr) return fail(SWAT_ERR_INVALID, "threshold is NaN");
  (void)cudaGetLastError();
  CU_OK(cudaSetDevice(ctx->device));
  swat_job* j = new swat_job();
  j->ctx = ctx;
  const size_t C = static_cast<size_t>(std::max(q->C, n_classes_alloc));
  j->n_classes_alloc = static_cast<int>(C);
  JobState& st = j->st;
  memset(&st, 0, sizeof(st));
  st.cap = static_cast<uint32_t>(std::min<int64_t>(cap, 0x7fffffff));
  st.n_lists = 2048;
  const int64_t private_lists = std::max(1, ctx->sm_count) * static_cast<int64_t>(kTcEpiWarps);   // one list per epilogue warp of the tcgen05 kernel
  st.list_cap = static_cast<uint32_t>(std::min<int64_t>(((list_entries + private_lists - 1) / private_lists + 255) / 256 * 256, 0x7fffff00));
  job_set_params(j, q, k_fetch, thr);
  cudaError_t e = cudaMalloc(&st.tau_enc, C * 4);
  if (e == cudaSuccess) e = cudaMalloc(&st.count, C * 4);
  if (e == cudaSuccess) e = cudaMalloc(&st.hist, C * kHistBins * 4);
  if (e == cudaSuccess) e = cudaMalloc(&st.cand, C * static_cast<size_t>(st.cap) * 8);
  if (e == cudaSuccess) e = cudaMalloc(&st.list, static_cast<size_t>(st.n_lists) * st.list_cap * sizeof(uint4));
  if (e == cudaSuccess) e = cudaMalloc(&st.list_count, static_cast<size_t>(st.n_lists) * 4);
  if (e == cudaSuccess) e = cudaMalloc(&st.flags, 16);
  if (e == cudaSuccess) e = cudaMalloc(&j->d_k_class, C * 4);
  if (getenv("SWAT_DEBUG")) fprintf(stderr, "[swat] job allocated: C=%zu k_fetch=%d cap=%u lists=%u x %u entries (%.1f MB)\n", C, k_fetch, st.cap,
                                    st.n_lists, st.list_cap, static_cast<double>(st.n_lists) * st.list_cap * 16 / 1e6);
  if (e != cudaSuccess) {
    swat_job_destroy(j);
    return fail(SWAT_ERR_CUDA, "job allocation failed (C=%zu, cap=%lld, list entries=%lld): %s", C, (long long)cap,
                (long long)list_entries, cudaGetErrorString(e));
  }
  *out = j;
  return SWAT_OK;
}

// whole-pipeline calls reuse one job allocation per ctx as long as the sizes fit
int32_t acquire_job(swat_ctx* ctx, const swat_queries* q, int32_t k_fetch, float thr, int64_t cap, int64_t list_entries, swat_job** out) {
  swat_job* j = ctx->cached_job;
  struct Timer {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    ~Timer() {
      const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
      if (ms > 1.0 && getenv("SWAT_DEBUG")) fprintf(stderr, "[swat] acquire_job took %.2f ms\n", ms);
    }
  } timer;
  const int64_t private_lists = std::max(1, ctx->sm_count) * static_cast<int64_t>(kTcEpiWarps);
  if (j && j->n_classes_alloc >= q->C && j->st.cap >= cap && static_cast<int64_t>(j->st.list_cap) * private_lists >= list_entries) {
    if (k_fetch < 1 || k_fetch > kMaxKFetch) return fail(SWAT_ERR_UNSUPPORTED, "k_fetch must be in [1, %d], got %d", kMaxKFetch, k_fetch);
    job_set_params(j, q, k_fetch, thr);
    *out = j;
    return SWAT_OK;
  }
  int c_alloc = q->C;
  if (j) {   // grow-only: alternating callers (main pass / escalated sub-pass) must not thrash the allocation
    c_alloc = std::max(c_alloc, j->n_classes_alloc);
    cap = std::max<int64_t>(cap, j->st.cap);
    list_entries = std::max<int64_t>(list_entries, static_cast<int64_t>(j->st.list_cap) * private_lists);
    swat_job_destroy(j);
    ctx->cached_job = nullptr;
  }
  SW_OK(job_create_cap(ctx, q, k_fetch, thr, cap, list_entries, &j, c_alloc));
  ctx->cached_job = j;
  *out = j;
  return SWAT_OK;
}

int32_t job_flags(swat_job* job, uint32_t* flags) {
  CU_OK(cudaMemcpyAsync(flags, job->st.flags, 4, cudaMemcpyDeviceToHost, job->last_stream));
  CU_OK(cudaStreamSynchronize(job->last_stream));
  return SWAT_OK;
}

// describes where the banks live for the whole-pipeline calls
struct BankSrc {
  bool host = false;
  const void* t2t = nullptr; const void* t2i = nullptr;
  int dtype = 0; int64_t n_rows = 0;
  const int32_t* row_class = nullptr; const uint32_t* exclude = nullptr;
  // host banks in pinned (page-locked) memory: device-visible aliases, the re-score kernel reads candidate rows
  // straight over PCIe instead of a host-side gather
  const void* t2t_mapped = nullptr; const void* t2i_mapped = nullptr;
};

size_t elem_size(int dtype) { return dtype == SWAT_BF16 ? 2 : 4; }

// device-visible alias of a page-locked host allocation (cudaHostAlloc / cudaHostRegister), nullptr for pageable memory
const void* mapped_alias(const void* h) {
  if (!h) return nullptr;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, h) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
  return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
}

// one pass over the whole bank, folding every view into `job`
int32_t scan_all(swat_ctx* ctx, swat_job* job, const BankSrc& b, const ScanBits* bits, cudaStream_t stream) {
  const bool dual = false;         // the fp32-FMA in-pass predicate is reachable through swat_job_scan only
  const float t2i_thr = 0.0f;
  if (!b.host) {
    const int64_t max_view = 0x40000000ll;   // TMA coordinates are int32
    for (int64_t r0 = 0; r0 < b.n_rows; r0 += max_view) {
      const int64_t n = std::min(max_view, b.n_rows - r0);
      const size_t off = static_cast<size_t>(r0) * kDim * elem_size(b.dtype);
      SW_OK(scan_view(job, static_cast<const char*>(b.t2t) + off, b.dtype, n, r0,
                      dual ? static_cast<const char*>(b.t2i) + off : nullptr, t2i_thr,
                      b.row_class ? b.row_class + r0 : nullptr, b.exclude ? b.exclude + r0 / 32 : nullptr, SWAT_ENGINE_AUTO,
                      nullptr, stream, bits));
    }
    return SWAT_OK;
  }
  // host bank: stream chunks through three device staging buffers, copies overlapped with the scan
  const int64_t chunk = std::max<int64_t>(4096, ctx->host_chunk_rows / 32 * 32);
  const size_t row_bytes = static_cast<size_t>(kDim) * elem_size(b.dtype);
  const int nbuf = 3;
  for (int i = 0; i < nbuf; ++i) {
    SW_OK(ctx->w_stage[i].ensure(static_cast<size_t>(chunk) * row_bytes * (dual ? 2 : 1)));
    if (b.row_class) SW_OK(ctx->w_rc[i].ensure(static_cast<size_t>(chunk) * 4));
    if (b.exclude) SW_OK(ctx->w_ex[i].ensure(static_cast<size_t>(chunk) / 8 + 8));
  }
  CU_OK(cudaEventRecord(ctx->ev_used[0], stream));   // order the copy stream after prior work on `stream`
  CU_OK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_used[0], 0));
  int64_t i = 0;
  for (int64_t r0 = 0; r0 < b.n_rows; r0 += chunk, ++i) {
    const int s = static_cast<int>(i % nbuf);
    const int64_t n = std::min(chunk, b.n_rows - r0);
    if (i >= nbuf) CU_OK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_used[s], 0));
    char* dst = ctx->w_stage[s].as<char>();
    CU_OK(cudaMemcpyAsync(dst, static_cast<const char*>(b.t2t) + static_cast<size_t>(r0) * row_bytes, static_cast<size_t>(n) * row_bytes,
                          cudaMemcpyHostToDevice, ctx->copy_stream));
    ctx->timing[5] += static_cast<double>(n) * row_bytes;
    char* dst2 = nullptr;
    if (dual) {
      dst2 = dst + static_cast<size_t>(chunk) * row_bytes;
      CU_OK(cudaMemcpyAsync(dst2, static_cast<const char*>(b.t2i) + static_cast<size_t>(r0) * row_bytes, static_cast<size_t>(n) * row_bytes,
                            cudaMemcpyHostToDevice, ctx->copy_stream));
      ctx->timing[5] += static_cast<double>(n) * row_bytes;
    }
    if (b.row_class) {
      CU_OK(cudaMemcpyAsync(ctx->w_rc[s].p, b.row_class + r0, static_cast<size_t>(n) * 4, cudaMemcpyHostToDevice, ctx->copy_stream));
      ctx->timing[5] += static_cast<double>(n) * 4;
    }
    if (b.exclude) {
      CU_OK(cudaMemcpyAsync(ctx->w_ex[s].p, b.exclude + r0 / 32, static_cast<size_t>((n + 31) / 32) * 4, cudaMemcpyHostToDevice, ctx->copy_stream));
      ctx->timing[5] += static_cast<double>((n + 31) / 32) * 4;
    }
    CU_OK(cudaEventRecord(ctx->ev_copied[s], ctx->copy_stream));
    CU_OK(cudaStreamWaitEvent(stream, ctx->ev_copied[s], 0));
    SW_OK(scan_view(job, dst, b.dtype, n, r0, dst2, t2i_thr, b.row_class ? ctx->w_rc[s].as<int32_t>() : nullptr,
                    b.exclude ? ctx->w_ex[s].as<uint32_t>() : nullptr, SWAT_ENGINE_AUTO, nullptr, stream, bits));
    CU_OK(cudaEventRecord(ctx->ev_used[s], stream));
  }
  return SWAT_OK;
}

int32_t ensure_status(swat_ctx* ctx, size_t n_ints) {
  if (n_ints <= ctx->h_status_cap) return SWAT_OK;
  if (ctx->h_status) cudaFreeHost(ctx->h_status);
  ctx->h_status = nullptr; ctx->h_status_cap = 0;
  CU_OK(cudaMallocHost(reinterpret_cast<void**>(&ctx->h_status), n_ints * sizeof(int32_t)));
  ctx->h_status_cap = n_ints;
  return SWAT_OK;
}
int32_t ensure_pinned(swat_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->h_pinned_cap) return SWAT_OK;
  if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
  ctx->h_pinned = nullptr; ctx->h_pinned_cap = 0;
  CU_OK(cudaMallocHost(&ctx->h_pinned, bytes));
  ctx->h_pinned_cap = bytes;
  return SWAT_OK;
}

int32_t run_pipeline(swat_ctx* ctx, const swat_queries* q, const BankSrc& b, int64_t row_offset, int32_t k, float thr, float t2i_thr,
                     float* d_out_scores, int64_t* d_out_rows, float* d_out_t2i, int32_t* d_out_counts, cudaStream_t stream,
                     int32_t k_fetch_init, int depth);

constexpr int32_t kSwapPass = kMaxKFetch + 1;    // k_fetch_init beyond the widest over-fetch: try the bank-swap pass, then the in-pass predicate
constexpr int32_t kForceDual = kMaxKFetch + 2;   // ... go straight to the in-pass predicate

// per-class candidate lists as swat_job_select leaves them (ranked on the scan's approximate score)
struct CandLists {
  const float* scores; const int64_t* rows; const int32_t* counts; const int32_t* trunc; int stride;
};

// Exact re-score + accept walk of candidate lists against the banks of `b` (select.cu): canonical fp32 scores for the
// ranking bank and, with use_aux, the predicate bank; results in walk order.  Host banks: pinned memory is read in
// place by the kernel (zero-copy over PCIe, only the candidates' rows move); pageable memory is gathered on the host.
int32_t walk_candidates(swat_ctx* ctx, const swat_queries* q, const BankSrc& b, bool use_aux, int64_t row_offset, const CandLists& cl,
                        int32_t k, float thr, float aux_thr, float eps, bool all_or_nothing, float* o_scores, int64_t* o_rows,
                        float* o_aux, int32_t* o_counts, float* o_limit, int32_t* o_incomplete, cudaStream_t stream,
                        const swat_queries* q_aux = nullptr, int32_t* d_eps_violation = nullptr, int32_t* d_status = nullptr,
                        const uint32_t* d_job_flags = nullptr) {
  const int C = q->C;
  const size_t n_slots = static_cast<size_t>(C) * cl.stride;
  SW_OK(ctx->w_exact.ensure(n_slots * 4));
  SW_OK(ctx->w_aux.ensure(n_slots * 4));
  WalkArgs w;
  memset(&w, 0, sizeof(w));
  w.dtype = b.dtype;
  w.queries = (b.dtype == SWAT_BF16) ? static_cast<const void*>(q->d_q_bf16) : static_cast<const void*>(q->d_q_f32);
  w.class_begin = q->d_class_begin;
  w.reduce = q->reduce;
  const swat_queries* qa = q_aux ? q_aux : q;
  w.aux_queries = (b.dtype == SWAT_BF16) ? static_cast<const void*>(qa->d_q_bf16) : static_cast<const void*>(qa->d_q_f32);
  w.aux_class_begin = qa->d_class_begin;
  w.aux_reduce = qa->reduce;
  w.cand_scores = cl.scores; w.cand_rows = cl.rows; w.cand_counts = cl.counts; w.truncated = cl.trunc;
  w.stride = cl.stride; w.k = k; w.n_classes = C;
  w.thr = thr; w.aux_thr = aux_thr; w.eps = eps; w.all_or_nothing = all_or_nothing ? 1 : 0;
  w.exact_scratch = ctx->w_exact.as<float>(); w.aux_scratch = ctx->w_aux.as<float>();
  w.out_scores = o_scores; w.out_rows = o_rows; w.out_aux = o_aux; w.out_counts = o_counts; w.out_limit = o_limit; w.incomplete = o_incomplete;
  w.eps_violation = d_eps_violation;
  w.status = d_status; w.job_flags = d_job_flags;
  w.key_row_base = row_offset;
  const bool mapped = b.host && b.t2t_mapped && (!use_aux || b.t2i_mapped);
  if (!b.host || mapped) {
    w.t2t_bank = b.host ? b.t2t_mapped : b.t2t;
    w.aux_bank = use_aux ? (b.host ? b.t2i_mapped : b.t2i) : nullptr;
    w.bank_rows = b.n_rows;
    w.bank_row_base = row_offset;
    w.gather_index = nullptr;
    w.lazy_t2t = (b.host && use_aux) ? 1 : 0;
    if (b.host) ctx->timing[5] += static_cast<double>(n_slots) * kDim * elem_size(b.dtype) * (use_aux ? 2 : 1);   // upper bound: every slot filled
  } else {
    // pageable host banks: gather the candidates' rows on the host, ship the compact blocks
    const size_t row_bytes = static_cast<size_t>(kDim) * elem_size(b.dtype);
    std::vector<int64_t> h_rows(n_slots);
    std::vector<int32_t> h_counts(C);
    CU_OK(cudaMemcpyAsync(h_rows.data(), cl.rows, n_slots * 8, cudaMemcpyDeviceToHost, stream));
    CU_OK(cudaMemcpyAsync(h_counts.data(), cl.counts, static_cast<size_t>(C) * 4, cudaMemcpyDeviceToHost, stream));
    CU_OK(cudaStreamSynchronize(stream));
    ctx->timing[6] += static_cast<double>(n_slots * 8 + C * 4);
    std::vector<int64_t> h_index(n_slots, -1);
    std::vector<int64_t> src;
    src.reserve(n_slots);
    for (int c = 0; c < C; ++c)
      for (int j = 0; j < std::min(h_counts[c], cl.stride); ++j) {
        h_index[static_cast<size_t>(c) * cl.stride + j] = static_cast<int64_t>(src.size());
        src.push_back(h_rows[static_cast<size_t>(c) * cl.stride + j] - row_offset);
      }
    const size_t n_src = src.size();
    const int n_banks = use_aux ? 2 : 1;
    SW_OK(ensure_pinned(ctx, std::max<size_t>(n_src, 1) * row_bytes * n_banks));
    char* stage = static_cast<char*>(ctx->h_pinned);
    const char* banks[2] = {static_cast<const char*>(b.t2t), static_cast<const char*>(b.t2i)};
    const unsigned nt = n_src * n_banks < 4096 ? 1u : std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
    auto work = [&](unsigned wi) {
      for (int bk = 0; bk < n_banks; ++bk)
        for (size_t i = wi; i < n_src; i += nt)
          memcpy(stage + (static_cast<size_t>(bk) * n_src + i) * row_bytes, banks[bk] + static_cast<size_t>(src[i]) * row_bytes, row_bytes);
    };
    if (nt == 1) work(0);
    else {
      std::vector<std::thread> th;
      for (unsigned wi = 0; wi < nt; ++wi) th.emplace_back(work, wi);
      for (auto& x : th) x.join();
    }
    SW_OK(ctx->w_img.ensure(std::max<size_t>(n_src, 1) * row_bytes * n_banks));
    SW_OK(ctx->w_idx.ensure(n_slots * 8));
    CU_OK(cudaMemcpyAsync(ctx->w_img.p, stage, n_src * row_bytes * n_banks, cudaMemcpyHostToDevice, stream));
    CU_OK(cudaMemcpyAsync(ctx->w_idx.p, h_index.data(), n_slots * 8, cudaMemcpyHostToDevice, stream));
    CU_OK(cudaStreamSynchronize(stream));   // h_index is pageable and dies at scope end
    ctx->timing[5] += static_cast<double>(n_src * row_bytes * n_banks + n_slots * 8);
    w.t2t_bank = ctx->w_img.p;
    w.aux_bank = use_aux ? ctx->w_img.as<char>() + n_src * row_bytes : nullptr;
    w.bank_rows = static_cast<int64_t>(n_src);
    w.bank_row_base = 0;
    w.gather_index = ctx->w_idx.as<int64_t>();
  }
  CU_OK(launch_rescore_walk(w, stream));
  ctx->launches += kWalkLaunches;
  return SWAT_OK;
}

// splice result rows of some classes into the final arrays with one kernel (dst class ids / source indices uploaded once)
int32_t splice_results(swat_ctx* ctx, const std::vector<int32_t>& dst_cls, const std::vector<int32_t>& src_idx, int32_t k,
                       const float* s_scores, const int64_t* s_rows, const float* s_aux, const int32_t* s_counts, float* d_scores,
                       int64_t* d_rows, float* d_aux, int32_t* d_counts, cudaStream_t stream) {
  const size_t n = dst_cls.size();
  if (n == 0) return SWAT_OK;
  SW_OK(ctx->w_splice.ensure(2 * n * 4));
  std::vector<int32_t> both(dst_cls);
  both.insert(both.end(), src_idx.begin(), src_idx.end());
  CU_OK(cudaMemcpyAsync(ctx->w_splice.p, both.data(), 2 * n * 4, cudaMemcpyHostToDevice, stream));
  CU_OK(launch_splice(ctx->w_splice.as<int32_t>(), ctx->w_splice.as<int32_t>() + n, static_cast<int>(n), k, s_scores, s_rows, s_aux, s_counts,
                      d_scores, d_rows, d_aux, d_counts, stream));
  CU_OK(cudaStreamSynchronize(stream));        // `both` is pageable and dies at scope end
  ctx->launches += 1;
  return SWAT_OK;
}

// Bank-swap escalation pass.  A class whose T2T-ordered walk ran out of candidates has FEW rows passing the T2I
// predicate (that is why k were not found among the best 4096 by T2T).  So enumerate the passers instead: scan the
// IMAGE bank with the T2I threshold (minus the scan's error bound) as the row threshold; if fewer than 4096 rows of the
// class survive, that list holds every row that can pass the predicate.  The walk then re-scores them exactly against
// both banks and orders them by T2T: exactly the walk of add_t2t_ranked_t2i_tshd_to_split (:507-527), at the cost of
// one tensor-core pass instead of the fp32 two-bank scan.
// unresolved: classes with 4096 or more such rows (plenty of passers, all with a low T2T score) -- left to the caller.
// only: nullable [C] mask -- results are written for these classes only (the others keep what the ladder proved).
int32_t swap_pass(swat_ctx* ctx, const swat_queries* q, const BankSrc& b, int64_t row_offset, int32_t k, float thr, float t2i_thr,
                  float* d_out_scores, int64_t* d_out_rows, float* d_out_t2i, int32_t* d_out_counts, cudaStream_t stream,
                  const std::vector<int>* only, std::vector<int>* unresolved) {
  const int C = q->C;
  const int32_t kf = kMaxKFetch;
  BankSrc sb = b;
  sb.t2t = b.t2i;              // rows are ranked by their image score here
  sb.t2i = nullptr;
  const float eps = scan_eps(q, b.dtype, resolve_engine(q, b.dtype, false));
  int64_t cap = auto_cap(ctx, kf), list_entries = auto_list_entries(ctx, C, kf);
  const size_t n = static_cast<size_t>(C) * kf;
  SW_OK(ctx->w_scores.ensure(n * 4)); SW_OK(ctx->w_rows.ensure(n * 8)); SW_OK(ctx->w_counts.ensure(static_cast<size_t>(C) * 4));
  SW_OK(ctx->w_trunc.ensure(static_cast<size_t>(C) * 4));
  for (int rounds = 0;; ++rounds) {
    if (rounds > 8) return fail(SWAT_ERR_OVERFLOW, "bank-swap pass: retry budget exhausted");
    swat_job* job = nullptr;
    SW_OK(acquire_job(ctx, q, kf, t2i_thr - eps, cap, list_entries, &job));
    SW_OK(swat_job_reset(job, stream));
    CU_OK(cudaEventRecord(ctx->ev[0], stream));
    SW_OK(scan_all(ctx, job, sb, nullptr, stream));
    CU_OK(cudaEventRecord(ctx->ev[1], stream));
    CU_OK(launch_select(job->st, C, row_offset, ctx->w_scores.as<float>(), ctx->w_rows.as<int64_t>(), ctx->w_counts.as<int32_t>(),
                        ctx->w_trunc.as<int32_t>(), stream));
    job->last_stream = stream;
    ctx->launches += kSelectLaunches;
    uint32_t flags = 0;
    SW_OK(job_flags(job, &flags));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    ctx->timing[0] += ms;
    ctx->timing[4] += 1;
    if (!(flags & 3u)) break;
    const int64_t limit = std::max<int64_t>(b.n_rows, 1 << 16);
    if ((flags & 1u) && cap >= limit) return fail(SWAT_ERR_OVERFLOW, "class candidate overflow with cap >= n_rows");
    if (flags & 1u) cap = std::min<int64_t>(cap * 4, limit);
    if (flags & 2u) list_entries *= 4;
  }
  // walk into scratch outputs, then copy the wanted classes
  DevBuf* w = ctx->w_swap;
  SW_OK(w[0].ensure(static_cast<size_t>(C) * k * 4)); SW_OK(w[1].ensure(static_cast<size_t>(C) * k * 8));
  SW_OK(w[2].ensure(static_cast<size_t>(C) * k * 4)); SW_OK(w[3].ensure(static_cast<size_t>(C) * 4)); SW_OK(w[4].ensure(static_cast<size_t>(C) * 4));
  CandLists cl{ctx->w_scores.as<float>(), ctx->w_rows.as<int64_t>(), ctx->w_counts.as<int32_t>(), ctx->w_trunc.as<int32_t>(), kf};
  SW_OK(walk_candidates(ctx, q, b, true, row_offset, cl, k, thr, t2i_thr, eps, true, w[0].as<float>(), w[1].as<int64_t>(), w[2].as<float>(),
                        w[3].as<int32_t>(), nullptr, w[4].as<int32_t>(), stream));
  SW_OK(ensure_status(ctx, static_cast<size_t>(C) + 1));
  CU_OK(cudaMemcpyAsync(ctx->h_status + 1, w[4].as<int32_t>(), static_cast<size_t>(C) * 4, cudaMemcpyDeviceToHost, stream));
  CU_OK(cudaStreamSynchronize(stream));
  std::vector<char> want(C, only ? 0 : 1);
  if (only) for (int c : *only) want[c] = 1;
  unresolved->clear();
  std::vector<int32_t> take;
  for (int c = 0; c < C; ++c) {
    if (!want[c]) continue;
    if (ctx->h_status[1 + c] != 0) unresolved->push_back(c);
    else take.push_back(c);
  }
  SW_OK(splice_results(ctx, take, take, k, w[0].as<float>(), w[1].as<int64_t>(), d_out_t2i ? w[2].as<float>() : nullptr, w[3].as<int32_t>(),
                       d_out_scores, d_out_rows, d_out_t2i, d_out_counts, stream));
  return SWAT_OK;
}

// Targeted escalation: re-run only the classes whose walk could not be proven exact, with a wider over-fetch (and
// finally the in-pass predicate), then splice their rows into the result.  Partitioned data (one class per row): the
// sub-run gets a row_class array renumbered to the sub-query set's classes (rows of other classes become -1).
int32_t escalate_classes(swat_ctx* ctx, const swat_queries* q, const BankSrc& b, int64_t row_offset, int32_t k, float thr, float t2i_thr,
                         const std::vector<int>& classes, int32_t k_fetch_next, float* d_out_scores, int64_t* d_out_rows,
                         float* d_out_t2i, int32_t* d_out_counts, cudaStream_t stream, int depth) {
  const int n = static_cast<int>(classes.size());
  if (depth > 3) return fail(SWAT_ERR_INCOMPLETE, "escalation nested too deep");
  std::vector<float> hq;
  std::vector<int32_t> coq;
  for (int i = 0; i < n; ++i) {
    const int c = classes[i];
    for (int qi = q->class_begin[c]; qi < q->class_begin[c + 1]; ++qi) {
      hq.insert(hq.end(), q->h_q.begin() + static_cast<size_t>(qi) * kDim, q->h_q.begin() + static_cast<size_t>(qi + 1) * kDim);
      coq.push_back(i);
    }
  }
  swat_queries* sub = nullptr;
  const bool cached = depth == 0;   // nested levels own their sub-query set: the cache entry is in use above them
  for (int i = 0; cached && i < 2 && !sub; ++i)
    if (ctx->esc_q[i] && ctx->esc_parent[i] == q && ctx->esc_classes[i] == classes) sub = ctx->esc_q[i];
  if (!sub) {
    const int slot = ctx->esc_next;
    if (cached && ctx->esc_q[slot]) { swat_queries_destroy(ctx->esc_q[slot]); ctx->esc_q[slot] = nullptr; }
    SW_OK(swat_queries_create(ctx, hq.data(), static_cast<int32_t>(coq.size()), coq.data(), n, q->reduce, &sub));
    if (cached) { ctx->esc_q[slot] = sub; ctx->esc_parent[slot] = q; ctx->esc_classes[slot] = classes; ctx->esc_next = slot ^ 1; }
  }
  BankSrc sb = b;
  int32_t rc = SWAT_OK;
  std::vector<int32_t> h_sub_rc;
  if (b.row_class != nullptr) {
    std::vector<int32_t> map(q->C, -1);
    for (int i = 0; i < n; ++i) map[classes[i]] = i;
    if (b.host) {               // host banks carry a host row_class: renumber it here
      h_sub_rc.resize(static_cast<size_t>(b.n_rows));
      for (int64_t i = 0; i < b.n_rows; ++i) {
        const int32_t c = b.row_class[i];
        h_sub_rc[static_cast<size_t>(i)] = (c >= 0 && c < q->C) ? map[c] : -1;
      }
      sb.row_class = h_sub_rc.data();
    } else {
      DevBuf &m = ctx->e_remap[depth][0], &o = ctx->e_remap[depth][1];
      rc = m.ensure(static_cast<size_t>(q->C) * 4);
      if (rc == SWAT_OK) rc = o.ensure(static_cast<size_t>(std::max<int64_t>(b.n_rows, 1)) * 4);
      if (rc == SWAT_OK) {
        cudaError_t e = cudaMemcpyAsync(m.p, map.data(), static_cast<size_t>(q->C) * 4, cudaMemcpyHostToDevice, stream);
        if (e == cudaSuccess) e = launch_remap_classes(b.row_class, m.as<int32_t>(), q->C, b.n_rows, o.as<int32_t>(), stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);      // `map` is pageable and dies at scope end
        if (e != cudaSuccess) rc = fail(SWAT_ERR_CUDA, "renumbering row classes failed: %s", cudaGetErrorString(e));
        ctx->launches += 1;
      }
      sb.row_class = o.as<int32_t>();
    }
  }
  DevBuf &o_s = ctx->e_bufs[depth][0], &o_r = ctx->e_bufs[depth][1], &o_t = ctx->e_bufs[depth][2], &o_c = ctx->e_bufs[depth][3];
  if (rc == SWAT_OK) rc = o_s.ensure(static_cast<size_t>(n) * k * 4);
  if (rc == SWAT_OK) rc = o_r.ensure(static_cast<size_t>(n) * k * 8);
  if (rc == SWAT_OK) rc = o_t.ensure(static_cast<size_t>(n) * k * 4);
  if (rc == SWAT_OK) rc = o_c.ensure(static_cast<size_t>(n) * 4);
  if (rc == SWAT_OK)
    rc = run_pipeline(ctx, sub, sb, row_offset, k, thr, t2i_thr, o_s.as<float>(), o_r.as<int64_t>(), d_out_t2i ? o_t.as<float>() : nullptr,
                      o_c.as<int32_t>(), stream, k_fetch_next, depth + 1);
  if (rc == SWAT_OK) {
    std::vector<int32_t> dst(classes.begin(), classes.end()), src(n);
    for (int i = 0; i < n; ++i) src[i] = i;
    rc = splice_results(ctx, dst, src, k, o_s.as<float>(), o_r.as<int64_t>(), d_out_t2i ? o_t.as<float>() : nullptr, o_c.as<int32_t>(),
                        d_out_scores, d_out_rows, d_out_t2i, d_out_counts, stream);
  }
  q->last_k_fetch = sub->last_k_fetch;
  if (!cached) swat_queries_destroy(sub);
  return rc;
}

int32_t round_up32(int64_t x) { return static_cast<int32_t>((x + 31) / 32 * 32); }

// First over-fetch of a walk.  Without a predicate the candidates must reach 2 eps below the k-th score (the walk only
// trusts rows above the list's frontier + eps); with one, deep enough for k rows to pass.
int32_t default_k_fetch(const swat_ctx* ctx, int32_t k, bool want_t2i, bool host, float eps) {
  const bool wide = eps > 10.0f * kEpsAccum;        // bf16-rounded scan of an fp32 bank: ~1e-2 of score between k-th and frontier
  int64_t kf;
  if (!want_t2i) kf = wide ? round_up32(k + std::max(1024, k)) : round_up32(k + std::max(64, k / 8));
  else if (ctx->overfetch > 0) kf = ctx->overfetch;
  // Host banks stream over PCIe (~20x slower than the scan): a second pass costs far more than a
  // wider first one, so over-fetch 4k there; HBM-resident banks start at 2k.
  else kf = (host ? std::max(4 * k, 2048) : std::max(2 * k, 1024)) + (wide ? 1024 : 0);
  return static_cast<int32_t>(std::min<int64_t>(std::max<int64_t>(kf, k), kMaxKFetch));
}

// The whole pipeline.  Results land in d_out_* (device).  See swat_topk / swat_topk_host.
int32_t run_pipeline(swat_ctx* ctx, const swat_queries* q, const BankSrc& b, int64_t row_offset, int32_t k, float thr, float t2i_thr,
                     float* d_out_scores, int64_t* d_out_rows, float* d_out_t2i, int32_t* d_out_counts, cudaStream_t stream,
                     int32_t k_fetch_init, int depth) {
  if (k < 1 || k > kMaxKFetch) return fail(SWAT_ERR_UNSUPPORTED, "k must be in [1, %d], got %d", kMaxKFetch, k);
  if (row_offset < 0 || b.n_rows < 0 || row_offset + b.n_rows > 0xFFFFFFFEll)
    return fail(SWAT_ERR_INVALID, "row_offset + n_rows = %lld exceeds the 32-bit row ids of one shard's keys", (long long)(row_offset + b.n_rows));
  const int C = q->C;
  const bool want_t2i = b.t2i != nullptr;
  if (depth == 0) for (int i = 0; i < 8; ++i) ctx->timing[i] = 0;
  const bool can_swap = want_t2i && ctx->swap_pass;
  const bool hinted_swap = k_fetch_init == 0 && depth == 0 && q->all_few_hint && ctx->overfetch == 0;
  if ((k_fetch_init == kSwapPass || hinted_swap) && can_swap) {
    // escalation beyond the widest over-fetch (or a query set whose classes all had too few T2I passers last time):
    // enumerate the T2I passers from the image bank (one tensor-core pass); classes with too many of them for that fall
    // through to the in-pass predicate
    cudaEvent_t ev_b = ctx->ev[6];
    if (depth == 0) CU_OK(cudaEventRecord(ev_b, stream));
    std::vector<int> unresolved;
    SW_OK(swap_pass(ctx, q, b, row_offset, k, thr, t2i_thr, d_out_scores, d_out_rows, d_out_t2i, d_out_counts, stream, nullptr, &unresolved));
    if (hinted_swap && !unresolved.empty()) {
      q->all_few_hint = false;          // a different bank: plenty of passers here, run the T2T-ordered walk after all
    } else {
      if (!unresolved.empty())
        SW_OK(escalate_classes(ctx, q, b, row_offset, k, thr, t2i_thr, unresolved, kForceDual, d_out_scores, d_out_rows, d_out_t2i,
                               d_out_counts, stream, depth));
      q->last_k_fetch = kMaxKFetch;
      if (depth == 0) {
        CU_OK(cudaEventRecord(ctx->ev[7], stream));
        CU_OK(cudaStreamSynchronize(stream));
        float ms = 0;
        cudaEventElapsedTime(&ms, ev_b, ctx->ev[7]);
        ctx->timing[3] = ms;
      }
      return SWAT_OK;
    }
  }
  // in-pass predicate: one tensor-core pass over the image bank writes the per-class bitmap of rows passing T2I, the
  // caption scan then keeps only survivors whose bit is set -- the walk of :507-527 for ANY data, at two passes
  bool dual = want_t2i && k_fetch_init > kMaxKFetch;
  float eps = scan_eps(q, b.dtype, resolve_engine(q, b.dtype, false));
  // every candidate of the in-pass mode passed the (loosened) predicate: k plus slack for the frontier suffices
  int32_t k_fetch = dual ? default_k_fetch(ctx, k, false, b.host, eps)
                         : (k_fetch_init > 0 ? std::min(std::max(k_fetch_init, k), kMaxKFetch) : default_k_fetch(ctx, k, want_t2i, b.host, eps));
  // Classes whose walk needed a deeper over-fetch before start there: the depth is per class (one class
  // with a block of duplicates should not make the other 199 collect four times the candidates).
  // Shards of one dataset behave alike; an escalation re-reads the whole bank.
  std::vector<uint32_t> k_class;
  const bool use_hint = !dual && depth == 0 && k_fetch_init == 0 && ctx->overfetch == 0 && static_cast<int>(q->kclass_hint.size()) == C;
  if (use_hint) {
    int32_t deepest = k_fetch;
    for (int c = 0; c < C; ++c) deepest = std::max(deepest, q->kclass_hint[c]);
    if (deepest > k_fetch) {
      k_class.resize(C);
      for (int c = 0; c < C; ++c) k_class[c] = static_cast<uint32_t>(std::max(k_fetch, q->kclass_hint[c]));
      k_fetch = deepest;
    }
  }
  int64_t cap = auto_cap(ctx, k_fetch);
  int64_t list_entries = auto_list_entries(ctx, C, k_fetch);
  cudaEvent_t ev_begin = ctx->ev[6];
  if (depth == 0) CU_OK(cudaEventRecord(ev_begin, stream));
  for (int rounds = 0;; ++rounds) {
    if (rounds > 24) return fail(SWAT_ERR_OVERFLOW, "retry budget exhausted (cap=%lld lists=%lld k_fetch=%d)", (long long)cap,
                                 (long long)list_entries, k_fetch);
    const int32_t kf = k_fetch;
    swat_job* job = nullptr;
    // the scan ranks by approximate scores: rows within eps below the user threshold may still qualify exactly
    SW_OK(acquire_job(ctx, q, kf, thr - eps, cap, list_entries, &job));
    if (!k_class.empty()) {
      CU_OK(cudaMemcpyAsync(job->d_k_class, k_class.data(), static_cast<size_t>(C) * 4, cudaMemcpyHostToDevice, stream));
      job->h_k_class.clear();
      job->st.k_class = job->d_k_class;
    }
    SW_OK(swat_job_reset(job, stream));
    CU_OK(cudaEventRecord(ctx->ev[0], stream));
    if (dual) {
      ScanBits pa, pb;
      pa.words = pb.words = (b.n_rows + 31) / 32;
      SW_OK(ctx->w_bits.ensure(static_cast<size_t>(C) * std::max<int64_t>(pa.words, 1) * 4));
      pa.out = ctx->w_bits.as<uint32_t>();
      pa.thr = t2i_thr - eps;
      pb.pass = ctx->w_bits.as<uint32_t>();
      BankSrc ib = b;              // pass A ranks nothing: class scores of the image rows against the threshold
      ib.t2t = b.t2i; ib.t2i = nullptr; ib.t2t_mapped = b.t2i_mapped; ib.t2i_mapped = nullptr;
      ib.row_class = nullptr; ib.exclude = nullptr;
      SW_OK(scan_all(ctx, job, ib, &pa, stream));
      SW_OK(scan_all(ctx, job, b, &pb, stream));
      ctx->timing[4] += 1;
    } else {
      SW_OK(scan_all(ctx, job, b, nullptr, stream));
    }
    CU_OK(cudaEventRecord(ctx->ev[1], stream));
    SW_OK(ctx->w_scores.ensure(static_cast<size_t>(C) * kf * 4));
    SW_OK(ctx->w_rows.ensure(static_cast<size_t>(C) * kf * 8));
    SW_OK(ctx->w_counts.ensure(static_cast<size_t>(C) * 4));
    SW_OK(ctx->w_trunc.ensure(static_cast<size_t>(C) * 4));
    // everything the host reads back after the step, one block (WalkArgs::status): [0] overflow word, [1..C] incomplete
    // flags, [1+C..2C] accepted counts, [1+2C] error bound violated (cleared by the select)
    SW_OK(ctx->w_incomplete.ensure((2 * static_cast<size_t>(C) + 2) * 4));
    int32_t* d_status = ctx->w_incomplete.as<int32_t>();
    int32_t* d_violation = d_status + 1 + 2 * C;
    // resident banks: rows leave the select as global ids (row_offset + shard-local row)
    // without a predicate the walk needs the k best rows only: candidates more than 2 eps below the k-th approximate
    // score cannot be among them and are not re-scored (select_kernel)
    CU_OK(launch_select(job->st, C, row_offset, ctx->w_scores.as<float>(), ctx->w_rows.as<int64_t>(), ctx->w_counts.as<int32_t>(),
                        ctx->w_trunc.as<int32_t>(), stream, want_t2i ? 0u : static_cast<uint32_t>(k), 2.0f * eps + 1.0e-6f, d_violation));
    job->last_stream = stream;
    ctx->launches += kSelectLaunches;
    CU_OK(cudaEventRecord(ctx->ev[2], stream));
    // Resident banks run the walk optimistically and read the overflow word together with its `incomplete` flags:
    // one host sync per step instead of two.  (Overflowed lists hold valid rows, just not all.)
    const bool late_check = !b.host;
    uint32_t flags = 0;
    auto account_scan = [&]() {
      float ms = 0;
      cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]); ctx->timing[0] += ms;
      cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]); ctx->timing[1] += ms;
      ctx->timing[4] += 1;
    };
    auto grow_buffers = [&]() -> int32_t {
      const int64_t limit = std::max<int64_t>(b.n_rows, 1 << 16);
      if ((flags & 1u) && cap >= limit) return fail(SWAT_ERR_OVERFLOW, "class candidate overflow with cap >= n_rows");
      if (flags & 1u) cap = std::min<int64_t>(cap * 4, limit);
      if (flags & 2u) list_entries *= 4;
      ctx->timing[7] += 1;
      return SWAT_OK;
    };
    if (!late_check) {
      SW_OK(job_flags(job, &flags));
      account_scan();
      if (flags & 3u) { SW_OK(grow_buffers()); continue; }
    }
    // ---- exact re-score of the candidates + accept walk
    CU_OK(cudaEventRecord(ctx->ev[3], stream));
    CandLists cl{ctx->w_scores.as<float>(), ctx->w_rows.as<int64_t>(), ctx->w_counts.as<int32_t>(), ctx->w_trunc.as<int32_t>(), kf};
    SW_OK(walk_candidates(ctx, q, b, want_t2i, row_offset, cl, k, thr, t2i_thr, eps, false, d_out_scores, d_out_rows, d_out_t2i,
                          d_out_counts, nullptr, nullptr, stream, nullptr, d_violation, d_status, job->st.flags));
    CU_OK(cudaEventRecord(ctx->ev[4], stream));
    SW_OK(ensure_status(ctx, 2 * static_cast<size_t>(C) + 2));
    CU_OK(cudaMemcpyAsync(ctx->h_status, d_status, (2 * static_cast<size_t>(C) + 2) * 4, cudaMemcpyDeviceToHost, stream));
    CU_OK(cudaStreamSynchronize(stream));
    if (ctx->h_status[1 + 2 * C] != 0)
      return fail(SWAT_ERR_INVALID, "a candidate's exact score differs from the score the scan ranked it by by more than the error bound "
                                    "(%g): bank rows and queries must be L2-normalised cosine features", static_cast<double>(eps));
    const int32_t* inc = ctx->h_status + 1;
    {
      float ms = 0;
      cudaEventElapsedTime(&ms, ctx->ev[3], ctx->ev[4]);
      ctx->timing[2] += ms;
    }
    if (late_check) {
      flags = static_cast<uint32_t>(ctx->h_status[0]);
      account_scan();
      if (flags & 3u) { SW_OK(grow_buffers()); continue; }
    }
    std::vector<int> bad;
    for (int c = 0; c < C; ++c) if (inc[c] != 0) bad.push_back(c);
    if (bad.empty()) break;
    // Some class ran out of trusted candidates before k rows were accepted although more rows were eligible: widen the
    // over-fetch for those classes only, then (T2I walks) enumerate the passers from the image bank, finally the
    // in-pass predicate.
    ctx->timing[7] += 1;
    if (dual) {
      if (k_fetch >= kMaxKFetch)
        return fail(SWAT_ERR_INCOMPLETE, "%d classes not provably exact at the widest over-fetch (%d candidates; more than that many "
                                         "rows tie with the k-th score?)", (int)bad.size(), kMaxKFetch);
      k_fetch = std::min(kMaxKFetch, k_fetch * 2);
      k_class.clear();
      cap = std::max(cap, auto_cap(ctx, k_fetch));
      list_entries = std::max(list_entries, auto_list_entries(ctx, C, k_fetch));
      continue;
    }
    int32_t from = k_fetch;                           // the escalated classes were walked to this depth
    if (!k_class.empty()) { from = kMaxKFetch; for (int c : bad) from = std::min<int32_t>(from, static_cast<int32_t>(k_class[c])); }
    const bool ladder_left = from < kMaxKFetch;
    // A class that accepted p of the d candidates walked so far needs about d*k/p of them.  Where that is beyond the
    // widest over-fetch the ladder would only waste passes: those classes go straight to the bank-swap pass.
    std::vector<int> deeper, few;
    const int32_t* accepted = ctx->h_status + 1 + C;           // copied out before any nested call reuses the buffer
    for (int c : bad) {
      const int64_t d = k_class.empty() ? k_fetch : static_cast<int64_t>(k_class[c]);
      const int64_t p = accepted[c];
      const bool hopeless = p <= 0 || 4 * d * k / p > 5 * kMaxKFetch;   // expected need 25 % beyond the widest over-fetch
      ((hopeless || !ladder_left) && can_swap ? few : deeper).push_back(c);
    }
    if (!few.empty()) {
      if (static_cast<int>(few.size()) == C) {
        std::vector<int> unresolved;
        SW_OK(swap_pass(ctx, q, b, row_offset, k, thr, t2i_thr, d_out_scores, d_out_rows, d_out_t2i, d_out_counts, stream, nullptr, &unresolved));
        if (!unresolved.empty())
          SW_OK(escalate_classes(ctx, q, b, row_offset, k, thr, t2i_thr, unresolved, kForceDual, d_out_scores, d_out_rows, d_out_t2i,
                                 d_out_counts, stream, depth));
        if (depth == 0 && unresolved.empty()) q->all_few_hint = true;     // next call on this query set starts here
      } else {
        SW_OK(escalate_classes(ctx, q, b, row_offset, k, thr, t2i_thr, few, kSwapPass, d_out_scores, d_out_rows, d_out_t2i, d_out_counts,
                               stream, depth));
      }
    }
    if (deeper.empty()) break;
    if (!ladder_left) {
      // widest over-fetch and no bank-swap pass to fall back on
      if (!want_t2i)
        return fail(SWAT_ERR_INCOMPLETE, "%d classes not provably exact at the widest over-fetch (%d candidates; more than that many "
                                         "rows tie with the k-th score?)", (int)deeper.size(), kMaxKFetch);
      if (static_cast<int>(deeper.size()) == C) {
        dual = true;
        k_fetch = default_k_fetch(ctx, k, false, b.host, eps);
        k_class.clear();
        continue;
      }
      SW_OK(escalate_classes(ctx, q, b, row_offset, k, thr, t2i_thr, deeper, kForceDual, d_out_scores, d_out_rows, d_out_t2i, d_out_counts,
                             stream, depth));
      break;
    }
    if (static_cast<int>(deeper.size()) == C) {
      // every class is short: escalate the whole set, x4 per round
      k_class.clear();
      k_fetch = std::min(kMaxKFetch, k_fetch * 4);
      cap = std::max(cap, auto_cap(ctx, k_fetch));
      list_entries = std::max(list_entries, auto_list_entries(ctx, C, k_fetch));
      continue;
    }
    // targeted: a sub-query set of just those classes, twice as deep
    SW_OK(escalate_classes(ctx, q, b, row_offset, k, thr, t2i_thr, deeper, std::min(kMaxKFetch, from * 2), d_out_scores, d_out_rows,
                           d_out_t2i, d_out_counts, stream, depth));
    if (depth == 0 && ctx->overfetch == 0 && q->last_k_fetch > 0) {      // remember the depth that worked, per class
      if (static_cast<int>(q->kclass_hint.size()) != C) q->kclass_hint.assign(C, 0);
      for (int c : deeper) q->kclass_hint[c] = std::max(q->kclass_hint[c], q->last_k_fetch);
    }
    break;
  }
  q->last_k_fetch = dual ? 0 : k_fetch;
  if (depth == 0) {
    CU_OK(cudaEventRecord(ctx->ev[7], stream));
    CU_OK(cudaStreamSynchronize(stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, ev_begin, ctx->ev[7]);
    ctx->timing[3] = ms;
  }
  return SWAT_OK;
}

}  // namespace

// ================================================================================== C-ABI
extern "C" {

int32_t swat_version(void) { return SWAT_VERSION; }
const char* swat_last_error(void) { return g_err.c_str(); }

int32_t swat_ctx_create(int32_t device, swat_ctx** out) {
  if (!out) return fail(SWAT_ERR_INVALID, "out is null");
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(SWAT_ERR_NO_DEVICE, "no CUDA device: swat_b200 has no CPU fallback");
  }
  if (device < 0 || device >= n) return fail(SWAT_ERR_INVALID, "device %d out of range (%d devices)", device, n);
  cudaDeviceProp prop;
  CU_OK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(SWAT_ERR_NO_DEVICE, "device %d is sm_%d%d; swat_b200 is built for sm_100a only", device, prop.major, prop.minor);
  CU_OK(cudaSetDevice(device));
  swat_ctx* ctx = new swat_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  cudaDeviceGetAttribute(&ctx->clock_khz, cudaDevAttrClockRate, device);
  ctx->smem_optin = prop.sharedMemPerBlockOptin;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
    delete ctx;
    return fail(SWAT_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  }
  ctx->encode = reinterpret_cast<EncodeTiledFn>(fn);
  cudaError_t ce = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
  if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&ctx->work_stream, cudaStreamNonBlocking);
  for (auto& ev : ctx->ev) if (ce == cudaSuccess) ce = cudaEventCreate(&ev);
  for (int i = 0; i < 3 && ce == cudaSuccess; ++i) {
    ce = cudaEventCreateWithFlags(&ctx->ev_copied[i], cudaEventDisableTiming);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&ctx->ev_used[i], cudaEventDisableTiming);
  }
  if (ce == cudaSuccess) {
    // optional: without the mapped word the automatic lockstep window simply stays off
    if (cudaHostAlloc(reinterpret_cast<void**>(&ctx->h_probe), 64, cudaHostAllocMapped) == cudaSuccess) {
      *ctx->h_probe = 0;
      if (cudaHostGetDevicePointer(reinterpret_cast<void**>(&ctx->d_probe), ctx->h_probe, 0) != cudaSuccess) {
        cudaFreeHost(ctx->h_probe); ctx->h_probe = nullptr; ctx->d_probe = nullptr;
      }
    } else {
      ctx->h_probe = nullptr;
    }
    (void)cudaGetLastError();
  }
  if (ce != cudaSuccess) {
    swat_ctx_destroy(ctx);        // releases whatever was created
    return fail(SWAT_ERR_CUDA, "context setup failed: %s", cudaGetErrorString(ce));
  }
  *out = ctx;
  return SWAT_OK;
}

int32_t swat_ctx_destroy(swat_ctx* ctx) {
  if (!ctx) return SWAT_OK;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  DevBuf* bufs[] = {&ctx->w_scores, &ctx->w_rows, &ctx->w_counts, &ctx->w_trunc, &ctx->w_exact, &ctx->w_aux, &ctx->w_incomplete, &ctx->w_keys,
                    &ctx->w_stage[0], &ctx->w_stage[1], &ctx->w_stage[2], &ctx->w_rc[0], &ctx->w_rc[1], &ctx->w_rc[2],
                    &ctx->w_ex[0], &ctx->w_ex[1], &ctx->w_ex[2], &ctx->w_img, &ctx->w_idx,
                    &ctx->w_out_scores, &ctx->w_out_rows, &ctx->w_out_t2i, &ctx->w_out_counts, &ctx->w_boot, &ctx->w_progress, &ctx->w_bits, &ctx->w_splice, &ctx->w_tiles,
                    &ctx->w_swap[0], &ctx->w_swap[1], &ctx->w_swap[2], &ctx->w_swap[3], &ctx->w_swap[4], &ctx->w_swap[5], &ctx->w_swap[6],
                    &ctx->w_swap[7], &ctx->w_swap[8], &ctx->w_swap[9]};
  for (DevBuf* b : bufs) b->release();
  if (ctx->cached_job) swat_job_destroy(ctx->cached_job);
  for (int i = 0; i < 2; ++i) if (ctx->esc_q[i]) { swat_queries* e = ctx->esc_q[i]; ctx->esc_q[i] = nullptr; swat_queries_destroy(e); }
  for (auto& lvl : ctx->e_bufs) for (auto& bf : lvl) bf.release();
  for (auto& lvl : ctx->e_remap) for (auto& bf : lvl) bf.release();
  if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
  if (ctx->h_status) cudaFreeHost(ctx->h_status);
  if (ctx->h_probe) cudaFreeHost(ctx->h_probe);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->work_stream) cudaStreamDestroy(ctx->work_stream);
  for (auto& ev : ctx->ev) if (ev) cudaEventDestroy(ev);
  for (int i = 0; i < 3; ++i) { if (ctx->ev_copied[i]) cudaEventDestroy(ctx->ev_copied[i]); if (ctx->ev_used[i]) cudaEventDestroy(ctx->ev_used[i]); }
  delete ctx;
  return SWAT_OK;
}

int32_t swat_ctx_set_option(swat_ctx* ctx, const char* name, int64_t value) {
  if (!ctx || !name) return fail(SWAT_ERR_INVALID, "null argument");
  const std::string n(name);
  if (n == "cta_group") { if (value != 1 && value != 2) return fail(SWAT_ERR_INVALID, "cta_group must be 1 or 2"); ctx->cta_group = (int)value; }
  else if (n == "max_ctas") ctx->max_ctas = static_cast<int>(value);
  else if (n == "cand_cap") ctx->cand_cap = value;
  else if (n == "list_entries") ctx->list_entries = value;
  else if (n == "overfetch") ctx->overfetch = static_cast<int>(value);
  else if (n == "host_chunk_rows") ctx->host_chunk_rows = value;
  else if (n == "unit_plan") ctx->unit_plan = value != 0;
  else if (n == "swap_pass") ctx->swap_pass = value != 0;
  else if (n == "zero_copy") ctx->zero_copy = value != 0;
  else if (n == "f32_op_stages") { if (value < 2 || value > 4) return fail(SWAT_ERR_INVALID, "f32_op_stages must be 2..4"); ctx->f32_op_stages = static_cast<int>(value); }
  else if (n == "lock_window") { ctx->lock_window = static_cast<int>(std::max<int64_t>(-1, value)); ctx->lock_auto_on = false; }
  else if (n == "dyn_tiles") ctx->dyn_tiles = value != 0;
  else if (n == "bootstrap_rows") ctx->bootstrap_rows = std::max<int64_t>(0, value);
  else return fail(SWAT_ERR_INVALID, "unknown option '%s'", name);
  return SWAT_OK;
}

int64_t swat_ctx_launch_count(const swat_ctx* ctx) { return ctx ? ctx->launches : 0; }

int32_t swat_ctx_last_timing(const swat_ctx* ctx, double out[8]) {
  if (!ctx || !out) return fail(SWAT_ERR_INVALID, "null argument");
  for (int i = 0; i < 8; ++i) out[i] = ctx->timing[i];
  return SWAT_OK;
}

int32_t swat_queries_create(swat_ctx* ctx, const float* h_queries, int32_t n_queries, const int32_t* h_class_of_query,
                            int32_t n_classes, int32_t reduce, swat_queries** out) {
  if (!ctx || !h_queries || !out) return fail(SWAT_ERR_INVALID, "null argument");
  if (n_queries < 1 || n_classes < 1) return fail(SWAT_ERR_INVALID, "need at least one query and one class");
  if (reduce < SWAT_REDUCE_NONE || reduce > SWAT_REDUCE_MIN) return fail(SWAT_ERR_INVALID, "bad reduce mode %d", reduce);
  if (!h_class_of_query && n_queries != n_classes) return fail(SWAT_ERR_INVALID, "class_of_query is required when n_queries != n_classes");
  (void)cudaGetLastError();   // drop stale errors left by other libraries in this process
  CU_OK(cudaSetDevice(ctx->device));
  std::vector<int32_t> cb(n_classes + 1, 0);
  for (int i = 0; i < n_queries; ++i) {
    const int c = h_class_of_query ? h_class_of_query[i] : i;
    if (c < 0 || c >= n_classes) return fail(SWAT_ERR_INVALID, "class_of_query[%d]=%d out of range", i, c);
    if (i > 0 && c < (h_class_of_query ? h_class_of_query[i - 1] : i - 1)) return fail(SWAT_ERR_INVALID, "class_of_query must be non-decreasing (queries of one class adjacent)");
    cb[c + 1] += 1;
  }
  for (int c = 0; c < n_classes; ++c) {
    if (cb[c + 1] == 0) return fail(SWAT_ERR_INVALID, "class %d has no query", c);
    if (reduce == SWAT_REDUCE_NONE && cb[c + 1] != 1) return fail(SWAT_ERR_INVALID, "SWAT_REDUCE_NONE needs exactly one query per class (class %d has %d)", c, cb[c + 1]);
    cb[c + 1] += cb[c];
  }
  swat_queries* q = new swat_queries();
  q->ctx = ctx; q->Q = n_queries; q->C = n_classes; q->reduce = reduce; q->class_begin = cb; q->ctas = ctx->cta_group;
  q->h_q.assign(h_queries, h_queries + static_cast<size_t>(n_queries) * kDim);
  const int hard_cols = (q->ctas == 2) ? 256 : 144;
  int max_group = 1;
  for (int c = 0; c < n_classes; ++c) max_group = std::max(max_group, cb[c + 1] - cb[c]);
  // column layout of every block.  Grouped reduces: a class may not straddle the column where the second epilogue
  // warp set starts, which can cost up to max_group-1 padding columns per block -- first try the full width (the
  // padding is often not needed: Q = 256 in groups of 2 is one block), then plan with that slack held back.
  std::vector<int> first;
  std::vector<std::vector<std::pair<int, int>>> place;   // per block: (class, first column)
  std::vector<int32_t> split;
  int widest = 0;
  auto layout = [&](int max_cols) -> bool {
    if (max_cols < max_group || !plan_blocks(cb, max_cols, q->n_qb, q->n_blk, first)) return false;
    place.assign(q->n_qb, {});
    split.assign(q->n_qb, 0);
    widest = 0;
    for (int b = 0; b < q->n_qb; ++b) {
      const int cols = cb[first[b + 1]] - cb[first[b]];
      const int H = (reduce == SWAT_REDUCE_NONE) ? (1 << 30) : ((cols + 1) / 2 + 31) / 32 * 32;
      int col = 0;
      for (int c = first[b]; c < first[b + 1]; ++c) {
        const int R = cb[c + 1] - cb[c];
        if (col < H && col + R > H) col = H;          // never straddle the split
        place[b].push_back({c, col});
        col += R;
      }
      split[b] = H;
      widest = std::max(widest, col);
    }
    return widest <= hard_cols;
  };
  if (!layout(hard_cols) && (reduce == SWAT_REDUCE_NONE || !layout(hard_cols - (max_group - 1)))) {
    delete q;
    return fail(SWAT_ERR_UNSUPPORTED, "a class has %d queries; cannot keep it resident", max_group);
  }
  q->n_blk = std::max(16, (widest + 15) / 16 * 16);
  for (int b = 0; b < q->n_qb; ++b) split[b] = std::min(split[b], q->n_blk);
  q->n_cols = q->n_qb * q->n_blk;
  q->n_stages = tc_pick_stages(q->n_blk, q->ctas, ctx->smem_optin);
  q->n_fstages = tc_pick_fstages(q->n_blk, q->ctas, ctx->smem_optin, &q->n_opstages_f32, ctx->f32_op_stages);
  {
    // fp32 banks are scanned as bf16-rounded rows against bf16-rounded queries.  With x~ = bf16(x) (round to nearest
    // even: |x~_i - x_i| <= 2^-8 |x_i|, 8 significant bits) and q~ = bf16(q):
    //   |x.q - x~.q~| = |x.(q - q~) + (x - x~).q~| <= |x| |q - q~| + 2^-8 |x| |q~|
    // |q - q~| and |q~| are computed here; rows are L2-normalised (extract_mined_feature.py:121,181), 0.1 % slack.
    // Group reduces (mean / max / min over a class's queries) cannot move by more than their worst member.
    double worst = 0.0;
    for (int i = 0; i < n_queries; ++i) {
      double d2 = 0.0, n2 = 0.0;
      for (int j = 0; j < kDim; ++j) {
        const float f = h_queries[static_cast<size_t>(i) * kDim + j];
        const uint32_t bits = static_cast<uint32_t>(f32_to_bf16_rne(f)) << 16;
        float r;
        memcpy(&r, &bits, 4);
        d2 += (static_cast<double>(f) - r) * (static_cast<double>(f) - r);
        n2 += static_cast<double>(r) * r;
      }
      worst = std::max(worst, std::sqrt(d2) + std::sqrt(n2) / 256.0);
    }
    q->eps_conv = static_cast<float>(1.001 * worst) + kEpsAccum;
  }
  // host staging: padded layouts
  const size_t Q = n_queries, NC = q->n_cols;
  std::vector<uint16_t> h_bf(Q * kDim), h_pbf(NC * kDim, 0);
  std::vector<float> h_pf(NC * kDim, 0.0f), h_cnt(NC, 0.0f);
  std::vector<int32_t> h_cls(NC, -1);
  for (size_t i = 0; i < Q * kDim; ++i) h_bf[i] = f32_to_bf16_rne(h_queries[i]);
  for (int b = 0; b < q->n_qb; ++b) {
    for (const auto& pc : place[b]) {
      const int c = pc.first;
      int col = b * q->n_blk + pc.second;
      for (int qi = cb[c]; qi < cb[c + 1]; ++qi, ++col) {
        memcpy(&h_pf[static_cast<size_t>(col) * kDim], &h_queries[static_cast<size_t>(qi) * kDim], kDim * 4);
        memcpy(&h_pbf[static_cast<size_t>(col) * kDim], &h_bf[static_cast<size_t>(qi) * kDim], kDim * 2);
        h_cls[col] = c;
        if (qi == cb[c + 1] - 1) h_cnt[col] = static_cast<float>(cb[c + 1] - cb[c]);
      }
    }
  }
  // one device arena + one staged upload: creating a query set costs one cudaMalloc and one copy
  auto up256 = [](size_t x) { return (x + 255) / 256 * 256; };
  const size_t o_qf = 0, o_qb = o_qf + up256(Q * kDim * 4), o_cb = o_qb + up256(Q * kDim * 2),
               o_pf = o_cb + up256((n_classes + 1) * 4), o_pb = o_pf + up256(NC * kDim * 4), o_cc = o_pb + up256(NC * kDim * 2),
               o_cn = o_cc + up256(NC * 4), o_bc = o_cn + up256(NC * 4), o_bs = o_bc + up256((q->n_qb + 1) * 4),
               total = o_bs + up256(q->n_qb * 4);
  std::vector<char> stage(total, 0);
  memcpy(&stage[o_qf], h_queries, Q * kDim * 4);
  memcpy(&stage[o_qb], h_bf.data(), Q * kDim * 2);
  memcpy(&stage[o_cb], cb.data(), (n_classes + 1) * 4);
  memcpy(&stage[o_pf], h_pf.data(), NC * kDim * 4);
  memcpy(&stage[o_pb], h_pbf.data(), NC * kDim * 2);
  memcpy(&stage[o_cc], h_cls.data(), NC * 4);
  memcpy(&stage[o_cn], h_cnt.data(), NC * 4);
  memcpy(&stage[o_bc], first.data(), (q->n_qb + 1) * 4);
  memcpy(&stage[o_bs], split.data(), q->n_qb * 4);
  cudaError_t e = cudaMalloc(&q->d_arena, total);
  if (e == cudaSuccess) e = cudaMemcpy(q->d_arena, stage.data(), total, cudaMemcpyHostToDevice);
  char* base = static_cast<char*>(q->d_arena);
  q->d_q_f32 = reinterpret_cast<float*>(base + o_qf);
  q->d_q_bf16 = reinterpret_cast<uint16_t*>(base + o_qb);
  q->d_class_begin = reinterpret_cast<int32_t*>(base + o_cb);
  q->d_qp_f32 = reinterpret_cast<float*>(base + o_pf);
  q->d_qp_bf16 = reinterpret_cast<uint16_t*>(base + o_pb);
  q->d_col_class = reinterpret_cast<int32_t*>(base + o_cc);
  q->d_col_count = reinterpret_cast<float*>(base + o_cn);
  q->d_blk_class = reinterpret_cast<int32_t*>(base + o_bc);
  q->d_blk_split = reinterpret_cast<int32_t*>(base + o_bs);
  if (e != cudaSuccess) {
    swat_queries_destroy(q);
    return fail(SWAT_ERR_CUDA, "query upload failed: %s", cudaGetErrorString(e));
  }
  int32_t rc = encode_2d_bf16(ctx, &q->tm_q, q->d_qp_bf16, NC, static_cast<uint32_t>(q->n_blk / q->ctas));
  if (rc != SWAT_OK) { swat_queries_destroy(q); return rc; }
  *out = q;
  return SWAT_OK;
}

int32_t swat_queries_destroy(swat_queries* q) {
  if (!q) return SWAT_OK;
  for (int i = 0; i < 2; ++i) if (q->ctx->esc_parent[i] == q) q->ctx->esc_parent[i] = nullptr;
  cudaSetDevice(q->ctx->device);
  cudaFree(q->d_arena);
  delete q;
  return SWAT_OK;
}

int32_t swat_job_create(swat_ctx* ctx, const swat_queries* q, int32_t k_fetch, float t2t_threshold, swat_job** out) {
  if (!ctx) return fail(SWAT_ERR_INVALID, "null ctx");
  SW_OK(job_create_cap(ctx, q, k_fetch, t2t_threshold, auto_cap(ctx, k_fetch), auto_list_entries(ctx, q->C, k_fetch), out));
  return swat_job_reset(*out, nullptr);
}

int32_t swat_job_reset(swat_job* job, void* stream) {
  if (!job) return fail(SWAT_ERR_INVALID, "null job");
  (void)cudaGetLastError();
  CU_OK(cudaSetDevice(job->ctx->device));
  CU_OK(launch_job_reset(job->st, job->q->C, static_cast<cudaStream_t>(stream)));
  job->fresh = true;
  job->last_stream = static_cast<cudaStream_t>(stream);
  job->ctx->launches += 1;
  return SWAT_OK;
}

int32_t swat_job_set_class_depth(swat_job* job, const int32_t* h_depth, void* stream) {
  if (!job) return fail(SWAT_ERR_INVALID, "null job");
  (void)cudaGetLastError();
  CU_OK(cudaSetDevice(job->ctx->device));
  if (!h_depth) { job->st.k_class = nullptr; return SWAT_OK; }
  const int C = job->q->C;
  std::vector<uint32_t> d(C);
  for (int c = 0; c < C; ++c) {
    if (h_depth[c] < 1 || static_cast<uint32_t>(h_depth[c]) > job->st.k_fetch)
      return fail(SWAT_ERR_INVALID, "class depth %d of class %d outside [1, k_fetch=%u]", h_depth[c], c, job->st.k_fetch);
    d[c] = static_cast<uint32_t>(h_depth[c]);
  }
  if (d != job->h_k_class) {          // steady state: the same depths every step, nothing to upload and no sync
    CU_OK(cudaMemcpyAsync(job->d_k_class, d.data(), static_cast<size_t>(C) * 4, cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream)));
    CU_OK(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));     // d is pageable and dies here
    job->h_k_class = d;
  }
  job->st.k_class = job->d_k_class;
  return SWAT_OK;
}

int32_t swat_job_scan(swat_job* job, const void* d_bank, int32_t dtype, int64_t n_rows, int64_t row_base, const void* d_t2i_bank,
                      float t2i_threshold, const int32_t* d_row_class, const uint32_t* d_exclude, int32_t engine, void* stream) {
  if (!job || (!d_bank && n_rows > 0)) return fail(SWAT_ERR_INVALID, "null argument");
  (void)cudaGetLastError();
  CU_OK(cudaSetDevice(job->ctx->device));
  return scan_view(job, d_bank, dtype, n_rows, row_base, d_t2i_bank, t2i_threshold, d_row_class, d_exclude, engine, nullptr,
                   static_cast<cudaStream_t>(stream));
}

int32_t swat_job_select(swat_job* job, int64_t row_offset, float* d_scores, int64_t* d_rows, int32_t* d_counts, int32_t* d_truncated,
                        void* stream) {
  if (!job || !d_scores || !d_rows || !d_counts) return fail(SWAT_ERR_INVALID, "null argument");
  if (row_offset < 0) return fail(SWAT_ERR_INVALID, "row_offset must be >= 0");
  (void)cudaGetLastError();
  CU_OK(cudaSetDevice(job->ctx->device));
  CU_OK(launch_select(job->st, job->q->C, row_offset, d_scores, d_rows, d_counts, d_truncated, static_cast<cudaStream_t>(stream)));
  job->last_stream = static_cast<cudaStream_t>(stream);
  job->ctx->launches += kSelectLaunches;
  return SWAT_OK;
}

int32_t swat_job_export_flags(swat_job* job, int32_t* d_flags, void* stream) {
  if (!job || !d_flags) return fail(SWAT_ERR_INVALID, "null argument");
  (void)cudaGetLastError();
  CU_OK(cudaSetDevice(job->ctx->device));
  CU_OK(cudaMemcpyAsync(d_flags, job->st.flags, sizeof(int32_t), cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
  return SWAT_OK;
}

int32_t swat_job_status(swat_job* job, int32_t* overflowed) {
  if (!job || !overflowed) return fail(SWAT_ERR_INVALID, "null argument");
  (void)cudaGetLastError();
  CU_OK(cudaSetDevice(job->ctx->device));
  uint32_t flags = 0;
  SW_OK(job_flags(job, &flags));
  *overflowed = static_cast<int32_t>(flags & 3u);   // bit0: class candidates ("cand_cap"), bit1: survivor lists ("list_entries")
  return SWAT_OK;
}

int32_t swat_job_destroy(swat_job* job) {
  if (!job) return SWAT_OK;
  cudaSetDevice(job->ctx->device);
  cudaFree(job->st.tau_enc); cudaFree(job->st.count); cudaFree(job->st.hist); cudaFree(job->st.cand); cudaFree(job->st.flags);
  cudaFree(job->st.list); cudaFree(job->st.list_count); cudaFree(job->d_k_class);
  delete job;
  return SWAT_OK;
}

int32_t swat_scan_eps(const swat_queries* q, int32_t dtype, int32_t engine, float* eps) {
  if (!q || !eps) return fail(SWAT_ERR_INVALID, "null argument");
  if (dtype != SWAT_BF16 && dtype != SWAT_F32) return fail(SWAT_ERR_INVALID, "dtype must be SWAT_BF16 or SWAT_F32");
  if (engine == SWAT_ENGINE_AUTO) engine = resolve_engine(q, dtype, false);
  *eps = scan_eps(q, dtype, engine);
  return SWAT_OK;
}

int32_t swat_rescore_walk(swat_ctx* ctx, const swat_queries* q, const swat_queries* q_aux, const void* d_t2t_bank, const void* d_aux_bank, int32_t dtype,
                          int64_t bank_rows, int64_t bank_row_base, const float* d_cand_scores, const int64_t* d_cand_rows,
                          const int32_t* d_cand_counts, const int32_t* d_truncated, int32_t k_fetch, int32_t k, float t2t_threshold,
                          float aux_threshold, float eps, float* d_out_scores, int64_t* d_out_rows, float* d_out_aux,
                          int32_t* d_out_counts, float* d_out_limit, int32_t* d_incomplete, void* stream) {
  if (!ctx || !q || !d_t2t_bank || !d_cand_scores || !d_cand_rows || !d_cand_counts || !d_out_scores || !d_out_rows || !d_out_counts)
    return fail(SWAT_ERR_INVALID, "null argument");
  if (dtype != SWAT_BF16 && dtype != SWAT_F32) return fail(SWAT_ERR_INVALID, "dtype must be SWAT_BF16 or SWAT_F32");
  if (k_fetch < 1 || k_fetch > kMaxKFetch || k < 1 || k > k_fetch) return fail(SWAT_ERR_INVALID, "need 1 <= k <= k_fetch <= %d", kMaxKFetch);
  if (!(eps >= 0.0f)) return fail(SWAT_ERR_INVALID, "eps must be >= 0");
  if (q_aux && q_aux->C != q->C) return fail(SWAT_ERR_INVALID, "the predicate's query set must cover the same %d classes (has %d)", q->C, q_aux->C);
  (void)cudaGetLastError();   // drop stale errors left by other libraries in this process
  CU_OK(cudaSetDevice(ctx->device));
  BankSrc b;
  b.host = false; b.t2t = d_t2t_bank; b.t2i = d_aux_bank; b.dtype = dtype; b.n_rows = bank_rows;
  CandLists cl{d_cand_scores, d_cand_rows, d_cand_counts, d_truncated, k_fetch};
  return walk_candidates(ctx, q, b, d_aux_bank != nullptr, bank_row_base, cl, k, t2t_threshold, aux_threshold, eps, false, d_out_scores,
                         d_out_rows, d_out_aux, d_out_counts, d_out_limit, d_incomplete, static_cast<cudaStream_t>(stream), q_aux);
}

int32_t swat_merge_topk(swat_ctx* ctx, const float* d_scores, const int64_t* d_rows, const float* d_aux, const int32_t* d_counts,
                        const float* d_limit, int32_t n_shards, int64_t shard_stride_bytes, int32_t n_classes, int32_t k_in,
                        int32_t k_out, float aux_threshold, float* d_out_scores, int64_t* d_out_rows, float* d_out_aux,
                        int32_t* d_out_counts, int32_t* d_incomplete, void* stream) {
  if (!ctx || !d_scores || !d_rows || !d_counts || !d_out_scores || !d_out_rows || !d_out_counts) return fail(SWAT_ERR_INVALID, "null argument");
  if (shard_stride_bytes < 0 || (shard_stride_bytes & 7) != 0) return fail(SWAT_ERR_INVALID, "shard_stride_bytes must be a non-negative multiple of 8");
  if (n_shards < 1 || n_classes < 1 || k_in < 1 || k_out < 1 || k_out > kMaxKFetch)
    return fail(SWAT_ERR_INVALID, "bad merge shape G=%d C=%d k_in=%d k_out=%d", n_shards, n_classes, k_in, k_out);
  (void)cudaGetLastError();
  CU_OK(cudaSetDevice(ctx->device));
  SW_OK(ctx->w_keys.ensure(static_cast<size_t>(n_shards) * n_classes * k_in * 8));
  CU_OK(launch_merge(d_scores, d_rows, d_aux, aux_threshold, d_counts, d_limit, n_shards, shard_stride_bytes, n_classes, k_in, k_out,
                     ctx->w_keys.as<uint64_t>(), d_out_scores, d_out_rows, d_out_aux, d_out_counts, d_incomplete,
                     static_cast<cudaStream_t>(stream)));
  ctx->launches += 2;
  return SWAT_OK;
}

int32_t swat_near_duplicates(swat_ctx* ctx, const void* d_bank, int32_t dtype, int64_t n_rows, const int64_t* d_order,
                             const int32_t* d_class_start, int32_t n_classes, int32_t max_class_rows, float threshold,
                             uint8_t* d_dup, void* stream) {
  if (!ctx || !d_bank || !d_order || !d_class_start || !d_dup) return fail(SWAT_ERR_INVALID, "null argument");
  if (dtype != SWAT_BF16 && dtype != SWAT_F32) return fail(SWAT_ERR_INVALID, "dtype must be SWAT_BF16 or SWAT_F32");
  if (n_rows < 0 || n_classes < 0 || max_class_rows < 0 || n_classes > 65535) return fail(SWAT_ERR_INVALID, "bad shape");
  (void)cudaGetLastError();
  CU_OK(cudaSetDevice(ctx->device));
  CU_OK(launch_near_dup(d_bank, dtype, d_order, d_class_start, n_classes, max_class_rows, threshold, d_dup, static_cast<cudaStream_t>(stream)));
  ctx->launches += 1;
  return SWAT_OK;
}

int32_t swat_scores_dense(swat_ctx* ctx, const swat_queries* q, const void* d_bank, int32_t dtype, int64_t n_rows, float* d_out,
                          int32_t engine, void* stream) {
  if (!ctx || !q || !d_bank || !d_out) return fail(SWAT_ERR_INVALID, "null argument");
  (void)cudaGetLastError();   // drop stale errors left by other libraries in this process
  CU_OK(cudaSetDevice(ctx->device));
  swat_job tmp;   // dense mode never touches job state
  tmp.ctx = ctx; tmp.q = q;
  memset(&tmp.st, 0, sizeof(tmp.st));
  return scan_view(&tmp, d_bank, dtype, n_rows, 0, nullptr, 0.0f, nullptr, nullptr, engine, d_out, static_cast<cudaStream_t>(stream));
}

int32_t swat_score_rows(swat_ctx* ctx, const swat_queries* q, const void* d_bank, int32_t dtype, int64_t n_rows,
                        const int32_t* d_row_class, float* d_out, void* stream) {
  if (!ctx || !q || (n_rows > 0 && (!d_bank || !d_row_class || !d_out))) return fail(SWAT_ERR_INVALID, "null argument");
  if (dtype != SWAT_BF16 && dtype != SWAT_F32) return fail(SWAT_ERR_INVALID, "dtype must be SWAT_BF16 or SWAT_F32");
  if (n_rows < 0) return fail(SWAT_ERR_INVALID, "n_rows must be >= 0");
  (void)cudaGetLastError();
  CU_OK(cudaSetDevice(ctx->device));
  CU_OK(launch_score_rows(d_bank, dtype, d_row_class, n_rows,
                          (dtype == SWAT_BF16) ? static_cast<const void*>(q->d_q_bf16) : static_cast<const void*>(q->d_q_f32),
                          q->d_class_begin, q->C, q->reduce, d_out, static_cast<cudaStream_t>(stream)));
  ctx->launches += 1;
  return SWAT_OK;
}

int32_t swat_zeroshot_predict(swat_ctx* ctx, const swat_queries* q, const void* d_bank, int32_t dtype, int64_t n_rows, int32_t* d_pred,
                              int32_t engine, void* stream) {
  if (!ctx || !q || (!d_bank && n_rows > 0) || (!d_pred && n_rows > 0)) return fail(SWAT_ERR_INVALID, "null argument");
  if (dtype != SWAT_BF16 && dtype != SWAT_F32) return fail(SWAT_ERR_INVALID, "dtype must be SWAT_BF16 or SWAT_F32");
  if (n_rows < 0) return fail(SWAT_ERR_INVALID, "n_rows must be >= 0");
  (void)cudaGetLastError();
  CU_OK(cudaSetDevice(ctx->device));
  const int64_t C = q->C;
  const int64_t chunk = std::max<int64_t>(256, ((256ll << 20) / (4 * C)) / 256 * 256);    // <= 256 MB of scores at a time
  swat_job tmp;   // dense mode never touches job state
  tmp.ctx = ctx; tmp.q = q;
  memset(&tmp.st, 0, sizeof(tmp.st));
  const size_t row_bytes = static_cast<size_t>(kDim) * elem_size(dtype);
  for (int64_t r0 = 0; r0 < n_rows; r0 += chunk) {
    const int64_t n = std::min(chunk, n_rows - r0);
    SW_OK(ctx->w_boot.ensure(static_cast<size_t>(n) * C * 4));
    SW_OK(scan_view(&tmp, static_cast<const char*>(d_bank) + static_cast<size_t>(r0) * row_bytes, dtype, n, 0, nullptr, 0.0f, nullptr, nullptr,
                    engine, ctx->w_boot.as<float>(), static_cast<cudaStream_t>(stream)));
    CU_OK(launch_argmax_rows(ctx->w_boot.as<float>(), n, static_cast<int>(C), d_pred + r0, static_cast<cudaStream_t>(stream)));
    ctx->launches += 1;
  }
  return SWAT_OK;
}

int32_t swat_topk(swat_ctx* ctx, const swat_queries* q, const void* d_t2t_bank, const void* d_t2i_bank, int32_t dtype, int64_t n_rows,
                  int64_t row_offset, int32_t k, float t2t_threshold, float t2i_threshold, const int32_t* d_row_class,
                  const uint32_t* d_exclude, float* d_out_scores, int64_t* d_out_rows, float* d_out_t2i, int32_t* d_out_counts,
                  void* stream) {
  if (!ctx || !q || (!d_t2t_bank && n_rows > 0) || !d_out_scores || !d_out_rows || !d_out_counts) return fail(SWAT_ERR_INVALID, "null argument");
  (void)cudaGetLastError();   // drop stale errors left by other libraries in this process
  CU_OK(cudaSetDevice(ctx->device));
  BankSrc b;
  b.host = false; b.t2t = d_t2t_bank; b.t2i = d_t2i_bank; b.dtype = dtype; b.n_rows = n_rows; b.row_class = d_row_class; b.exclude = d_exclude;
  return run_pipeline(ctx, q, b, row_offset, k, t2t_threshold, t2i_threshold, d_out_scores, d_out_rows, d_out_t2i, d_out_counts,
                      static_cast<cudaStream_t>(stream), 0, 0);
}

int32_t swat_topk_host(swat_ctx* ctx, const swat_queries* q, const void* h_t2t_bank, const void* h_t2i_bank, int32_t dtype, int64_t n_rows,
                       int64_t row_offset, int32_t k, float t2t_threshold, float t2i_threshold, const int32_t* h_row_class,
                       const uint32_t* h_exclude, float* h_out_scores, int64_t* h_out_rows, float* h_out_t2i, int32_t* h_out_counts) {
  if (!ctx || !q || (!h_t2t_bank && n_rows > 0) || !h_out_scores || !h_out_rows || !h_out_counts) return fail(SWAT_ERR_INVALID, "null argument");
  (void)cudaGetLastError();   // drop stale errors left by other libraries in this process
  CU_OK(cudaSetDevice(ctx->device));
  const size_t C = q->C;
  SW_OK(ctx->w_out_t2i.ensure(C * k * 4));
  DevBuf o_scores, o_rows, o_counts;
  int32_t rc = o_scores.ensure(C * k * 4);
  if (rc == SWAT_OK) rc = o_rows.ensure(C * k * 8);
  if (rc == SWAT_OK) rc = o_counts.ensure(C * 4);
  BankSrc b;
  b.host = true; b.t2t = h_t2t_bank; b.t2i = h_t2i_bank; b.dtype = dtype; b.n_rows = n_rows; b.row_class = h_row_class; b.exclude = h_exclude;
  if (ctx->zero_copy) {          // pinned banks are visible to the device: candidates' rows are read in place
    b.t2t_mapped = mapped_alias(h_t2t_bank);
    b.t2i_mapped = mapped_alias(h_t2i_bank);
  }
  cudaStream_t s = ctx->work_stream;
  if (rc == SWAT_OK)
    rc = run_pipeline(ctx, q, b, 0, k, t2t_threshold, t2i_threshold, o_scores.as<float>(), o_rows.as<int64_t>(),
                      (h_out_t2i && h_t2i_bank) ? ctx->w_out_t2i.as<float>() : nullptr, o_counts.as<int32_t>(), s, 0, 0);
  if (rc == SWAT_OK) {
    cudaError_t e = cudaMemcpyAsync(h_out_scores, o_scores.p, C * k * 4, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_out_rows, o_rows.p, C * k * 8, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_out_counts, o_counts.p, C * 4, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess && h_out_t2i && h_t2i_bank) e = cudaMemcpyAsync(h_out_t2i, ctx->w_out_t2i.p, C * k * 4, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) rc = fail(SWAT_ERR_CUDA, "result copy failed: %s", cudaGetErrorString(e));
    ctx->timing[6] += static_cast<double>(C * k * 12 + C * 4 + ((h_out_t2i && h_t2i_bank) ? C * k * 4 : 0));
    if (rc == SWAT_OK && row_offset != 0)
      for (size_t i = 0; i < C * static_cast<size_t>(k); ++i) if (h_out_rows[i] >= 0) h_out_rows[i] += row_offset;
  }
  o_scores.release(); o_rows.release(); o_counts.release();
  return rc;
}

}  // extern "C"
