// SIMT fp32-FMA scan kernel: the exact-arithmetic path.
//
// Used (a) for the DENSE scores of fp32 banks (the S1 primitives t2t_similarity / cal_t2i_similarity and the zero-shot
// logits; the reference's own dtype, utils/extras.py:163 model.float()), accumulated in fp32 in ascending-k order like a
// plain dot product, (b) for swat_job_scan with a second bank: the full reference predicate
// `t2t >= thr and t2i >= t2i_thr` (sample_retrieval.py:511-514) evaluated for every row in one pass, and (c) as the
// on-device checker the tcgen05 kernel is tested against.  The whole-pipeline calls never select with it.
// Same selection epilogue as the tensor-core kernel (epilogue.cuh): thread = bank row.
#include "common.cuh"
#include "epilogue.cuh"
#include "scan_tc.h"

namespace swat {
namespace {

template <typename T> __device__ __forceinline__ float to_f32(T x);
template <> __device__ __forceinline__ float to_f32<float>(float x) { return x; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 x) { return __bfloat162float(x); }

constexpr int kRows = 128;   // rows per CTA = threads per CTA
constexpr int kKc = 32;      // k chunk
constexpr int kNc = 32;      // query columns per pass

template <typename T, int RED, bool PART, bool DUAL, bool DENSE>
__global__ void __launch_bounds__(kRows)
scan_simt_kernel(const ScanArgs a, const T* __restrict__ bank, const T* __restrict__ bank2, const T* __restrict__ queries) {
  __shared__ float s_x[kKc][kRows + 1];
  __shared__ float s_x2[DUAL ? kKc : 1][kRows + 1];
  __shared__ __align__(16) float s_q[kKc][kNc];
  __shared__ float s_tau[kNc];
  __shared__ int32_t s_cls[kNc];
  __shared__ float s_cnt[kNc];
  __shared__ uint32_t s_endmask;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t row0 = static_cast<int64_t>(blockIdx.x) * kRows;
  const int64_t row = row0 + tid;

  EpiCtx cx;
  cx.tau_col = s_tau;
  cx.cls_col = s_cls;
  cx.cnt_col = s_cnt;
  cx.row = static_cast<uint32_t>(row);
  cx.row_valid = row < a.n_rows && !row_excluded(a.exclude, static_cast<uint32_t>(row));
  cx.my_cls = -1;
  if (PART && cx.row_valid) cx.my_cls = a.row_class[row];
  cx.acc = red_init<RED>();
  cx.acc2 = red_init<RED>();
  cx.list_pos = 0;
  // shared lists: slots reserved atomically
  const SlowCtx sc = make_slow_ctx(a, DENSE ? 0u : (blockIdx.x * 4u + static_cast<uint32_t>(warp)) % a.st.n_lists);
  // every CTA refreshes a few class thresholds from the histograms on entry (round-robin over classes)
  if (!DENSE) refresh_tau(a.st, static_cast<int>((blockIdx.x * 4u + static_cast<uint32_t>(warp)) % static_cast<uint32_t>(a.n_classes)));

  for (int c0 = 0; c0 < a.n_cols; c0 += kNc) {
    if (tid < kNc) {
      const int c = c0 + tid;
      const bool in = c < a.n_cols;
      const int cls = in ? a.col_class[c] : -1;
      const float cnt = in ? a.col_count[c] : 0.0f;
      s_cls[tid] = cls;
      s_cnt[tid] = cnt;
      float t = (PART || DUAL) ? __int_as_float(0x7fc00000) : INFINITY;   // never passes (see scan_tc.cu)
      if (!DENSE && cls >= 0 && cnt > 0.0f) t = fast_tau<RED>(f32_dec(ld_cg_u32(&a.st.tau_enc[cls])), cnt);
      s_tau[tid] = t;
      const uint32_t m = __ballot_sync(0xffffffffu, cnt > 0.0f);
      if (tid == 0) s_endmask = m;
    }
    float acc[kNc], acc2[DUAL ? kNc : 1];
#pragma unroll
    for (int j = 0; j < kNc; ++j) acc[j] = 0.0f;
    if (DUAL) {
#pragma unroll
      for (int j = 0; j < kNc; ++j) acc2[DUAL ? j : 0] = 0.0f;
    }
    for (int k0 = 0; k0 < kDim; k0 += kKc) {
      __syncthreads();
      // bank tile, transposed into [k][row]: warp w stages rows w*32 .. w*32+31, lane = k
#pragma unroll 4
      for (int r = 0; r < 32; ++r) {
        const int64_t rg = row0 + warp * 32 + r;
        float x = 0.0f, y = 0.0f;
        if (rg < a.n_rows) {
          x = to_f32<T>(bank[rg * kDim + k0 + lane]);
          if (DUAL) y = to_f32<T>(bank2[rg * kDim + k0 + lane]);
        }
        s_x[lane][warp * 32 + r] = x;
        if (DUAL) s_x2[lane][warp * 32 + r] = y;
      }
      // query tile [k][col]: warp w stages columns w*8 .. w*8+7
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) {
        const int c = c0 + warp * 8 + cc;
        s_q[lane][warp * 8 + cc] = (c < a.n_cols) ? to_f32<T>(queries[static_cast<size_t>(c) * kDim + k0 + lane]) : 0.0f;
      }
      __syncthreads();
#pragma unroll 4
      for (int kk = 0; kk < kKc; ++kk) {
        const float x = s_x[kk][tid];
        const float y = DUAL ? s_x2[kk][tid] : 0.0f;
#pragma unroll
        for (int j4 = 0; j4 < kNc / 4; ++j4) {
          const float4 q = *reinterpret_cast<const float4*>(&s_q[kk][4 * j4]);
          acc[4 * j4 + 0] = fmaf(x, q.x, acc[4 * j4 + 0]);
          acc[4 * j4 + 1] = fmaf(x, q.y, acc[4 * j4 + 1]);
          acc[4 * j4 + 2] = fmaf(x, q.z, acc[4 * j4 + 2]);
          acc[4 * j4 + 3] = fmaf(x, q.w, acc[4 * j4 + 3]);
          if (DUAL) {
            acc2[DUAL ? 4 * j4 + 0 : 0] = fmaf(y, q.x, acc2[DUAL ? 4 * j4 + 0 : 0]);
            acc2[DUAL ? 4 * j4 + 1 : 0] = fmaf(y, q.y, acc2[DUAL ? 4 * j4 + 1 : 0]);
            acc2[DUAL ? 4 * j4 + 2 : 0] = fmaf(y, q.z, acc2[DUAL ? 4 * j4 + 2 : 0]);
            acc2[DUAL ? 4 * j4 + 3 : 0] = fmaf(y, q.w, acc2[DUAL ? 4 * j4 + 3 : 0]);
          }
        }
      }
    }
    const uint32_t endmask = s_endmask;
    if (RED != RED_NONE) {
      // zero-padding columns between Q blocks must not leak into the next class's max / min
#pragma unroll
      for (int j = 0; j < kNc; ++j) {
        if (s_cls[j] < 0) {
          acc[j] = red_init<RED>();
          if (DUAL) acc2[DUAL ? j : 0] = red_init<RED>();
        }
      }
    }
    if constexpr (DUAL) {
      process_chunk<kNc, RED, PART, DUAL, DENSE, true>(a, sc, cx, acc, reinterpret_cast<float(&)[kNc]>(acc2), 0, endmask);
    } else {
      process_chunk<kNc, RED, PART, false, DENSE, true>(a, sc, cx, acc, acc, 0, endmask);
    }
    __syncthreads();   // tables of this column pass are dead only after every warp left the epilogue
  }
}

template <typename T, int RED>
cudaError_t launch_t_red(const ScanArgs& a, const void* bank, const void* bank2, const void* q, bool part, bool dense, cudaStream_t s) {
  const unsigned grid = static_cast<unsigned>((a.n_rows + kRows - 1) / kRows);
  const T* b = static_cast<const T*>(bank);
  const T* b2 = static_cast<const T*>(bank2);
  const T* qq = static_cast<const T*>(q);
  if (grid == 0) return cudaSuccess;
  if (dense) scan_simt_kernel<T, RED, false, false, true><<<grid, kRows, 0, s>>>(a, b, b2, qq);
  else if (b2 && part) scan_simt_kernel<T, RED, true, true, false><<<grid, kRows, 0, s>>>(a, b, b2, qq);
  else if (b2) scan_simt_kernel<T, RED, false, true, false><<<grid, kRows, 0, s>>>(a, b, b2, qq);
  else if (part) scan_simt_kernel<T, RED, true, false, false><<<grid, kRows, 0, s>>>(a, b, b2, qq);
  else scan_simt_kernel<T, RED, false, false, false><<<grid, kRows, 0, s>>>(a, b, b2, qq);
  return cudaGetLastError();
}
template <typename T>
cudaError_t launch_t(const ScanArgs& a, const void* bank, const void* bank2, const void* q, int red, bool part, bool dense, cudaStream_t s) {
  switch (red) {
    case RED_NONE: return launch_t_red<T, RED_NONE>(a, bank, bank2, q, part, dense, s);
    case RED_MEAN: return launch_t_red<T, RED_MEAN>(a, bank, bank2, q, part, dense, s);
    case RED_MAX: return launch_t_red<T, RED_MAX>(a, bank, bank2, q, part, dense, s);
    default: return launch_t_red<T, RED_MIN>(a, bank, bank2, q, part, dense, s);
  }
}

}  // namespace

cudaError_t launch_scan_simt(const ScanArgs& a, const void* bank, const void* bank2, const void* queries_padded,
                             int dtype, int reduce, bool partitioned, bool dense, cudaStream_t stream) {
  if (dtype == 0) return launch_t<__nv_bfloat16>(a, bank, bank2, queries_padded, reduce, partitioned, dense, stream);
  return launch_t<float>(a, bank, bank2, queries_padded, reduce, partitioned, dense, stream);
}

}  // namespace swat
