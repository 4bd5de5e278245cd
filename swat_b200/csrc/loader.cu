// Feature-shard loader of the C-ABI: rows of a flat shard file -> device memory.
//
// Replaces, for the bank matrices, torch.load(pre_extracted_feats_fn) + .cuda() of
// /root/reference/retrieval/sample_retrieval.py:1473-1476, :337, :399 (a pickled dict that is read, unpickled and
// copied row block by row block through pageable memory).  The flat shard (swat_b200/shards.py: raw row-major
// [n_rows,512] bf16 | f32 files) is read straight into the bank's place in HBM:
//   * pread() into two pinned staging buffers, the read of chunk i+1 overlapping the H2D copy of chunk i (default);
//   * GPUDirect Storage when SWAT_GDS=1 is set and libcufile loads and accepts the file (cuFileRead into device
//     memory, no host staging).
#include <cuda_runtime.h>
#include <cufile.h>
#include <dlfcn.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/swat_b200.h"
#include "common.cuh"

namespace swat {
int32_t api_fail(int32_t code, const char* fmt, ...);      // api.cu: sets swat_last_error()
int api_ctx_device(const swat_ctx* ctx);

namespace {

struct CuFileApi {
  void* lib = nullptr;
  CUfileError_t (*driver_open)() = nullptr;
  CUfileError_t (*handle_register)(CUfileHandle_t*, CUfileDescr_t*) = nullptr;
  void (*handle_deregister)(CUfileHandle_t) = nullptr;
  ssize_t (*read)(CUfileHandle_t, void*, size_t, off_t, off_t) = nullptr;
  bool ok = false, tried = false;
};

CuFileApi& cufile() {
  static CuFileApi api;
  if (api.tried) return api;
  api.tried = true;
  // Opt-in: on hosts without the nvidia-fs kernel module cuFileDriverOpen / cuFileRead can block for minutes in their
  // compatibility mode, so GPUDirect Storage is used only when the operator asks for it (SWAT_GDS=1).
  const char* want = getenv("SWAT_GDS");
  if (!want || want[0] == '0') return api;
  api.lib = dlopen("libcufile.so.0", RTLD_NOW | RTLD_LOCAL);
  if (!api.lib) return api;
  api.driver_open = reinterpret_cast<decltype(api.driver_open)>(dlsym(api.lib, "cuFileDriverOpen"));
  api.handle_register = reinterpret_cast<decltype(api.handle_register)>(dlsym(api.lib, "cuFileHandleRegister"));
  api.handle_deregister = reinterpret_cast<decltype(api.handle_deregister)>(dlsym(api.lib, "cuFileHandleDeregister"));
  api.read = reinterpret_cast<decltype(api.read)>(dlsym(api.lib, "cuFileRead"));
  if (!api.driver_open || !api.handle_register || !api.handle_deregister || !api.read) return api;
  api.ok = api.driver_open().err == CU_FILE_SUCCESS;
  return api;
}

// GPUDirect Storage path; false = not available for this file (the caller falls back to the staged path)
bool load_gds(const char* path, size_t file_off, size_t bytes, void* d_dst) {
  CuFileApi& api = cufile();
  if (!api.ok) return false;
  const int fd = open(path, O_RDONLY | O_DIRECT);
  if (fd < 0) return false;
  CUfileDescr_t descr;
  memset(&descr, 0, sizeof(descr));
  descr.handle.fd = fd;
  descr.type = CU_FILE_HANDLE_TYPE_OPAQUE_FD;
  CUfileHandle_t fh;
  bool done = false;
  if (api.handle_register(&fh, &descr).err == CU_FILE_SUCCESS) {
    done = true;
    const size_t piece = size_t(64) << 20;
    for (size_t o = 0; o < bytes && done; o += piece) {
      const size_t n = std::min(piece, bytes - o);
      const ssize_t r = api.read(fh, d_dst, n, static_cast<off_t>(file_off + o), static_cast<off_t>(o));
      if (r != static_cast<ssize_t>(n)) done = false;
    }
    api.handle_deregister(fh);
  }
  close(fd);
  return done;
}

}  // namespace
}  // namespace swat

using namespace swat;

extern "C" int32_t swat_bank_load(swat_ctx* ctx, const char* path, int32_t dtype, int64_t row_begin, int64_t row_end, void* d_dst,
                                  int64_t chunk_rows, int32_t* used_gds, void* stream_) {
  if (!ctx || !path || (!d_dst && row_end > row_begin)) return api_fail(SWAT_ERR_INVALID, "null argument");
  if (dtype != SWAT_BF16 && dtype != SWAT_F32) return api_fail(SWAT_ERR_INVALID, "dtype must be SWAT_BF16 or SWAT_F32");
  if (row_begin < 0 || row_end < row_begin) return api_fail(SWAT_ERR_INVALID, "bad row range [%lld, %lld)", (long long)row_begin, (long long)row_end);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  (void)cudaGetLastError();
  if (cudaSetDevice(api_ctx_device(ctx)) != cudaSuccess) return api_fail(SWAT_ERR_CUDA, "cudaSetDevice failed");
  if (used_gds) *used_gds = 0;
  const size_t row_bytes = static_cast<size_t>(kDim) * (dtype == SWAT_BF16 ? 2 : 4);
  const size_t off = static_cast<size_t>(row_begin) * row_bytes, bytes = static_cast<size_t>(row_end - row_begin) * row_bytes;
  struct stat st;
  if (stat(path, &st) != 0) return api_fail(SWAT_ERR_INVALID, "cannot stat %s: %s", path, strerror(errno));
  if (static_cast<size_t>(st.st_size) < off + bytes)
    return api_fail(SWAT_ERR_INVALID, "%s holds %lld bytes, rows [%lld, %lld) need %zu", path, (long long)st.st_size, (long long)row_begin,
                    (long long)row_end, off + bytes);
  if (bytes == 0) return SWAT_OK;
  if (load_gds(path, off, bytes, d_dst)) {
    if (used_gds) *used_gds = 1;
    return SWAT_OK;
  }
  // staged path: two pinned buffers, pread of the next chunk overlaps the H2D copy of the previous one
  const size_t chunk = std::max<size_t>(row_bytes, static_cast<size_t>(chunk_rows > 0 ? chunk_rows : (1 << 16)) * row_bytes);
  const int fd = open(path, O_RDONLY);
  if (fd < 0) return api_fail(SWAT_ERR_INVALID, "cannot open %s: %s", path, strerror(errno));
  posix_fadvise(fd, static_cast<off_t>(off), static_cast<off_t>(bytes), POSIX_FADV_SEQUENTIAL);
  char* stage[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  int32_t rc = SWAT_OK;
  for (int i = 0; i < 2 && rc == SWAT_OK; ++i) {
    if (cudaMallocHost(reinterpret_cast<void**>(&stage[i]), std::min(chunk, bytes)) != cudaSuccess ||
        cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming) != cudaSuccess)
      rc = api_fail(SWAT_ERR_CUDA, "staging buffers: %s", cudaGetErrorString(cudaGetLastError()));
  }
  bool busy[2] = {false, false};
  int b = 0;
  for (size_t o = 0; o < bytes && rc == SWAT_OK; o += chunk, b ^= 1) {
    const size_t n = std::min(chunk, bytes - o);
    if (busy[b] && cudaEventSynchronize(ev[b]) != cudaSuccess) { rc = api_fail(SWAT_ERR_CUDA, "event sync failed"); break; }
    // the page cache -> pinned copy is a memcpy in the kernel: a few reader threads per chunk reach the memory bandwidth
    const unsigned nt = n >= (size_t(8) << 20) ? std::max(1u, std::min(8u, std::thread::hardware_concurrency())) : 1u;
    const size_t seg = (n / nt + 4095) / 4096 * 4096;
    std::vector<int> err(nt, 0);
    auto reader = [&](unsigned t) {
      const size_t lo = std::min(n, static_cast<size_t>(t) * seg), hi = (t + 1 == nt) ? n : std::min(n, lo + seg);
      size_t got = lo;
      while (got < hi) {
        const ssize_t r = pread(fd, stage[b] + got, hi - got, static_cast<off_t>(off + o + got));
        if (r <= 0) { err[t] = r == 0 ? -1 : errno; return; }
        got += static_cast<size_t>(r);
      }
    };
    if (nt == 1) reader(0);
    else {
      std::vector<std::thread> th;
      for (unsigned t = 0; t < nt; ++t) th.emplace_back(reader, t);
      for (auto& x : th) x.join();
    }
    for (unsigned t = 0; t < nt && rc == SWAT_OK; ++t)
      if (err[t] != 0) rc = api_fail(SWAT_ERR_INVALID, "read of %s failed near offset %zu: %s", path, off + o, err[t] < 0 ? "unexpected end of file" : strerror(err[t]));
    if (rc != SWAT_OK) break;
    if (cudaMemcpyAsync(static_cast<char*>(d_dst) + o, stage[b], n, cudaMemcpyHostToDevice, stream) != cudaSuccess ||
        cudaEventRecord(ev[b], stream) != cudaSuccess)
      rc = api_fail(SWAT_ERR_CUDA, "H2D copy failed: %s", cudaGetErrorString(cudaGetLastError()));
    busy[b] = true;
  }
  if (cudaStreamSynchronize(stream) != cudaSuccess && rc == SWAT_OK) rc = api_fail(SWAT_ERR_CUDA, "stream sync failed");
  for (int i = 0; i < 2; ++i) {
    if (ev[i]) cudaEventDestroy(ev[i]);
    if (stage[i]) cudaFreeHost(stage[i]);
  }
  close(fd);
  return rc;
}
