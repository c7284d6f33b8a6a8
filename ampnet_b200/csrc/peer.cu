// Peer memory over NVLink 5 / NVSwitch for the destination-partitioned path (one process per GPU).
// The reference has no counterpart: its only multi-process code is the unsynchronised gloo demo
// experiments/cora_benchmark_graphsaint_distributed.py:28,63,83.
//
// Design: every rank owns a few cudaMalloc'ed windows (K|V rows of its halo sources, the dK|dV rows its peers return,
// a handful of flag words), exports them as CUDA IPC handles and maps its peers' windows into its own address space.
// Data then moves with plain stream-ordered device-to-device copies INTO THE PEER'S WINDOW -- executed by the copy
// engines over NVLink, so the persistent attention kernels keep every SM -- followed by a 4-byte copy that raises a flag
// in the receiver's window.  The receiver's compute stream waits for the flag with a one-thread kernel (bounded spin on
// ld.acquire.sys) right before the kernel that consumes the rows: the transfer of ring phase t+1 overlaps the math of
// phase t, and nothing on the data path is a collective.
#include <cstring>

#include "common.cuh"

namespace ampconv {
namespace {

// first failure wins, as in the tcgen05 kernels: status = code | (expected & 0xffff) << 16
__global__ void peer_wait_kernel(const int32_t* __restrict__ flag, int32_t expected, int* __restrict__ status,
                                 unsigned long long budget_ns) {
  unsigned long long t0;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
  for (;;) {
    int32_t v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if (v == expected) return;
    unsigned long long t1;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
    if (t1 - t0 > budget_ns) {
      if (status) atomicCAS(status, 0, (int)(601 | ((expected & 0x7fff) << 16)));
      return;
    }
    __nanosleep(200);
  }
}

__global__ void ramp_kernel(int32_t* __restrict__ ramp, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) ramp[i] = i;
}

// rows [send_idx[i], :] of src -> dst[i, :]  (row = row_vec16 * 16 bytes); one warp per 512 bytes of a row
__global__ void gather_rows_kernel(const uint4* __restrict__ src, const int64_t* __restrict__ idx, uint4* __restrict__ dst,
                                   int64_t n_rows, int row_vec16) {
  const int64_t total = n_rows * row_vec16;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / row_vec16;
    const int c = (int)(i - r * row_vec16);
    dst[i] = __ldg(src + idx[r] * row_vec16 + c);
  }
}

}  // namespace
}  // namespace ampconv

using namespace ampconv;

extern "C" int ampconv_peer_alloc(size_t bytes, void** ptr) {
  AMPCONV_REQUIRE(ptr != nullptr);
  *ptr = nullptr;
  if (bytes == 0) bytes = 256;
  AMPCONV_CUDA_TRY(cudaMalloc(ptr, bytes));
  AMPCONV_CUDA_TRY(cudaMemset(*ptr, 0, bytes));
  return AMPCONV_OK;
}

extern "C" int ampconv_peer_free(void* ptr) {
  if (ptr) AMPCONV_CUDA_TRY(cudaFree(ptr));
  return AMPCONV_OK;
}

extern "C" int ampconv_peer_export(void* ptr, void* handle_out) {
  AMPCONV_REQUIRE(ptr && handle_out);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  cudaIpcMemHandle_t h;
  AMPCONV_CUDA_TRY(cudaIpcGetMemHandle(&h, ptr));
  memcpy(handle_out, &h, sizeof(h));
  return AMPCONV_OK;
}

extern "C" int ampconv_peer_open(const void* handle, void** ptr) {
  AMPCONV_REQUIRE(handle && ptr);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  AMPCONV_CUDA_TRY(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return AMPCONV_OK;
}

extern "C" int ampconv_peer_close(void* ptr) {
  if (ptr) AMPCONV_CUDA_TRY(cudaIpcCloseMemHandle(ptr));
  return AMPCONV_OK;
}

extern "C" int ampconv_peer_copy(void* dst, const void* src, size_t bytes, void* stream) {
  if (bytes == 0) return AMPCONV_OK;
  AMPCONV_REQUIRE(dst && src);
  AMPCONV_CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, as_stream(stream)));
  return AMPCONV_OK;
}

extern "C" int ampconv_peer_ramp(int32_t* ramp, int n, void* stream) {
  AMPCONV_REQUIRE(ramp && n > 0);
  ramp_kernel<<<(n + 255) / 256, 256, 0, as_stream(stream)>>>(ramp, n);
  AMPCONV_CHECK_LAUNCH();
  return AMPCONV_OK;
}

extern "C" int ampconv_peer_signal(int32_t* peer_flag, const int32_t* ramp, int32_t value, void* stream) {
  AMPCONV_REQUIRE(peer_flag && ramp && value >= 0 && value < 65536);
  AMPCONV_CUDA_TRY(cudaMemcpyAsync(peer_flag, ramp + value, sizeof(int32_t), cudaMemcpyDeviceToDevice, as_stream(stream)));
  return AMPCONV_OK;
}

extern "C" int ampconv_peer_wait(const int32_t* flag, int32_t expected, void* workspace, double budget_seconds, void* stream) {
  AMPCONV_REQUIRE(flag && budget_seconds > 0);
  int* status = workspace ? reinterpret_cast<int*>(workspace) + 1 : nullptr;
  peer_wait_kernel<<<1, 1, 0, as_stream(stream)>>>(flag, expected, status, (unsigned long long)(budget_seconds * 1e9));
  AMPCONV_CHECK_LAUNCH();
  return AMPCONV_OK;
}

extern "C" int ampconv_gather_rows(const void* src, const int64_t* idx, void* dst, int64_t n_rows, int64_t row_bytes,
                                   void* stream) {
  AMPCONV_REQUIRE(n_rows >= 0 && row_bytes > 0 && row_bytes % 16 == 0);
  if (n_rows == 0) return AMPCONV_OK;
  AMPCONV_REQUIRE(src && idx && dst);
  const int64_t total = n_rows * (row_bytes / 16);
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  gather_rows_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(reinterpret_cast<const uint4*>(src), idx,
                                                               reinterpret_cast<uint4*>(dst), n_rows, (int)(row_bytes / 16));
  AMPCONV_CHECK_LAUNCH();
  return AMPCONV_OK;
}
