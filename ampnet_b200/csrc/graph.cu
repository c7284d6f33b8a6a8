// Graph preparation: int64 edge_index [2,E] -> destination-sorted and source-sorted CSR views.
// Replaces the per-call gather/scatter bookkeeping of PyG MessagePassing.propagate
// (reference src/ampnet/conv/amp_conv.py:24-26) with structures built once per edge_index.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace ampconv {

thread_local int g_last_cuda_error = 0;
unsigned long long g_launch_count = 0;

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
    cached = n;
  }
  return cached;
}

namespace {

__global__ void split_edges_kernel(const int64_t* __restrict__ edge_index, int64_t E, int64_t N_src, int64_t N_dst,
                                   int32_t* __restrict__ src, int32_t* __restrict__ dst,
                                   int32_t* __restrict__ iota, int* __restrict__ bad) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= E) return;
  int64_t s = edge_index[i], t = edge_index[E + i];
  if (s < 0 || s >= N_src || t < 0 || t >= N_dst) {
    atomicExch(bad, 1);
    s = 0;
    t = 0;
  }
  src[i] = (int32_t)s;
  dst[i] = (int32_t)t;
  iota[i] = (int32_t)i;
}

// rowptr[n] = first slot whose (sorted) key is >= n.
__global__ void rowptr_kernel(const int32_t* __restrict__ sorted_keys, int64_t E, int64_t N,
                              int32_t* __restrict__ rowptr) {
  int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (n > N) return;
  int64_t lo = 0, hi = E;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (sorted_keys[mid] < n) lo = mid + 1; else hi = mid;
  }
  rowptr[n] = (int32_t)lo;
}

__global__ void gather_i32_kernel(const int32_t* __restrict__ table, const int32_t* __restrict__ idx,
                                  int64_t E, int32_t* __restrict__ out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < E) out[i] = table[idx[i]];
}

__global__ void degree_kernel(const int32_t* __restrict__ rowptr, int64_t N,
                              float* __restrict__ inv_deg, float* __restrict__ has_in) {
  int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (n >= N) return;
  int deg = rowptr[n + 1] - rowptr[n];
  inv_deg[n] = 1.0f / (float)(deg > 0 ? deg : 1);
  has_in[n] = deg > 0 ? 1.0f : 0.0f;
}

struct GraphWs {
  int32_t *src, *dst, *iota, *keys_out, *tmp_pos;
  int* bad;
  void* cub_tmp;
  size_t cub_bytes, total;
};

int plan(int64_t E, void* base, GraphWs* ws) {
  size_t cub_bytes = 0;
  cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const int32_t*)nullptr, (int32_t*)nullptr,
                                                  (const int32_t*)nullptr, (int32_t*)nullptr, (int)E);
  if (e != cudaSuccess) return cuda_fail(e);
  size_t n = align_up((size_t)(E > 0 ? E : 1) * sizeof(int32_t), 256);
  char* p = reinterpret_cast<char*>(base);
  size_t off = 0;
  ws->src = (int32_t*)(p + off); off += n;
  ws->dst = (int32_t*)(p + off); off += n;
  ws->iota = (int32_t*)(p + off); off += n;
  ws->keys_out = (int32_t*)(p + off); off += n;
  ws->tmp_pos = (int32_t*)(p + off); off += n;
  ws->bad = (int*)(p + off); off += 256;
  ws->cub_tmp = (void*)(p + off);
  ws->cub_bytes = cub_bytes;
  off += align_up(cub_bytes, 256);
  ws->total = off;
  return AMPCONV_OK;
}

}  // namespace
}  // namespace ampconv

using namespace ampconv;

extern "C" int ampconv_abi_version(void) { return AMPCONV_ABI_VERSION; }

extern "C" const char* ampconv_strerror(int status) {
  switch (status) {
    case AMPCONV_OK: return "ok";
    case AMPCONV_ERR_INVALID_ARGUMENT: return "invalid argument";
    case AMPCONV_ERR_UNSUPPORTED: return "shape not supported by this kernel family";
    case AMPCONV_ERR_INDEX_RANGE: return "edge_index holds a node id outside [0, N)";
    case AMPCONV_ERR_WORKSPACE: return "workspace too small";
    case AMPCONV_ERR_CUDA: return "CUDA runtime error (see ampconv_last_cuda_error)";
    case AMPCONV_ERR_NO_DEVICE: return "no sm_100 CUDA device";
    default: return "unknown status";
  }
}

extern "C" int ampconv_last_cuda_error(void) { return g_last_cuda_error; }

extern "C" uint64_t ampconv_launch_count(void) { return g_launch_count; }

extern "C" int ampconv_device_info(int* sm, int* cc_major, int* cc_minor) {
  AMPCONV_REQUIRE(sm && cc_major && cc_minor);
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return AMPCONV_ERR_NO_DEVICE;
  AMPCONV_CUDA_TRY(cudaDeviceGetAttribute(sm, cudaDevAttrMultiProcessorCount, dev));
  AMPCONV_CUDA_TRY(cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev));
  AMPCONV_CUDA_TRY(cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
  return AMPCONV_OK;
}

extern "C" int ampconv_graph_workspace_bytes(int64_t E, int64_t N, size_t* bytes) {
  AMPCONV_REQUIRE(bytes && E >= 0 && N >= 0 && E < (int64_t)INT32_MAX && N < (int64_t)INT32_MAX);
  GraphWs ws;
  int rc = plan(E, nullptr, &ws);
  if (rc != AMPCONV_OK) return rc;
  *bytes = ws.total;
  return AMPCONV_OK;
}

extern "C" int ampconv_graph_build_bipartite(const int64_t* edge_index, int64_t E, int64_t N, int64_t N_src,
                                   int32_t* dst_rowptr, int32_t* dst_src, int32_t* dst_eid,
                                   int32_t* src_rowptr, int32_t* src_dst, int32_t* src_pos,
                                   float* inv_deg, float* has_in,
                                   void* workspace, size_t workspace_bytes, void* stream_) {
  AMPCONV_REQUIRE(E >= 0 && N >= 0 && N_src >= 0 && E < (int64_t)INT32_MAX && N < (int64_t)INT32_MAX && N_src < (int64_t)INT32_MAX);
  AMPCONV_REQUIRE(dst_rowptr && src_rowptr && workspace && (N == 0 || (inv_deg && has_in)));   // a rank may own no destination
  AMPCONV_REQUIRE(E == 0 || (edge_index && dst_src && dst_eid && src_dst && src_pos));
  cudaStream_t stream = as_stream(stream_);
  GraphWs ws;
  int rc = plan(E, workspace, &ws);
  if (rc != AMPCONV_OK) return rc;
  if (ws.total > workspace_bytes) return AMPCONV_ERR_WORKSPACE;
  const int T = 256;
  AMPCONV_CUDA_TRY(cudaMemsetAsync(ws.bad, 0, sizeof(int), stream));
  int end_bit = 1;
  const int64_t n_max = N > N_src ? N : N_src;
  while ((1ll << end_bit) < (n_max > 1 ? n_max : 2)) ++end_bit;   // keys are < max(N, N_src): sort only the bits in use
  if (E > 0) {
    unsigned gb = (unsigned)ceil_div<int64_t>(E, T);
    split_edges_kernel<<<gb, T, 0, stream>>>(edge_index, E, N_src, N, ws.src, ws.dst, ws.iota, ws.bad);
    AMPCONV_CHECK_LAUNCH();
    // destination-sorted view (stable: ties keep edge_index order)
    size_t cub_bytes = ws.cub_bytes;
    AMPCONV_CUDA_TRY(cub::DeviceRadixSort::SortPairs(ws.cub_tmp, cub_bytes, (const int32_t*)ws.dst, ws.keys_out,
                                                     (const int32_t*)ws.iota, dst_eid, (int)E, 0, end_bit, stream));
    gather_i32_kernel<<<gb, T, 0, stream>>>(ws.src, dst_eid, E, dst_src);
    AMPCONV_CHECK_LAUNCH();
  }
  unsigned gn = (unsigned)ceil_div<int64_t>(N + 1, T);
  rowptr_kernel<<<gn, T, 0, stream>>>(ws.keys_out, E, N, dst_rowptr);
  AMPCONV_CHECK_LAUNCH();
  if (N > 0) {
    degree_kernel<<<(unsigned)ceil_div<int64_t>(N, T), T, 0, stream>>>(dst_rowptr, N, inv_deg, has_in);
    AMPCONV_CHECK_LAUNCH();
  }
  if (E > 0) {
    unsigned gb = (unsigned)ceil_div<int64_t>(E, T);
    // source-sorted view of the destination-sorted slots: keys = dst_src, values = slot
    // ws.dst is reused to hold the sorted destination keys of every slot (keys_out) before it is overwritten
    AMPCONV_CUDA_TRY(cudaMemcpyAsync(ws.dst, ws.keys_out, (size_t)E * sizeof(int32_t), cudaMemcpyDeviceToDevice, stream));
    size_t cub_bytes = ws.cub_bytes;
    AMPCONV_CUDA_TRY(cub::DeviceRadixSort::SortPairs(ws.cub_tmp, cub_bytes, (const int32_t*)dst_src, ws.keys_out,
                                                     (const int32_t*)ws.iota, src_pos, (int)E, 0, end_bit, stream));
    gather_i32_kernel<<<gb, T, 0, stream>>>(ws.dst, src_pos, E, src_dst);
    AMPCONV_CHECK_LAUNCH();
  }
  rowptr_kernel<<<(unsigned)ceil_div<int64_t>(N_src + 1, T), T, 0, stream>>>(ws.keys_out, E, N_src, src_rowptr);
  AMPCONV_CHECK_LAUNCH();
  int bad_host = 0;
  AMPCONV_CUDA_TRY(cudaMemcpyAsync(&bad_host, ws.bad, sizeof(int), cudaMemcpyDeviceToHost, stream));
  AMPCONV_CUDA_TRY(cudaStreamSynchronize(stream));
  return bad_host ? AMPCONV_ERR_INDEX_RANGE : AMPCONV_OK;
}

extern "C" int ampconv_graph_build(const int64_t* edge_index, int64_t E, int64_t N,
                                   int32_t* dst_rowptr, int32_t* dst_src, int32_t* dst_eid,
                                   int32_t* src_rowptr, int32_t* src_dst, int32_t* src_pos,
                                   float* inv_deg, float* has_in,
                                   void* workspace, size_t workspace_bytes, void* stream_) {
  return ampconv_graph_build_bipartite(edge_index, E, N, N, dst_rowptr, dst_src, dst_eid, src_rowptr, src_dst, src_pos,
                                       inv_deg, has_in, workspace, workspace_bytes, stream_);
}
