// Multi-GPU halo exchange helpers: the rows that come back from the peers in the backward pass are
// partial dK | dV sums for this rank's own source nodes.  They are added into the fp32 gradient in a fixed order
// (sender rank, then position), one thread per output element: deterministic, no atomics.
#include <cuda_bf16.h>

#include "common.cuh"

namespace ampconv {
namespace {

// acc[tgt[i], :] += sum_{j in [rowptr[i], rowptr[i+1])} recv[pos[j], :]     (row = row_vec4 * 4 elements)
// A received row is `tokens` token rows of tok_vec4 float4 each; in acc a token row starts every acc_ld_vec4 float4 at column
// offset acc_col_vec4 (so the sum can land in columns [d, 3d) of a [rows, 3d] gradient tensor as well as in a dense [rows, 2d]).
__global__ void halo_add_bf16_kernel(const uint2* __restrict__ recv, const int32_t* __restrict__ tgt,
                                     const int32_t* __restrict__ rowptr, const int32_t* __restrict__ pos,
                                     float4* __restrict__ acc, int64_t n_tgt, int row_vec4, int tok_vec4, int acc_ld_vec4,
                                     int acc_col_vec4) {
  const int64_t total = n_tgt * row_vec4;
  const int tokens = row_vec4 / tok_vec4;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx / row_vec4;
    const int c = (int)(idx - i * row_vec4);
    const int tok = c / tok_vec4, cc = c - tok * tok_vec4;
    float4* dst = acc + ((int64_t)tgt[i] * tokens + tok) * acc_ld_vec4 + acc_col_vec4 + cc;
    float4 a = *dst;
    for (int j = rowptr[i]; j < rowptr[i + 1]; ++j) {
      const uint2 r = recv[(int64_t)pos[j] * row_vec4 + c];      // four bf16
      a.x += __uint_as_float(r.x << 16);
      a.y += __uint_as_float(r.x & 0xFFFF0000u);
      a.z += __uint_as_float(r.y << 16);
      a.w += __uint_as_float(r.y & 0xFFFF0000u);
    }
    *dst = a;
  }
}

}  // namespace
}  // namespace ampconv

using namespace ampconv;

static int halo_add_impl(const void* recv_bf16, const int32_t* tgt, const int32_t* rowptr, const int32_t* pos, float* acc,
                         int64_t n_tgt, int64_t row_elems, int64_t tok_elems, int64_t acc_ld, int64_t acc_col, void* stream_) {
  AMPCONV_REQUIRE(n_tgt >= 0 && row_elems > 0 && row_elems % 4 == 0);
  AMPCONV_REQUIRE(tok_elems > 0 && tok_elems % 4 == 0 && row_elems % tok_elems == 0 && acc_ld % 4 == 0 && acc_col % 4 == 0 &&
                  acc_col >= 0 && acc_col + tok_elems <= acc_ld);
  if (n_tgt == 0) return AMPCONV_OK;
  AMPCONV_REQUIRE(recv_bf16 && tgt && rowptr && pos && acc);
  const int64_t total = n_tgt * (row_elems / 4);
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  halo_add_bf16_kernel<<<(int)blocks, 256, 0, as_stream(stream_)>>>(
      reinterpret_cast<const uint2*>(recv_bf16), tgt, rowptr, pos, reinterpret_cast<float4*>(acc), n_tgt, (int)(row_elems / 4),
      (int)(tok_elems / 4), (int)(acc_ld / 4), (int)(acc_col / 4));
  AMPCONV_CHECK_LAUNCH();
  return AMPCONV_OK;
}

extern "C" int ampconv_halo_add_bf16(const void* recv_bf16, const int32_t* tgt, const int32_t* rowptr, const int32_t* pos,
                                     float* acc, int64_t n_tgt, int64_t row_elems, void* stream_) {
  return halo_add_impl(recv_bf16, tgt, rowptr, pos, acc, n_tgt, row_elems, row_elems, row_elems, 0, stream_);
}

// Same, into a strided accumulator: a received row holds row_elems / tok_elems token rows of tok_elems elements; token row t
// of node n lives at acc[(n * tokens + t) * acc_ld + acc_col] (e.g. the dK | dV columns [d, 3d) of d_qkv [rows, 3d]).
extern "C" int ampconv_halo_add_bf16_strided(const void* recv_bf16, const int32_t* tgt, const int32_t* rowptr, const int32_t* pos,
                                             float* acc, int64_t n_tgt, int64_t row_elems, int64_t tok_elems, int64_t acc_ld,
                                             int64_t acc_col, void* stream_) {
  return halo_add_impl(recv_bf16, tgt, rowptr, pos, acc, n_tgt, row_elems, tok_elems, acc_ld, acc_col, stream_);
}
