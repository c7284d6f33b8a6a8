// Strict-mode node-level projections and their gradients (C ABI wrappers over gemm_f32).
// in_proj / out_proj of torch.nn.MultiheadAttention (reference era copy
// src/ampnet/conv/custom_multihead_attn_forward.py:4031-4084, 4436-4437) executed once per
// node token; out_proj is applied after the mean aggregation (the mean is linear).
#include "common.cuh"
#include "gemm_f32.cuh"

using namespace ampconv;

namespace {
constexpr int kColsumPartials = 592;   // 4 x 148 blocks
}  // namespace

static size_t ws_split_floats(int out_dim, int in_dim, int splits) { return (size_t)splits * out_dim * in_dim; }

extern "C" int ampconv_param_grad_workspace_bytes(int out_dim, int in_dim, size_t* bytes) {
  AMPCONV_REQUIRE(bytes && out_dim > 0 && in_dim > 0);
  // split-K partials of the weight gradient (choose_splits never exceeds 2*SMs) + column-sum partials
  const size_t max_splits = 2 * 160;
  *bytes = (ws_split_floats(out_dim, in_dim, (int)max_splits) + (size_t)kColsumPartials * out_dim) * sizeof(float);
  return AMPCONV_OK;
}

extern "C" int ampconv_qkv_proj_f32(const float* x, const float* w, const float* b, float* qkv,
                                    int64_t rows, int d, void* stream) {
  AMPCONV_REQUIRE(rows >= 0 && d > 0);
  if (rows == 0) return AMPCONV_OK;
  AMPCONV_REQUIRE(x && w && b && qkv);
  GemmEpilogue epi;
  epi.bias = b;
  // C[r, o] = sum_k x[r,k] * w[o,k]
  return gemm_f32(x, d, 1, w, 1, d, qkv, 3 * (int64_t)d, rows, 3 * (int64_t)d, d, epi, 1, nullptr, as_stream(stream));
}

extern "C" int ampconv_out_proj_f32(const float* agg, const float* w, const float* b, const float* has_in,
                                    float* out, int64_t N, int F, int d, void* stream) {
  AMPCONV_REQUIRE(N >= 0 && F > 0 && d > 0);
  if (N == 0) return AMPCONV_OK;
  AMPCONV_REQUIRE(agg && w && b && has_in && out);
  GemmEpilogue epi;
  epi.bias = b;
  epi.bias_gate = has_in;
  epi.rows_per_group = F;
  return gemm_f32(agg, d, 1, w, 1, d, out, d, N * F, d, d, epi, 1, nullptr, as_stream(stream));
}

static int out_proj_bwd_impl(const float* d_out, const float* agg, const float* w,
                             const float* inv_deg, const float* has_in,
                             float* d_agg, void* d_agg_bf16, float* d_w, float* d_b,
                             int64_t N, int F, int d, void* ws, size_t ws_bytes, void* stream_) {
  AMPCONV_REQUIRE(N >= 0 && F > 0 && d > 0 && d_w && d_b);
  cudaStream_t stream = as_stream(stream_);
  const int64_t rows = N * F;
  if (rows == 0) {
    AMPCONV_CUDA_TRY(cudaMemsetAsync(d_w, 0, sizeof(float) * d * d, stream));
    AMPCONV_CUDA_TRY(cudaMemsetAsync(d_b, 0, sizeof(float) * d, stream));
    return AMPCONV_OK;
  }
  AMPCONV_REQUIRE(d_out && agg && w && inv_deg && has_in && (d_agg || d_agg_bf16) && ws);
  const int splits = choose_splits(d, d, rows);
  const size_t need = (ws_split_floats(d, d, splits) + (size_t)kColsumPartials * d) * sizeof(float);
  if (need > ws_bytes) return AMPCONV_ERR_WORKSPACE;
  float* partials = reinterpret_cast<float*>(ws);
  float* col_partials = partials + ws_split_floats(d, d, splits);
  // d_agg[r,k] = inv_deg[node] * sum_o d_out[r,o] * w[o,k]
  GemmEpilogue epi;
  epi.row_scale = inv_deg;
  epi.rows_per_group = F;
  if (d_agg_bf16) {
    epi.split_out[0] = d_agg_bf16;
    epi.split_width = d;
  }
  int rc = gemm_f32(d_out, d, 1, w, d, 1, d_agg, d, rows, d, d, epi, 1, nullptr, stream);
  if (rc != AMPCONV_OK) return rc;
  // d_w[o,k] = sum_r d_out[r,o] * agg[r,k]
  rc = gemm_f32(d_out, 1, d, agg, d, 1, d_w, d, d, d, rows, GemmEpilogue(), splits, partials, stream);
  if (rc != AMPCONV_OK) return rc;
  // d_b[o] = sum over rows whose node has an in-edge
  return colsum_f32(d_out, d, rows, d, has_in, F, d_b, col_partials, kColsumPartials, stream);
}

extern "C" int ampconv_out_proj_bwd_f32(const float* d_out, const float* agg, const float* w,
                                        const float* inv_deg, const float* has_in,
                                        float* d_agg, float* d_w, float* d_b,
                                        int64_t N, int F, int d, void* ws, size_t ws_bytes, void* stream_) {
  return out_proj_bwd_impl(d_out, agg, w, inv_deg, has_in, d_agg, nullptr, d_w, d_b, N, F, d, ws, ws_bytes, stream_);
}

extern "C" int ampconv_out_proj_bwd_bf16(const float* d_out, const float* agg, const float* w,
                                         const float* inv_deg, const float* has_in,
                                         void* d_agg_bf16, float* d_w, float* d_b,
                                         int64_t N, int F, int d, void* ws, size_t ws_bytes, void* stream_) {
  return out_proj_bwd_impl(d_out, agg, w, inv_deg, has_in, nullptr, d_agg_bf16, d_w, d_b, N, F, d, ws, ws_bytes, stream_);
}

extern "C" int ampconv_qkv_proj_bwd_f32(const float* x, const float* d_qkv, const float* w,
                                        float* d_x, float* d_w, float* d_b, int64_t rows, int d,
                                        void* ws, size_t ws_bytes, void* stream_) {
  AMPCONV_REQUIRE(rows >= 0 && d > 0 && d_w && d_b);
  cudaStream_t stream = as_stream(stream_);
  const int64_t d3 = 3 * (int64_t)d;
  if (rows == 0) {
    AMPCONV_CUDA_TRY(cudaMemsetAsync(d_w, 0, sizeof(float) * d3 * d, stream));
    AMPCONV_CUDA_TRY(cudaMemsetAsync(d_b, 0, sizeof(float) * d3, stream));
    return AMPCONV_OK;
  }
  AMPCONV_REQUIRE(x && d_qkv && w && d_x && ws);
  const int splits = choose_splits(d3, d, rows);
  const size_t need = (ws_split_floats((int)d3, d, splits) + (size_t)kColsumPartials * d3) * sizeof(float);
  if (need > ws_bytes) return AMPCONV_ERR_WORKSPACE;
  float* partials = reinterpret_cast<float*>(ws);
  float* col_partials = partials + ws_split_floats((int)d3, d, splits);
  // d_x[r,k] = sum_o d_qkv[r,o] * w[o,k]
  int rc = gemm_f32(d_qkv, d3, 1, w, d, 1, d_x, d, rows, d, d3, GemmEpilogue(), 1, nullptr, stream);
  if (rc != AMPCONV_OK) return rc;
  // d_w[o,k] = sum_r d_qkv[r,o] * x[r,k]
  rc = gemm_f32(d_qkv, 1, d3, x, d, 1, d_w, d, d3, d, rows, GemmEpilogue(), splits, partials, stream);
  if (rc != AMPCONV_OK) return rc;
  return colsum_f32(d_qkv, d3, rows, d3, nullptr, 1, d_b, col_partials, kColsumPartials, stream);
}

// bf16 family: Q' = (x Wq^T + bq) * q_scale, K, V as three bf16 [rows, d] tensors (node-major tiles for TMA).
extern "C" int ampconv_qkv_proj_bf16(const float* x, const float* w, const float* b, void* q, void* k, void* v,
                                     int64_t rows, int d, float q_scale, void* stream) {
  AMPCONV_REQUIRE(rows >= 0 && d > 0);
  if (rows == 0) return AMPCONV_OK;
  AMPCONV_REQUIRE(x && w && b && q && k && v);
  GemmEpilogue epi;
  epi.bias = b;
  epi.split_out[0] = q;
  epi.split_out[1] = k;
  epi.split_out[2] = v;
  epi.split_width = d;
  epi.split_scale0 = q_scale;
  return gemm_f32(x, d, 1, w, 1, d, nullptr, 0, rows, 3 * (int64_t)d, d, epi, 1, nullptr, as_stream(stream));
}

// Parameter gradients only (the bf16 family computes the input gradients with the tcgen05 kernels of linear_tc.cu).
extern "C" int ampconv_out_proj_bwd_params_f32(const float* d_out, const float* agg, const float* has_in,
                                               float* d_w, float* d_b, int64_t N, int F, int d,
                                               void* ws, size_t ws_bytes, void* stream_) {
  AMPCONV_REQUIRE(N >= 0 && F > 0 && d > 0 && d_w && d_b);
  cudaStream_t stream = as_stream(stream_);
  const int64_t rows = N * F;
  if (rows == 0) {
    AMPCONV_CUDA_TRY(cudaMemsetAsync(d_w, 0, sizeof(float) * d * d, stream));
    AMPCONV_CUDA_TRY(cudaMemsetAsync(d_b, 0, sizeof(float) * d, stream));
    return AMPCONV_OK;
  }
  AMPCONV_REQUIRE(d_out && agg && has_in && ws);
  const int splits = choose_splits(d, d, rows);
  const size_t need = (ws_split_floats(d, d, splits) + (size_t)kColsumPartials * d) * sizeof(float);
  if (need > ws_bytes) return AMPCONV_ERR_WORKSPACE;
  float* partials = reinterpret_cast<float*>(ws);
  float* col_partials = partials + ws_split_floats(d, d, splits);
  int rc = gemm_f32(d_out, 1, d, agg, d, 1, d_w, d, d, d, rows, GemmEpilogue(), splits, partials, stream);
  if (rc != AMPCONV_OK) return rc;
  return colsum_f32(d_out, d, rows, d, has_in, F, d_b, col_partials, kColsumPartials, stream);
}

extern "C" int ampconv_qkv_proj_bwd_params_f32(const float* x, const float* d_qkv, float* d_w, float* d_b,
                                               int64_t rows, int d, void* ws, size_t ws_bytes, void* stream_) {
  AMPCONV_REQUIRE(rows >= 0 && d > 0 && d_w && d_b);
  cudaStream_t stream = as_stream(stream_);
  const int64_t d3 = 3 * (int64_t)d;
  if (rows == 0) {
    AMPCONV_CUDA_TRY(cudaMemsetAsync(d_w, 0, sizeof(float) * d3 * d, stream));
    AMPCONV_CUDA_TRY(cudaMemsetAsync(d_b, 0, sizeof(float) * d3, stream));
    return AMPCONV_OK;
  }
  AMPCONV_REQUIRE(x && d_qkv && ws);
  const int splits = choose_splits(d3, d, rows);
  const size_t need = (ws_split_floats((int)d3, d, splits) + (size_t)kColsumPartials * d3) * sizeof(float);
  if (need > ws_bytes) return AMPCONV_ERR_WORKSPACE;
  float* partials = reinterpret_cast<float*>(ws);
  float* col_partials = partials + ws_split_floats((int)d3, d, splits);
  int rc = gemm_f32(d_qkv, 1, d3, x, d, 1, d_w, d, d3, d, rows, GemmEpilogue(), splits, partials, stream);
  if (rc != AMPCONV_OK) return rc;
  return colsum_f32(d_qkv, d3, rows, d3, nullptr, 1, d_b, col_partials, kColsumPartials, stream);
}
