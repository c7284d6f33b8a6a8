// fp32 CUDA-core GEMM with arbitrary operand strides (strict mode node-level projections and
// their gradients).  C[m,n] = sum_k A(m,k) * B(k,n)  (+ epilogue), A(m,k) = A[m*sam + k*sak],
// B(k,n) = B[k*sbk + n*sbn].  64x64x16 tiles, 256 threads, 4x4 outputs per thread.
#pragma once
#include "common.cuh"

namespace ampconv {

struct GemmEpilogue {
  const float* bias = nullptr;       // [N] added to every row (scaled by bias_gate if given)
  const float* bias_gate = nullptr;  // [M / rows_per_group] multiplies the bias
  const float* row_scale = nullptr;  // [M / rows_per_group] multiplies the whole row
  int rows_per_group = 1;
  // optional bf16 output split into column blocks of `split_width`: block j of row m goes to
  // split_out[j][m * split_width + n % split_width]; block 0 is multiplied by split_scale0.
  void* split_out[3] = {nullptr, nullptr, nullptr};
  int split_width = 0;
  float split_scale0 = 1.f;
};

// Launches the GEMM on `stream`.  If splits > 1 the K range is divided and partial tiles go to
// `partials` ([splits, M, N] floats) followed by a deterministic reduction into C.
int gemm_f32(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk, int64_t sbn,
             float* C, int64_t ldc, int64_t M, int64_t N, int64_t K, const GemmEpilogue& epi,
             int splits, float* partials, cudaStream_t stream);

// out[n] = sum_m gate[m / rows_per_group] * A[m*lda + n]   (gate may be null).  Deterministic two-stage.
int colsum_f32(const float* A, int64_t lda, int64_t M, int64_t N, const float* gate, int rows_per_group,
               float* out, float* partials, int num_partials, cudaStream_t stream);

int choose_splits(int64_t M, int64_t N, int64_t K);

}  // namespace ampconv
