// fp32 CUDA-core GEMM / column-sum used by the strict-mode node-level projections.
// Replaces the aten::mm / addmm calls of F.multi_head_attention_forward
// (reference era copy src/ampnet/conv/custom_multihead_attn_forward.py:4031-4084, 4436-4437),
// executed once per node token instead of once per edge token.
#include <cuda_bf16.h>

#include "gemm_f32.cuh"

namespace ampconv {
namespace {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

__global__ void __launch_bounds__(NT)
gemm_f32_kernel(const float* __restrict__ A, int64_t sam, int64_t sak,
                const float* __restrict__ B, int64_t sbk, int64_t sbn,
                float* __restrict__ C, int64_t ldc, int64_t M, int64_t N, int64_t K,
                GemmEpilogue epi, int64_t k_per_split, float* __restrict__ partials) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int64_t n0 = (int64_t)blockIdx.y * BN;
  const int64_t k_begin = (int64_t)blockIdx.z * k_per_split;
  const int64_t k_end = min(K, k_begin + k_per_split);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const bool a_kcontig = (sak == 1);
  const bool b_kcontig = (sbk == 1) && (sbn != 1);

  for (int64_t k0 = k_begin; k0 < k_end; k0 += BK) {
#pragma unroll
    for (int i = 0; i < (BM * BK) / NT; ++i) {
      int e = tid + i * NT;
      int kk, m;
      if (a_kcontig) { kk = e & (BK - 1); m = e / BK; } else { m = e & (BM - 1); kk = e / BM; }
      int64_t gm = m0 + m, gk = k0 + kk;
      As[kk][m] = (gm < M && gk < k_end) ? A[gm * sam + gk * sak] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < (BN * BK) / NT; ++i) {
      int e = tid + i * NT;
      int kk, n;
      if (b_kcontig) { kk = e & (BK - 1); n = e / BK; } else { n = e & (BN - 1); kk = e / BN; }
      int64_t gn = n0 + n, gk = k0 + kk;
      Bs[kk][n] = (gn < N && gk < k_end) ? B[gk * sbk + gn * sbn] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      float av[4] = {a.x, a.y, a.z, a.w};
      float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
    float rs = 1.f, gate = 1.f;
    if (partials == nullptr) {
      int64_t g = m / epi.rows_per_group;
      if (epi.row_scale) rs = epi.row_scale[g];
      if (epi.bias_gate) gate = epi.bias_gate[g];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int64_t n = n0 + tx * 4 + j;
      if (n >= N) continue;
      if (partials) {
        partials[((int64_t)blockIdx.z * M + m) * N + n] = acc[i][j];
      } else {
        float v = acc[i][j] * rs;
        if (epi.bias) v += epi.bias[n] * gate;
        if (epi.split_width > 0) {
          const int blk = (int)(n / epi.split_width);
          if (blk == 0) v *= epi.split_scale0;
          reinterpret_cast<__nv_bfloat16*>(epi.split_out[blk])[m * epi.split_width + (n - (int64_t)blk * epi.split_width)] =
              __float2bfloat16(v);
        } else {
          C[m * ldc + n] = v;
        }
      }
    }
  }
}

__global__ void splitk_reduce_kernel(const float* __restrict__ partials, int splits, int64_t MN,
                                     int64_t N, float* __restrict__ C, int64_t ldc) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= MN) return;
  float s = 0.f;
  for (int k = 0; k < splits; ++k) s += partials[(int64_t)k * MN + i];
  C[(i / N) * ldc + (i % N)] = s;
}

__global__ void __launch_bounds__(256)
colsum_partial_kernel(const float* __restrict__ A, int64_t lda, int64_t M, int64_t N,
                      const float* __restrict__ gate, int rows_per_group,
                      float* __restrict__ partials, int64_t rows_per_block) {
  __shared__ float red[256];
  const int tid = threadIdx.x;
  const int64_t r_begin = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r_end = min(M, r_begin + rows_per_block);
  for (int64_t c0 = 0; c0 < N; c0 += 256) {
    const int width = (int)((N - c0) < 256 ? (N - c0) : 256);
    const int lanes = 256 / width;                 // row lanes working on the same column
    const int col = tid % width, lane = tid / width;
    float s = 0.f;
    if (lane < lanes) {
      for (int64_t r = r_begin + lane; r < r_end; r += lanes) {
        float g = gate ? gate[r / rows_per_group] : 1.f;
        s += g * A[r * lda + c0 + col];
      }
    }
    red[tid] = s;
    __syncthreads();
    if (tid < width) {
      float t = 0.f;
      for (int l = 0; l < lanes; ++l) t += red[l * width + tid];
      partials[(int64_t)blockIdx.x * N + c0 + tid] = t;
    }
    __syncthreads();
  }
}

__global__ void colsum_final_kernel(const float* __restrict__ partials, int num_partials, int64_t N,
                                    float* __restrict__ out) {
  int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s = 0.f;
  for (int p = 0; p < num_partials; ++p) s += partials[(int64_t)p * N + n];
  out[n] = s;
}

}  // namespace

int choose_splits(int64_t M, int64_t N, int64_t K) {
  int64_t tiles = ceil_div<int64_t>(M, BM) * ceil_div<int64_t>(N, BN);
  int64_t want = ceil_div<int64_t>(2 * (int64_t)sm_count(), tiles);
  int64_t max_by_k = ceil_div<int64_t>(K, 4 * BK);
  int64_t s = want < max_by_k ? want : max_by_k;
  if (s < 1) s = 1;
  if (s > 1024) s = 1024;
  return (int)s;
}

int gemm_f32(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk, int64_t sbn,
             float* C, int64_t ldc, int64_t M, int64_t N, int64_t K, const GemmEpilogue& epi,
             int splits, float* partials, cudaStream_t stream) {
  if (M <= 0 || N <= 0) return AMPCONV_OK;
  if (splits < 1) splits = 1;
  if (splits > 1 && partials == nullptr) return AMPCONV_ERR_WORKSPACE;
  int64_t k_per_split = ceil_div<int64_t>(ceil_div<int64_t>(K > 0 ? K : 1, splits), BK) * BK;
  dim3 grid((unsigned)ceil_div<int64_t>(M, BM), (unsigned)ceil_div<int64_t>(N, BN), (unsigned)splits);
  gemm_f32_kernel<<<grid, NT, 0, stream>>>(A, sam, sak, B, sbk, sbn, C, ldc, M, N, K, epi, k_per_split,
                                           splits > 1 ? partials : nullptr);
  AMPCONV_CHECK_LAUNCH();
  if (splits > 1) {
    int64_t MN = M * N;
    splitk_reduce_kernel<<<(unsigned)ceil_div<int64_t>(MN, 256), 256, 0, stream>>>(partials, splits, MN, N, C, ldc);
    AMPCONV_CHECK_LAUNCH();
  }
  return AMPCONV_OK;
}

int colsum_f32(const float* A, int64_t lda, int64_t M, int64_t N, const float* gate, int rows_per_group,
               float* out, float* partials, int num_partials, cudaStream_t stream) {
  if (N <= 0) return AMPCONV_OK;
  if (num_partials < 1 || partials == nullptr) return AMPCONV_ERR_WORKSPACE;
  int64_t rows_per_block = ceil_div<int64_t>(M > 0 ? M : 1, num_partials);
  colsum_partial_kernel<<<num_partials, 256, 0, stream>>>(A, lda, M, N, gate, rows_per_group, partials, rows_per_block);
  AMPCONV_CHECK_LAUNCH();
  colsum_final_kernel<<<(unsigned)ceil_div<int64_t>(N, 256), 256, 0, stream>>>(partials, num_partials, N, out);
  AMPCONV_CHECK_LAUNCH();
  return AMPCONV_OK;
}

}  // namespace ampconv
