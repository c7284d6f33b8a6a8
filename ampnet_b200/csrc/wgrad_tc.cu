// Parameter gradients of the two projections on tcgen05 (bf16 family):
//   d_w[a, b] = sum_r A[r, a] * B[r, b],   d_b[a] = sum_r gate(r) * A[r, a]
// with A = upstream gradient rows (d_out [rows,64] or d_qkv [rows,192]) and B = layer input rows
// (agg or x, [rows,64]); rows = N*F is tens of millions, so this is an HBM-bound reduction.
// (autograd of the aten::addmm calls in custom_multihead_attn_forward.py:4031-4084, 4436-4437.)
//
// Both operands are "row = contraction index" tiles, i.e. MN-major operands for tcgen05: a 128-row tile
// converted to bf16 and stored row-major with the 128B swizzle serves as A (M = gradient columns) and as
// B (N = input columns + one extra 16-column group holding gate(r), which yields the bias gradient from
// the same MMA).  Each persistent CTA accumulates its share of the row tiles in TMEM and writes one
// partial [MW, 64 + 1] block; a tiny second kernel reduces the partials deterministically.
#include <cuda_bf16.h>

#include "common.cuh"
#include "umma.cuh"

namespace ampconv {
namespace {

using namespace umma;

constexpr int kThreads = 192;       // 4 loader/epilogue warps + MMA warp + spare
constexpr int kAtomBytes = 128 * 128;

template <int MW>
struct WgSmem {
  static constexpr int NA = MW / 64;              // A atoms (64 gradient columns each)
  static constexpr int NS = MW == 64 ? 3 : 2;
  uint8_t a[NS][NA][kAtomBytes];
  uint8_t b[NS][2][kAtomBytes];                   // [x tile, gate tile (first 16 columns used)]
  uint64_t full[NS], empty[NS], done;
  uint32_t tmem_base;
};

// A_BF16: the gradient rows A are bf16 already (ampconv_attn_bwd_*_bf16_h): 16-byte chunks are copied as they are.
template <int MW, bool A_BF16 = false>
__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc_kernel(const float* __restrict__ A, const float* __restrict__ B, const float* __restrict__ gate,
                int tokens_per_node, int64_t rows, float* __restrict__ partials, int* __restrict__ status) {
  using Smem = WgSmem<MW>;
  constexpr int NA = Smem::NA, NS = Smem::NS;
  constexpr int NB = 80;                           // 64 input columns + 16 gate columns
  extern __shared__ uint8_t smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;

  if (warp == 4) tmem_alloc(&sm.tmem_base, 256);
  if (tid == 0) {
    for (int i = 0; i < NS; ++i) {
      mbar_init(&sm.full[i], 4);
      mbar_init(&sm.empty[i], 1);
    }
    mbar_init(&sm.done, 1);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;
  const int64_t num_tiles = (rows + 127) / 128;

  if (warp < 4) {
    // ------------------------------------------------------------------ loaders (then epilogue)
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const uint32_t st = it % NS;
      if (!mbar_wait(&sm.empty[st], ((it / NS) & 1) ^ 1)) { atomicCAS(status, 0, 501 | (blockIdx.x << 16)); break; }
      const int64_t row0 = tile * 128;
      // A tile: 128 rows x MW values -> NA atoms
      if constexpr (A_BF16) {
        const uint4* Ab = reinterpret_cast<const uint4*>(A);
#pragma unroll
        for (int base = 0; base < MW / 8; base += 8) {
          uint4 v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int q = (base + u) * 128 + tid;
            const int r = q / (MW / 8), c = q - r * (MW / 8);
            const int64_t grow = row0 + r;
            v[u] = grow < rows ? __ldg(Ab + grow * (MW / 8) + c) : make_uint4(0u, 0u, 0u, 0u);
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int q = (base + u) * 128 + tid;
            const int r = q / (MW / 8), c = q - r * (MW / 8);
            *reinterpret_cast<uint4*>(sm.a[st][c >> 3] + sw128_offset(r, (c & 7) * 16)) = v[u];
          }
        }
      } else {
#pragma unroll
      for (int base = 0; base < MW / 8; base += 8) {
        float4 v[8][2];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int q = (base + u) * 128 + tid;
          const int r = q / (MW / 8), c = q - r * (MW / 8);
          const int64_t grow = row0 + r;
          if (grow < rows) {
            const float4* src = reinterpret_cast<const float4*>(A + grow * MW + c * 8);
            v[u][0] = __ldg(src);
            v[u][1] = __ldg(src + 1);
          } else {
            v[u][0] = make_float4(0.f, 0.f, 0.f, 0.f);
            v[u][1] = v[u][0];
          }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int q = (base + u) * 128 + tid;
          const int r = q / (MW / 8), c = q - r * (MW / 8);
          uint4 pk;
          pk.x = pack_bf16x2(v[u][0].x, v[u][0].y);
          pk.y = pack_bf16x2(v[u][0].z, v[u][0].w);
          pk.z = pack_bf16x2(v[u][1].x, v[u][1].y);
          pk.w = pack_bf16x2(v[u][1].z, v[u][1].w);
          *reinterpret_cast<uint4*>(sm.a[st][c >> 3] + sw128_offset(r, (c & 7) * 16)) = pk;
        }
      }
      }
      // B tile: 128 rows x 64 floats
      {
        float4 v[8][2];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int q = u * 128 + tid;
          const int r = q >> 3, c = q & 7;
          const int64_t grow = row0 + r;
          if (grow < rows) {
            const float4* src = reinterpret_cast<const float4*>(B + grow * 64 + c * 8);
            v[u][0] = __ldg(src);
            v[u][1] = __ldg(src + 1);
          } else {
            v[u][0] = make_float4(0.f, 0.f, 0.f, 0.f);
            v[u][1] = v[u][0];
          }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int q = u * 128 + tid;
          const int r = q >> 3, c = q & 7;
          uint4 pk;
          pk.x = pack_bf16x2(v[u][0].x, v[u][0].y);
          pk.y = pack_bf16x2(v[u][0].z, v[u][0].w);
          pk.z = pack_bf16x2(v[u][1].x, v[u][1].y);
          pk.w = pack_bf16x2(v[u][1].z, v[u][1].w);
          *reinterpret_cast<uint4*>(sm.b[st][0] + sw128_offset(r, c * 16)) = pk;
        }
      }
      // gate tile: row r carries gate(r) in its first 16 columns (two 16-byte chunks)
      {
        const int r = tid;
        const int64_t grow = row0 + r;
        float gv = 0.f;
        if (grow < rows) gv = gate ? gate[grow / tokens_per_node] : 1.f;
        const uint32_t g2 = pack_bf16x2(gv, gv);
        const uint4 pk = make_uint4(g2, g2, g2, g2);
        *reinterpret_cast<uint4*>(sm.b[st][1] + sw128_offset(r, 0)) = pk;
        *reinterpret_cast<uint4*>(sm.b[st][1] + sw128_offset(r, 16)) = pk;
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.full[st]);
    }
    // ------------------------------------------------------------------ epilogue: partial block of this CTA
    if (mbar_wait(&sm.done, 0)) {
      tc_fence_after();
      const int row = warp * 32 + lane;               // TMEM lane = gradient column (within a 128-row MMA)
      const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
      float* out = partials + (int64_t)blockIdx.x * MW * 65;
#pragma unroll
      for (int blk = 0; blk < (MW + 127) / 128; ++blk) {
        const int a = blk * 128 + row;                // gradient column index
        const bool ok = (MW == 64) ? (row < 64) : (a < MW);
#pragma unroll
        for (int c0 = 0; c0 < 96; c0 += 32) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(lane_base + blk * 96 + c0, r);
          tmem_ld_wait();
          if (ok) {
            if (c0 < 64) {
#pragma unroll
              for (int j = 0; j < 32; ++j) out[(int64_t)a * 65 + c0 + j] = __uint_as_float(r[j]);
            } else {
              out[(int64_t)a * 65 + 64] = __uint_as_float(r[0]);
            }
          }
        }
      }
    } else {
      atomicCAS(status, 0, 503 | (blockIdx.x << 16));
    }
  } else if (warp == 4) {
    // ------------------------------------------------------------------ MMA issuer
    {
      const uint32_t idesc = idesc_bf16(128, NB, 1, 1);
      // M = 128 gradient columns per MMA = two 64-column atoms 16 KB apart.  Where the second atom does not exist
      // (MW == 64, or the v block of MW == 192) the MMA reads whatever shared memory follows; every accumulator row
      // depends only on its own A row, so those rows are garbage that the epilogue never reads.
      const uint32_t lbo_a01 = (uint32_t)kAtomBytes;
      uint32_t it = 0;
      bool ok = true;
      for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const uint32_t st = it % NS;
        if (!mbar_wait(&sm.full[st], (it / NS) & 1)) { atomicCAS(status, 0, 502 | (blockIdx.x << 16)); ok = false; break; }
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t db = smem_desc(smem_u32(sm.b[st][0]) + ks * 2048, kAtomBytes, 1024, LAYOUT_SW128);
          mma_ss_w(tmem, smem_desc(smem_u32(sm.a[st][0]) + ks * 2048, lbo_a01, 1024, LAYOUT_SW128), db, idesc,
                 (it | ks) != 0);
          if (MW == 192)
            mma_ss_w(tmem + 96, smem_desc(smem_u32(sm.a[st][2]) + ks * 2048, lbo_a01, 1024, LAYOUT_SW128), db, idesc,
                   (it | ks) != 0);
        }
        mma_commit_w(&sm.empty[st]);
      }
      if (ok) mma_commit_w(&sm.done);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 256);
}

// d_w[a, 0..63] and d_b[a] from the per-CTA partial blocks [num_parts][MW][65]
__global__ void wgrad_reduce_kernel(const float* __restrict__ partials, int num_parts, int MW,
                                    float* __restrict__ d_w, float* __restrict__ d_b) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= MW * 65) return;
  float s = 0.f;
  for (int p = 0; p < num_parts; ++p) s += partials[(int64_t)p * MW * 65 + idx];
  const int a = idx / 65, c = idx - a * 65;
  if (c < 64) d_w[a * 64 + c] = s; else d_b[a] = s;
}

template <int MW, bool A_BF16 = false>
int launch_wgrad(const float* A, const float* B, const float* gate, int tokens_per_node, int64_t rows,
                 float* d_w, float* d_b, float* partials, size_t partial_bytes, int* status, cudaStream_t stream) {
  const int64_t tiles = (rows + 127) / 128;
  const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
  if ((size_t)grid * MW * 65 * sizeof(float) > partial_bytes) return AMPCONV_ERR_WORKSPACE;
  const size_t smem = sizeof(WgSmem<MW>) + 1024;
  AMPCONV_CUDA_TRY(cudaFuncSetAttribute(wgrad_tc_kernel<MW, A_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  wgrad_tc_kernel<MW, A_BF16><<<grid, kThreads, smem, stream>>>(A, B, gate, tokens_per_node, rows, partials, status);
  AMPCONV_CHECK_LAUNCH();
  wgrad_reduce_kernel<<<(MW * 65 + 255) / 256, 256, 0, stream>>>(partials, grid, MW, d_w, d_b);
  AMPCONV_CHECK_LAUNCH();
  return AMPCONV_OK;
}

}  // namespace
}  // namespace ampconv

using namespace ampconv;

// ws: scratch of ampconv_param_grad_workspace_bytes(3d, d) bytes; workspace: the family's 256-byte workspace.
extern "C" int ampconv_out_proj_bwd_params_tc(const float* d_out, const float* agg, const float* has_in,
                                              float* d_w, float* d_b, int64_t N, int F, int d,
                                              void* ws, size_t ws_bytes, void* workspace, void* stream_) {
  AMPCONV_REQUIRE(N >= 0 && F > 0 && d_w && d_b);
  if (d != 64) return AMPCONV_ERR_UNSUPPORTED;
  cudaStream_t stream = as_stream(stream_);
  if (N == 0) {
    AMPCONV_CUDA_TRY(cudaMemsetAsync(d_w, 0, sizeof(float) * d * d, stream));
    AMPCONV_CUDA_TRY(cudaMemsetAsync(d_b, 0, sizeof(float) * d, stream));
    return AMPCONV_OK;
  }
  AMPCONV_REQUIRE(d_out && agg && has_in && ws && workspace);
  return launch_wgrad<64>(d_out, agg, has_in, F, N * F, d_w, d_b, reinterpret_cast<float*>(ws), ws_bytes,
                          reinterpret_cast<int*>(workspace) + 1, stream);
}

extern "C" int ampconv_qkv_proj_bwd_params_tc(const float* x, const float* d_qkv, float* d_w, float* d_b,
                                              int64_t rows, int d, void* ws, size_t ws_bytes, void* workspace,
                                              void* stream_) {
  AMPCONV_REQUIRE(rows >= 0 && d_w && d_b);
  if (d != 64) return AMPCONV_ERR_UNSUPPORTED;
  cudaStream_t stream = as_stream(stream_);
  if (rows == 0) {
    AMPCONV_CUDA_TRY(cudaMemsetAsync(d_w, 0, sizeof(float) * 3 * d * d, stream));
    AMPCONV_CUDA_TRY(cudaMemsetAsync(d_b, 0, sizeof(float) * 3 * d, stream));
    return AMPCONV_OK;
  }
  AMPCONV_REQUIRE(x && d_qkv && ws && workspace);
  return launch_wgrad<192>(d_qkv, x, nullptr, 1, rows, d_w, d_b, reinterpret_cast<float*>(ws), ws_bytes,
                           reinterpret_cast<int*>(workspace) + 1, stream);
}

// in_proj parameter gradients from bf16 gradient rows d_qkv_bf16 [rows, 192] (ampconv_attn_bwd_*_bf16_h).
extern "C" int ampconv_qkv_proj_bwd_params_tc_h(const float* x, const void* d_qkv_bf16, float* d_w, float* d_b,
                                                int64_t rows, int d, void* ws, size_t ws_bytes, void* workspace,
                                                void* stream_) {
  AMPCONV_REQUIRE(rows >= 0 && d_w && d_b);
  if (d != 64) return AMPCONV_ERR_UNSUPPORTED;
  cudaStream_t stream = as_stream(stream_);
  if (rows == 0) {
    AMPCONV_CUDA_TRY(cudaMemsetAsync(d_w, 0, sizeof(float) * 3 * d * d, stream));
    AMPCONV_CUDA_TRY(cudaMemsetAsync(d_b, 0, sizeof(float) * 3 * d, stream));
    return AMPCONV_OK;
  }
  AMPCONV_REQUIRE(x && d_qkv_bf16 && ws && workspace);
  return launch_wgrad<192, true>(reinterpret_cast<const float*>(d_qkv_bf16), x, nullptr, 1, rows, d_w, d_b,
                                 reinterpret_cast<float*>(ws), ws_bytes, reinterpret_cast<int*>(workspace) + 1, stream);
}
