// Node-level projections of the bf16 family on tcgen05: Y[rows, N] = X[rows, K] * B^T with tiny K, N
// (64 or 192) and tens of millions of rows, i.e. HBM-bound tall-skinny GEMMs.  They replace the
// aten::addmm calls of F.multi_head_attention_forward (reference era copy
// src/ampnet/conv/custom_multihead_attn_forward.py:4031-4084 in-projection, :4436-4437 out-projection)
// and their input gradients, executed once per node token instead of once per edge token.
//
// One persistent kernel, warp-specialised:
//   warps 0-7  loaders : coalesced fp32 loads of a 128-row tile, convert to bf16, store into the 128B-swizzled
//                        K-major layout tcgen05 expects (3-stage ring);
//   warp  8    MMA     : one thread issues M=128, N, K/16 tcgen05.mma per tile into one of two TMEM accumulators;
//   warps 9-12 epilogue: tcgen05.ld the accumulator rows, fused bias / gate / row scale / q scaling, bf16 or fp32 stores.
// The weight matrix is converted once per CTA and stays in shared memory.
#include <cuda_bf16.h>

#include "common.cuh"
#include "umma.cuh"

namespace ampconv {
namespace {

using namespace umma;

constexpr int kLoadWarps = 8;                      // loader warps: enough loads in flight per SM to cover the HBM latency
constexpr int kLoadThreads = kLoadWarps * 32;
constexpr int kMmaWarp = kLoadWarps;               // then the MMA warp, then four epilogue warps (one per TMEM lane quarter)
constexpr int kThreads = (kLoadWarps + 5) * 32;
constexpr int kStages = 3;

enum : int { EPI_QKV = 0, EPI_OUT = 1, EPI_DAGG = 2, EPI_DX = 3 };

struct LinearArgs {
  const float* x;        // [rows, K] fp32
  const float* w;        // EPI_QKV / EPI_OUT: [N, K];  EPI_DAGG / EPI_DX: [K, N] (used transposed)
  const float* bias;     // [N] or null
  const float* node_vec; // EPI_OUT: has_in[node] gates the bias; EPI_DAGG: inv_deg[node] scales the row
  void* out0;            // EPI_QKV: q (bf16); EPI_OUT / EPI_DX: fp32 [rows, 64]; EPI_DAGG: bf16 [rows, 64]
  void* out1;            // EPI_QKV: k
  void* out2;            // EPI_QKV: v
  int64_t rows;
  int tokens_per_node;   // F
  float q_scale;
};

template <int K, int N>
struct LinSmem {
  uint8_t a[kStages][K / 64][128 * 128];   // [stage][k atom][128 rows x 64 bf16, swizzled]
  uint8_t b[K / 64][N * 128];              // [k atom][N rows x 64 bf16, swizzled]
  float bias[N];
  uint64_t a_full[kStages], a_empty[kStages], d_full[2], d_empty[2];
  uint32_t tmem_base;
};

// IN_BF16: the input rows are bf16 already (the gradient rows of the attention backward): 16-byte chunks are copied as they are.
template <int K, int N, int EPI, bool IN_BF16 = false>
__global__ void __launch_bounds__(kThreads, 1) linear_tc_kernel(const LinearArgs args, int* __restrict__ status) {
  using Smem = LinSmem<K, N>;
  extern __shared__ uint8_t smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr bool kTransposedW = (EPI == EPI_DAGG || EPI == EPI_DX);
  constexpr int kTmemCols = 512;

  if (warp == kMmaWarp) tmem_alloc(&sm.tmem_base, kTmemCols);
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&sm.a_full[i], kLoadWarps);
      mbar_init(&sm.a_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sm.d_full[i], 1);
      mbar_init(&sm.d_empty[i], 4);
    }
    fence_barrier_init();
  }
  // weights -> bf16, swizzled K-major [N rows][K cols]
  for (int idx = threadIdx.x; idx < N * K; idx += kThreads) {
    const int n = idx / K, k = idx - n * K;
    const float wv = kTransposedW ? args.w[(int64_t)k * N + n] : args.w[(int64_t)n * K + k];
    *reinterpret_cast<__nv_bfloat16*>(sm.b[k >> 6] + sw128_offset(n, (k & 63) * 2)) = __float2bfloat16(wv);
  }
  for (int idx = threadIdx.x; idx < N; idx += kThreads) sm.bias[idx] = args.bias ? args.bias[idx] : 0.f;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;
  const int64_t num_tiles = (args.rows + 127) / 128;

  if (warp < kLoadWarps) {
    // ------------------------------------------------------------------ loaders
    constexpr int kChunksPerRow = K / 8;                 // 16-byte bf16 chunks (8 elements) per row
    constexpr int kChunksPerThread = 128 * kChunksPerRow / kLoadThreads;
    const int tid = threadIdx.x;
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const uint32_t st = it % kStages;
      if (!mbar_wait(&sm.a_empty[st], ((it / kStages) & 1) ^ 1)) { atomicCAS(status, 0, 401 | (blockIdx.x << 16)); break; }
      const int64_t row0 = tile * 128;
      if constexpr (IN_BF16) {
        const uint4* xb = reinterpret_cast<const uint4*>(args.x);     // bf16 [rows, K]: K / 8 chunks of 16 bytes per row
        uint4 v[kChunksPerThread > 8 ? 8 : kChunksPerThread];
#pragma unroll
        for (int base = 0; base < kChunksPerThread; base += 8) {
#pragma unroll
          for (int u = 0; u < 8 && base + u < kChunksPerThread; ++u) {
            const int q = (base + u) * kLoadThreads + tid;
            const int r = q / kChunksPerRow, c = q - r * kChunksPerRow;
            const int64_t grow = row0 + r;
            v[u] = grow < args.rows ? __ldg(xb + grow * kChunksPerRow + c) : make_uint4(0u, 0u, 0u, 0u);
          }
#pragma unroll
          for (int u = 0; u < 8 && base + u < kChunksPerThread; ++u) {
            const int q = (base + u) * kLoadThreads + tid;
            const int r = q / kChunksPerRow, c = q - r * kChunksPerRow;
            *reinterpret_cast<uint4*>(sm.a[st][c >> 3] + sw128_offset(r, (c & 7) * 16)) = v[u];
          }
        }
      } else {
      float4 v[kChunksPerThread > 8 ? 8 : kChunksPerThread][2];
#pragma unroll
      for (int base = 0; base < kChunksPerThread; base += 8) {
#pragma unroll
        for (int u = 0; u < 8 && base + u < kChunksPerThread; ++u) {
          const int q = (base + u) * kLoadThreads + tid;
          const int r = q / kChunksPerRow, c = q - r * kChunksPerRow;
          const int64_t grow = row0 + r;
          if (grow < args.rows) {
            const float4* src = reinterpret_cast<const float4*>(args.x + grow * K + c * 8);
            v[u][0] = __ldg(src);
            v[u][1] = __ldg(src + 1);
          } else {
            v[u][0] = make_float4(0.f, 0.f, 0.f, 0.f);
            v[u][1] = v[u][0];
          }
        }
#pragma unroll
        for (int u = 0; u < 8 && base + u < kChunksPerThread; ++u) {
          const int q = (base + u) * kLoadThreads + tid;
          const int r = q / kChunksPerRow, c = q - r * kChunksPerRow;
          uint4 pk;
          pk.x = pack_bf16x2(v[u][0].x, v[u][0].y);
          pk.y = pack_bf16x2(v[u][0].z, v[u][0].w);
          pk.z = pack_bf16x2(v[u][1].x, v[u][1].y);
          pk.w = pack_bf16x2(v[u][1].z, v[u][1].w);
          *reinterpret_cast<uint4*>(sm.a[st][c >> 3] + sw128_offset(r, (c & 7) * 16)) = pk;
        }
      }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.a_full[st]);
    }
  } else if (warp == kMmaWarp) {
    // ------------------------------------------------------------------ MMA issuer
    {
      const uint32_t idesc = idesc_bf16(128, N, 0, 0);
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const uint32_t st = it % kStages, buf = it & 1;
        if (!mbar_wait(&sm.a_full[st], (it / kStages) & 1)) { atomicCAS(status, 0, 402 | (blockIdx.x << 16)); break; }
        if (!mbar_wait(&sm.d_empty[buf], ((it >> 1) & 1) ^ 1)) { atomicCAS(status, 0, 403 | (blockIdx.x << 16)); break; }
        tc_fence_after();
#pragma unroll
        for (int ka = 0; ka < K / 64; ++ka)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            mma_ss_w(tmem + buf * 256, smem_desc(smem_u32(sm.a[st][ka]) + ks * 32, 16, 1024, LAYOUT_SW128),
                   smem_desc(smem_u32(sm.b[ka]) + ks * 32, 16, 1024, LAYOUT_SW128), idesc, (ka | ks) != 0);
        mma_commit_w(&sm.a_empty[st]);
        mma_commit_w(&sm.d_full[buf]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue
    const int quarter = warp & 3;
    const int row_in_tile = quarter * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(quarter * 32) << 16);
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it & 1;
      if (!mbar_wait(&sm.d_full[buf], (it >> 1) & 1)) { atomicCAS(status, 0, 404 | (blockIdx.x << 16)); break; }
      tc_fence_after();
      const int64_t grow = tile * 128 + row_in_tile;
      const bool ok = grow < args.rows;
      float node_v = 1.f;
      if ((EPI == EPI_OUT || EPI == EPI_DAGG) && ok) node_v = args.node_vec[grow / args.tokens_per_node];
#pragma unroll
      for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(lane_base + buf * 256 + c0, r);
        tmem_ld_wait();
        if (ok) {
          if (EPI == EPI_QKV) {
            const int blk = c0 / 64;
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(blk == 0 ? args.out0 : (blk == 1 ? args.out1 : args.out2)) +
                                 grow * 64 + (c0 & 63);
            const float sc = blk == 0 ? args.q_scale : 1.f;
            uint4 pk[4];
            uint32_t* pw = reinterpret_cast<uint32_t*>(pk);
#pragma unroll
            for (int j = 0; j < 16; ++j)
              pw[j] = pack_bf16x2((__uint_as_float(r[2 * j]) + sm.bias[c0 + 2 * j]) * sc,
                                  (__uint_as_float(r[2 * j + 1]) + sm.bias[c0 + 2 * j + 1]) * sc);
#pragma unroll
            for (int j = 0; j < 4; ++j) reinterpret_cast<uint4*>(dst)[j] = pk[j];
          } else if (EPI == EPI_DAGG) {
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(args.out0) + grow * 64 + c0;
            uint4 pk[4];
            uint32_t* pw = reinterpret_cast<uint32_t*>(pk);
#pragma unroll
            for (int j = 0; j < 16; ++j)
              pw[j] = pack_bf16x2(__uint_as_float(r[2 * j]) * node_v, __uint_as_float(r[2 * j + 1]) * node_v);
#pragma unroll
            for (int j = 0; j < 4; ++j) reinterpret_cast<uint4*>(dst)[j] = pk[j];
          } else {
            float* dst = reinterpret_cast<float*>(args.out0) + grow * 64 + c0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 o;
              if (EPI == EPI_OUT) {
                o.x = __uint_as_float(r[4 * j]) + sm.bias[c0 + 4 * j] * node_v;
                o.y = __uint_as_float(r[4 * j + 1]) + sm.bias[c0 + 4 * j + 1] * node_v;
                o.z = __uint_as_float(r[4 * j + 2]) + sm.bias[c0 + 4 * j + 2] * node_v;
                o.w = __uint_as_float(r[4 * j + 3]) + sm.bias[c0 + 4 * j + 3] * node_v;
              } else {
                o = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                                __uint_as_float(r[4 * j + 3]));
              }
              reinterpret_cast<float4*>(dst)[j] = o;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.d_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem, kTmemCols);
}

template <int K, int N, int EPI, bool IN_BF16 = false>
int launch_linear(const LinearArgs& args, int* status, cudaStream_t stream) {
  const size_t smem = sizeof(LinSmem<K, N>) + 1024;
  AMPCONV_CUDA_TRY(cudaFuncSetAttribute(linear_tc_kernel<K, N, EPI, IN_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t tiles = (args.rows + 127) / 128;
  const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
  linear_tc_kernel<K, N, EPI, IN_BF16><<<grid, kThreads, smem, stream>>>(args, status);
  AMPCONV_CHECK_LAUNCH();
  return AMPCONV_OK;
}

}  // namespace
}  // namespace ampconv

using namespace ampconv;

// All four entry points require d == 64 (the tcgen05 family); `workspace` is the family's 256-byte workspace
// (status word at int index 1).

extern "C" int ampconv_qkv_proj_tc(const float* x, const float* w, const float* b, void* q, void* k, void* v,
                                   int64_t rows, int d, float q_scale, void* workspace, void* stream) {
  AMPCONV_REQUIRE(rows >= 0 && d > 0);
  if (d != 64) return AMPCONV_ERR_UNSUPPORTED;
  if (rows == 0) return AMPCONV_OK;
  AMPCONV_REQUIRE(x && w && b && q && k && v && workspace);
  LinearArgs a{};
  a.x = x; a.w = w; a.bias = b; a.out0 = q; a.out1 = k; a.out2 = v; a.rows = rows; a.tokens_per_node = 1; a.q_scale = q_scale;
  return launch_linear<64, 192, EPI_QKV>(a, reinterpret_cast<int*>(workspace) + 1, as_stream(stream));
}

extern "C" int ampconv_out_proj_tc(const float* agg, const float* w, const float* b, const float* has_in, float* out,
                                   int64_t N, int F, int d, void* workspace, void* stream) {
  AMPCONV_REQUIRE(N >= 0 && F > 0 && d > 0);
  if (d != 64) return AMPCONV_ERR_UNSUPPORTED;
  if (N == 0) return AMPCONV_OK;
  AMPCONV_REQUIRE(agg && w && b && has_in && out && workspace);
  LinearArgs a{};
  a.x = agg; a.w = w; a.bias = b; a.node_vec = has_in; a.out0 = out; a.rows = N * F; a.tokens_per_node = F; a.q_scale = 1.f;
  return launch_linear<64, 64, EPI_OUT>(a, reinterpret_cast<int*>(workspace) + 1, as_stream(stream));
}

extern "C" int ampconv_out_proj_bwd_input_tc(const float* d_out, const float* w, const float* inv_deg, void* d_agg_bf16,
                                             int64_t N, int F, int d, void* workspace, void* stream) {
  AMPCONV_REQUIRE(N >= 0 && F > 0 && d > 0);
  if (d != 64) return AMPCONV_ERR_UNSUPPORTED;
  if (N == 0) return AMPCONV_OK;
  AMPCONV_REQUIRE(d_out && w && inv_deg && d_agg_bf16 && workspace);
  LinearArgs a{};
  a.x = d_out; a.w = w; a.node_vec = inv_deg; a.out0 = d_agg_bf16; a.rows = N * F; a.tokens_per_node = F; a.q_scale = 1.f;
  return launch_linear<64, 64, EPI_DAGG>(a, reinterpret_cast<int*>(workspace) + 1, as_stream(stream));
}

extern "C" int ampconv_qkv_proj_bwd_input_tc(const float* d_qkv, const float* w, float* d_x, int64_t rows, int d,
                                             void* workspace, void* stream) {
  AMPCONV_REQUIRE(rows >= 0 && d > 0);
  if (d != 64) return AMPCONV_ERR_UNSUPPORTED;
  if (rows == 0) return AMPCONV_OK;
  AMPCONV_REQUIRE(d_qkv && w && d_x && workspace);
  LinearArgs a{};
  a.x = d_qkv; a.w = w; a.out0 = d_x; a.rows = rows; a.tokens_per_node = 1; a.q_scale = 1.f;
  return launch_linear<192, 64, EPI_DX>(a, reinterpret_cast<int*>(workspace) + 1, as_stream(stream));
}

// d_x from bf16 gradient rows d_qkv_bf16 [rows, 192] (ampconv_attn_bwd_*_bf16_h).
extern "C" int ampconv_qkv_proj_bwd_input_tc_h(const void* d_qkv_bf16, const float* w, float* d_x, int64_t rows, int d,
                                               void* workspace, void* stream) {
  AMPCONV_REQUIRE(rows >= 0 && d > 0);
  if (d != 64) return AMPCONV_ERR_UNSUPPORTED;
  if (rows == 0) return AMPCONV_OK;
  AMPCONV_REQUIRE(d_qkv_bf16 && w && d_x && workspace);
  LinearArgs a{};
  a.x = reinterpret_cast<const float*>(d_qkv_bf16); a.w = w; a.out0 = d_x; a.rows = rows; a.tokens_per_node = 1; a.q_scale = 1.f;
  return launch_linear<192, 64, EPI_DX, true>(a, reinterpret_cast<int*>(workspace) + 1, as_stream(stream));
}
