// Hardware probe for the tcgen05 building blocks the AMPConv kernels rely on (run on the B200 box):
//   mode 0  S = Q_h K_h^T        SS MMA, K-major SW128 operands written by threads (manual swizzle)
//   mode 1  O = P V_h            TS MMA (A = bf16 P in TMEM), B = V tile as MN-major SW128, 8 K-steps
//   mode 2  O = P V_h            SS MMA, A = P in smem (K-major SW128, two 64-column atoms)
//   mode 3  TMA 3D tensor-map load (box 128 x 64 over an F=100 node): raw smem image vs expected swizzle
//   mode 4  mode 0 with both tiles brought in by TMA
//   mode 6  D[128x64] = A^T B    SS MMA, BOTH operands MN-major SW128 (A = two 64-column atoms, LBO = atom stride)
//   mode 7  MMA issue/completion cost: 256 back-to-back TS (h=0) or SS (h=1) MMAs, N = lbo arg, #accumulators = sbo arg
//   mode 5  MUFU ex2 throughput: f32 vs bf16x2 vs f16x2; fma.rn.f32x2 vs scalar fma
// Every mbarrier wait is bounded, so a wrong descriptor reports a timeout instead of hanging the GPU.
// Usage: umma_probe <mode> [h] [lbo_bytes] [sbo_bytes] [b_major_mn]
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_bf16.h>

#include "../ampnet_b200/csrc/umma.cuh"

using namespace ampconv;
using namespace ampconv::umma;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

struct Params {
  int mode, h, lbo, sbo, b_mn;
};

__global__ void __launch_bounds__(128)
probe_kernel(Params p, const __nv_bfloat16* __restrict__ Q, const __nv_bfloat16* __restrict__ K,
             const __nv_bfloat16* __restrict__ V, const __nv_bfloat16* __restrict__ P,
             const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
             float* __restrict__ out, int* __restrict__ status) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;              // 32 KB region (Q tile, or P as two 16 KB atoms)
  uint8_t* sB = smem + 32768;      // 16 KB region (K or V tile)
  __shared__ uint64_t bar_mma, bar_tma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;

  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    mbar_init(&bar_mma, 1);
    mbar_init(&bar_tma, 1);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
  bool ok = true;

  if (p.mode == 0 || p.mode == 1 || p.mode == 2 || p.mode == 6) {
    // manual swizzled fill: tile rows are 128 B (64 bf16)
    const __nv_bfloat16* srcB = (p.mode == 0) ? K : V;
    for (int idx = tid; idx < 128 * 8; idx += 128) {       // 16-byte chunks
      int row = idx >> 3, chunk = idx & 7;
      uint4 vb = *reinterpret_cast<const uint4*>(srcB + row * 64 + chunk * 8);
      *reinterpret_cast<uint4*>(sB + sw128_offset(row, chunk * 16)) = vb;
      if (p.mode == 0) {
        uint4 va = *reinterpret_cast<const uint4*>(Q + row * 64 + chunk * 8);
        *reinterpret_cast<uint4*>(sA + sw128_offset(row, chunk * 16)) = va;
      }
    }
    if (p.mode == 2 || p.mode == 6) {
      for (int idx = tid; idx < 128 * 16; idx += 128) {    // P is 128 x 128: two atoms of 64 columns
        int row = idx >> 4, chunk = idx & 15;
        uint4 va = *reinterpret_cast<const uint4*>(P + row * 128 + chunk * 8);
        *reinterpret_cast<uint4*>(sA + (chunk >> 3) * 16384 + sw128_offset(row, (chunk & 7) * 16)) = va;
      }
    }
    fence_proxy_async_smem();
  }
  if (p.mode == 1) {
    // P row of this thread -> TMEM columns [128, 192): two bf16 per 32-bit column, even k in the low half
    const int row = tid;
    for (int c0 = 0; c0 < 64; c0 += 16) {
      uint32_t r[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(P + row * 128 + 2 * (c0 + c));
        r[c] = *reinterpret_cast<uint32_t*>(&v);
      }
      tmem_st_32x32b_x16(lane_addr + 128 + c0, r);
    }
    tmem_st_wait();
  }
  if (p.mode == 3 || p.mode == 4) {
    if (tid == 0) {
      mbar_arrive_expect_tx(&bar_tma, p.mode == 4 ? 32768 : 16384);
      tma_load_3d(sA, &mapQ, &bar_tma, 0, 0, 1);
      if (p.mode == 4) tma_load_3d(sB, &mapK, &bar_tma, 0, 0, 1);
    }
    ok = mbar_wait(&bar_tma, 0);
    if (!ok && tid == 0) atomicExch(status, 1);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (p.mode == 3) {
    for (int i = tid; i < 4096; i += 128) out[i] = reinterpret_cast<float*>(sA)[i];
  } else if (p.mode != 5) {
    if (tid == 0) {
      if (p.mode == 0 || p.mode == 4) {
        uint64_t da = smem_desc(smem_u32(sA) + p.h * 32, p.lbo, p.sbo, LAYOUT_SW128);
        uint64_t db = smem_desc(smem_u32(sB) + p.h * 32, p.lbo, p.sbo, LAYOUT_SW128);
        mma_ss(tmem, da, db, idesc_bf16(128, 128, 0, 0), 0);
      } else if (p.mode == 1) {
        for (int s = 0; s < 8; ++s) {
          uint64_t db = smem_desc(smem_u32(sB) + s * 2048 + p.h * 32, p.lbo, p.sbo, LAYOUT_SW128);
          mma_ts(tmem + 256, tmem + 128 + 8 * s, db, idesc_bf16(128, 16, 0, p.b_mn), s > 0);
        }
      } else if (p.mode == 6) {
        for (int s = 0; s < 8; ++s) {
          uint64_t da = smem_desc(smem_u32(sA) + s * 2048, p.lbo, p.sbo, LAYOUT_SW128);
          uint64_t db = smem_desc(smem_u32(sB) + s * 2048, 16, 1024, LAYOUT_SW128);
          mma_ss(tmem + 256, da, db, idesc_bf16(128, 64, 1, 1), s > 0);
        }
      } else {  // mode 2
        for (int s = 0; s < 8; ++s) {
          uint64_t da = smem_desc(smem_u32(sA) + (s >> 2) * 16384 + (s & 3) * 32, 16, 1024, LAYOUT_SW128);
          uint64_t db = smem_desc(smem_u32(sB) + s * 2048 + p.h * 32, p.lbo, p.sbo, LAYOUT_SW128);
          mma_ss(tmem + 256, da, db, idesc_bf16(128, 16, 0, p.b_mn), s > 0);
        }
      }
      mma_commit(&bar_mma);
    }
    ok = mbar_wait(&bar_mma, 0);
    if (!ok && tid == 0) atomicExch(status, 2);
    tc_fence_after();
    if (ok) {
      const int row = tid;
      if (p.mode == 0 || p.mode == 4) {
        for (int c0 = 0; c0 < 128; c0 += 32) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(lane_addr + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 32; ++c) out[row * 128 + c0 + c] = __uint_as_float(r[c]);
        }
      } else if (p.mode == 6) {
        for (int c0 = 0; c0 < 64; c0 += 32) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(lane_addr + 256 + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 32; ++c) out[row * 64 + c0 + c] = __uint_as_float(r[c]);
        }
      } else {
        uint32_t r[16];
        tmem_ld_32x32b_x16(lane_addr + 256, r);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 16; ++c) out[row * 16 + c] = __uint_as_float(r[c]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------ mode 5: pipe throughput
template <int KIND>
__global__ void __launch_bounds__(256) pipe_kernel(float* out, int iters) {
  float a0 = -0.001f * threadIdx.x, a1 = a0 - 0.5f, a2 = a0 - 1.0f, a3 = a0 - 1.5f;
  uint32_t b0 = pack_bf16x2(a0, a1), b1 = pack_bf16x2(a2, a3), b2 = pack_bf16x2(a1, a2), b3 = pack_bf16x2(a3, a0);
  for (int i = 0; i < iters; ++i) {
    if (KIND == 0) {          // f32 ex2
      a0 = ex2_approx(a0) - 1.0f; a1 = ex2_approx(a1) - 1.0f; a2 = ex2_approx(a2) - 1.0f; a3 = ex2_approx(a3) - 1.0f;
    } else if (KIND == 1) {   // bf16x2 ex2
      b0 = ex2_bf16x2(b0) ^ 0x80008000u; b1 = ex2_bf16x2(b1) ^ 0x80008000u;
      b2 = ex2_bf16x2(b2) ^ 0x80008000u; b3 = ex2_bf16x2(b3) ^ 0x80008000u;
    } else if (KIND == 2) {   // f16x2 ex2
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(b0)); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(b1));
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(b2)); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(b3));
      b0 ^= 0x80008000u; b1 ^= 0x80008000u; b2 ^= 0x80008000u; b3 ^= 0x80008000u;
    } else if (KIND == 3) {   // scalar fma
      a0 = fmaf(a0, 1.0001f, 0.5f); a1 = fmaf(a1, 1.0001f, 0.5f); a2 = fmaf(a2, 1.0001f, 0.5f); a3 = fmaf(a3, 1.0001f, 0.5f);
    } else {                  // packed fma.rn.f32x2
      unsigned long long x = ((unsigned long long)__float_as_uint(a0) << 32) | __float_as_uint(a1);
      unsigned long long y = ((unsigned long long)__float_as_uint(a2) << 32) | __float_as_uint(a3);
      const unsigned long long m = ((unsigned long long)__float_as_uint(1.0001f) << 32) | __float_as_uint(1.0001f);
      const unsigned long long c = ((unsigned long long)__float_as_uint(0.5f) << 32) | __float_as_uint(0.5f);
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x) : "l"(m), "l"(c));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(y) : "l"(m), "l"(c));
      a0 = __uint_as_float((uint32_t)(x >> 32)); a1 = __uint_as_float((uint32_t)x);
      a2 = __uint_as_float((uint32_t)(y >> 32)); a3 = __uint_as_float((uint32_t)y);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + __uint_as_float(b0 ^ b1 ^ b2 ^ b3);
}

template <int KIND>
static int run_pipe(const char* name, int elems_per_iter) {
  float* out;
  CK(cudaMalloc(&out, 148 * 8 * 256 * sizeof(float)));
  const int iters = 20000;
  pipe_kernel<KIND><<<148 * 8, 256>>>(out, 100);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  pipe_kernel<KIND><<<148 * 8, 256>>>(out, iters);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  double elems = (double)148 * 8 * 256 * iters * elems_per_iter;
  printf("pipe %-12s: %.3f ms  %.1f Gelem/s  (%.1f elem/clk/SM at 1.965 GHz)\n", name, ms, elems / ms * 1e-6,
         elems / (ms * 1e-3) / 148 / 1.965e9);
  cudaFree(out);
  return 0;
}

__global__ void __launch_bounds__(128)
mma_cost_kernel(int use_ss, int N, int nacc, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 49152 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp == 1) {
    const uint32_t idesc = idesc_bf16(128, N, 0, use_ss ? 0 : 1);
    const long long t0 = clock64();
    for (int i = 0; i < 256; ++i) {
      const uint32_t d = tmem + 256 + (i % nacc) * (N < 32 ? 32 : N);
      if (use_ss)
        mma_ss_w(d, smem_desc(smem_u32(smem) + (i & 3) * 32, 16, 1024, LAYOUT_SW128),
                 smem_desc(smem_u32(smem + 16384) + (i & 3) * 32, 16, 1024, LAYOUT_SW128), idesc, 1);
      else
        mma_ts_w(d, tmem + 8 * (i & 7), smem_desc(smem_u32(smem + 16384) + (i & 7) * 2048, 16, 1024, LAYOUT_SW128), idesc, 1);
    }
    const long long t1 = clock64();
    mma_commit_w(&bar);
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    if (tid == 32) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main(int argc, char** argv) {
  Params p{0, 0, 16, 1024, 1};
  if (argc > 1) p.mode = atoi(argv[1]);
  if (argc > 2) p.h = atoi(argv[2]);
  if (argc > 3) p.lbo = atoi(argv[3]);
  if (argc > 4) p.sbo = atoi(argv[4]);
  if (argc > 5) p.b_mn = atoi(argv[5]);
  printf("mode=%d h=%d lbo=%d sbo=%d b_mn=%d\n", p.mode, p.h, p.lbo, p.sbo, p.b_mn);
  if (p.mode == 7) {
    long long* dout;
    CK(cudaMalloc(&dout, 16));
    const size_t smem = 49152 + 1024;
    CK(cudaFuncSetAttribute(mma_cost_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int rep = 0; rep < 2; ++rep) mma_cost_kernel<<<1, 128, smem>>>(p.h, p.lbo, p.sbo, dout);
    CK(cudaDeviceSynchronize());
    long long h[2];
    CK(cudaMemcpy(h, dout, 16, cudaMemcpyDeviceToHost));
    printf("RESULT mode=7 %s N=%d accumulators=%d: issue %.1f clk/MMA, issue+complete %.1f clk/MMA\n", p.h ? "SS" : "TS",
           p.lbo, p.sbo, h[0] / 256.0, h[1] / 256.0);
    return 0;
  }
  if (p.mode == 5) {
    run_pipe<0>("ex2.f32", 4);
    run_pipe<1>("ex2.bf16x2", 8);
    run_pipe<2>("ex2.f16x2", 8);
    run_pipe<3>("fma.f32", 4);
    run_pipe<4>("fma.f32x2", 4);
    return 0;
  }
  // node-major tensors [3 nodes][F=100][64]; the probe uses node 1
  const int F = 100, NODES = 3;
  std::vector<float> hq(NODES * F * 64), hk(NODES * F * 64), hv(128 * 64), hp(128 * 128);
  srand(1);
  auto rnd = []() { return (float)(rand() % 2001 - 1000) / 1000.0f; };
  for (auto& x : hq) x = bf(rnd());
  for (auto& x : hk) x = bf(rnd());
  for (auto& x : hv) x = bf(rnd());
  for (auto& x : hp) x = bf(fabsf(rnd()));
  std::vector<__nv_bfloat16> bq(hq.size()), bk(hk.size()), bv(hv.size()), bp(hp.size());
  for (size_t i = 0; i < hq.size(); ++i) bq[i] = __float2bfloat16(hq[i]);
  for (size_t i = 0; i < hk.size(); ++i) bk[i] = __float2bfloat16(hk[i]);
  for (size_t i = 0; i < hv.size(); ++i) bv[i] = __float2bfloat16(hv[i]);
  for (size_t i = 0; i < hp.size(); ++i) bp[i] = __float2bfloat16(hp[i]);
  __nv_bfloat16 *dq, *dk, *dv, *dp;
  float* dout;
  int* dstatus;
  CK(cudaMalloc(&dq, bq.size() * 2)); CK(cudaMalloc(&dk, bk.size() * 2));
  CK(cudaMalloc(&dv, bv.size() * 2)); CK(cudaMalloc(&dp, bp.size() * 2));
  CK(cudaMalloc(&dout, 128 * 128 * 4)); CK(cudaMalloc(&dstatus, 4));
  CK(cudaMemcpy(dq, bq.data(), bq.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dk, bk.data(), bk.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dv, bv.data(), bv.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dp, bp.data(), bp.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dout, 0, 128 * 128 * 4)); CK(cudaMemset(dstatus, 0, 4));
  CUtensorMap mq, mk;
  if (!make_tensor_map_bf16_3d(&mq, dq, 64, F, NODES, 64, 128) || !make_tensor_map_bf16_3d(&mk, dk, 64, F, NODES, 64, 128)) {
    printf("tensor map creation failed\n");
    return 3;
  }
  // for modes 0-2 the kernel reads Q/K rows of "node 1" laid out as a dense 128 x 64 tile: build those
  std::vector<float> tq(128 * 64, 0.f), tk(128 * 64, 0.f);
  for (int r = 0; r < 128; ++r)
    for (int c = 0; c < 64; ++c) {
      tq[r * 64 + c] = r < F ? hq[(1 * F + r) * 64 + c] : 0.f;
      tk[r * 64 + c] = r < F ? hk[(1 * F + r) * 64 + c] : 0.f;
    }
  std::vector<__nv_bfloat16> btq(128 * 64), btk(128 * 64);
  for (int i = 0; i < 128 * 64; ++i) { btq[i] = __float2bfloat16(tq[i]); btk[i] = __float2bfloat16(tk[i]); }
  __nv_bfloat16 *dtq, *dtk;
  CK(cudaMalloc(&dtq, 128 * 64 * 2)); CK(cudaMalloc(&dtk, 128 * 64 * 2));
  CK(cudaMemcpy(dtq, btq.data(), 128 * 64 * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dtk, btk.data(), 128 * 64 * 2, cudaMemcpyHostToDevice));

  const size_t smem = 49152 + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_kernel<<<1, 128, smem>>>(p, dtq, dtk, dv, dp, mq, mk, dout, dstatus);
  CK(cudaDeviceSynchronize());
  int hstatus = 0;
  std::vector<float> ho(128 * 128);
  CK(cudaMemcpy(&hstatus, dstatus, 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(ho.data(), dout, 128 * 128 * 4, cudaMemcpyDeviceToHost));
  if (hstatus) { printf("RESULT mode=%d TIMEOUT status=%d\n", p.mode, hstatus); return 1; }

  double max_err = 0, max_ref = 0;
  if (p.mode == 0 || p.mode == 4) {
    for (int i = 0; i < 128; ++i)
      for (int j = 0; j < 128; ++j) {
        double s = 0;
        for (int c = 0; c < 16; ++c) s += (double)tq[i * 64 + 16 * p.h + c] * tk[j * 64 + 16 * p.h + c];
        max_err = fmax(max_err, fabs(s - ho[i * 128 + j]));
        max_ref = fmax(max_ref, fabs(s));
      }
  } else if (p.mode == 1 || p.mode == 2) {
    for (int i = 0; i < 128; ++i)
      for (int c = 0; c < 16; ++c) {
        double s = 0;
        for (int j = 0; j < 128; ++j) s += (double)hp[i * 128 + j] * hv[j * 64 + 16 * p.h + c];
        max_err = fmax(max_err, fabs(s - ho[i * 16 + c]));
        max_ref = fmax(max_ref, fabs(s));
      }
  } else if (p.mode == 6) {
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 64; ++n) {
        double s = 0;
        for (int k = 0; k < 128; ++k) s += (double)hp[k * 128 + m] * hv[k * 64 + n];
        max_err = fmax(max_err, fabs(s - ho[m * 64 + n]));
        max_ref = fmax(max_ref, fabs(s));
      }
  } else if (p.mode == 3) {
    const uint16_t* img = reinterpret_cast<const uint16_t*>(ho.data());
    int bad = 0;
    for (int r = 0; r < 128; ++r)
      for (int c = 0; c < 64; ++c) {
        uint16_t got = img[sw128_offset(r, c * 2) / 2];
        __nv_bfloat16 e = __float2bfloat16(tq[r * 64 + c]);
        uint16_t exp_bits = *reinterpret_cast<uint16_t*>(&e);
        if (got != exp_bits) ++bad;
      }
    printf("RESULT mode=3 TMA swizzle image mismatches=%d of 8192 (rows >= %d must be zero-filled)\n", bad, F);
    return bad ? 1 : 0;
  }
  printf("RESULT mode=%d h=%d lbo=%d sbo=%d b_mn=%d max_abs_err=%.6f max_ref=%.3f %s\n", p.mode, p.h, p.lbo, p.sbo,
         p.b_mn, max_err, max_ref, max_err < 1e-2 * fmax(1.0, max_ref) ? "PASS" : "FAIL");
  return max_err < 1e-2 * fmax(1.0, max_ref) ? 0 : 1;
}
