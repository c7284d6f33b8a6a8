// Elementwise stage of the backward kernels in isolation (no barriers, no MMAs): 16 warps repeat
// tcgen05.ld (X, Y 2 x 16 columns) -> exp2 / multiply / pack -> tcgen05.st over resident TMEM contents.
// Prints cycles per (edge, head) item = two half-items, one per group of 8 warps (MUFU floor: 1024).
//   variant 0: MODE_DQ math (row statistic in a register)     variant 1: MODE_DKV math (column statistics from smem)
//   variant 2: variant 0 without tcgen05.st                   variant 3: variant 0 without tcgen05.ld/st (registers only)
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include "../ampnet_b200/csrc/umma.cuh"
using namespace ampconv;
using namespace ampconv::umma;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

template <int VARIANT>
__global__ void __launch_bounds__(640, 1) ew_kernel(float* out, long long* cyc, int iters) {
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float stat[2][512];
  __shared__ uint64_t gbar[2], mbar_done;
  extern __shared__ __align__(1024) uint8_t dsm_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 1024; i += blockDim.x) (&stat[0][0])[i] = 0.001f * (i % 89);
  uint8_t* dsm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dsm_raw) + 1023) & ~uintptr_t(1023));
  if (VARIANT >= 4) for (int i = tid; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(dsm)[i] = 0x3c003c00u;
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) { mbar_init(&gbar[0], 8); mbar_init(&gbar[1], 8); mbar_init(&mbar_done, 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  float acc = 0.f;
  if (warp < 16) {
    const uint32_t q4 = warp & 3, grp = warp >> 3, cb = (warp >> 2) & 1;
    const uint32_t lane_base = tmem + ((uint32_t)(q4 * 32) << 16);
    {
      uint32_t z[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) z[j] = __float_as_uint(-0.01f * ((lane + j) % 37));
      for (int c = 0; c < 32; ++c) tmem_st_32x32b_x16(lane_base + 16 * c, z);
      tmem_st_wait();
    }
    const float L = 0.5f + 0.01f * lane;
    float dl = 0.f;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t set = (2 * it + grp) % 3;
      const uint32_t xbase = lane_base + set * 128 + 32 * cb;
      uint32_t xs[2][16], ys[2][16];
      if (VARIANT != 3) {  // (variants 4..6 = variant 0 plus extras)
        tmem_ld_32x32b_x16(xbase, xs[0]);
        tmem_ld_32x32b_x16(xbase + 64, ys[0]);
        tmem_ld_32x32b_x16(xbase + 16, xs[1]);
        tmem_ld_32x32b_x16(xbase + 64 + 16, ys[1]);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          xs[0][j] = __float_as_uint(-0.01f * j - dl * 1e-9f); xs[1][j] = __float_as_uint(-0.02f * j);
          ys[0][j] = __float_as_uint(0.5f + j); ys[1][j] = __float_as_uint(0.25f + j);
          asm volatile("mov.b32 %0, %0;" : "+r"(xs[0][j])); asm volatile("mov.b32 %0, %0;" : "+r"(xs[1][j]));
          asm volatile("mov.b32 %0, %0;" : "+r"(ys[0][j])); asm volatile("mov.b32 %0, %0;" : "+r"(ys[1][j]));
        }
      }
      const float* Ls = stat[0] + (it & 3) * 64;
      const float* Ds = stat[1] + (it & 3) * 64;
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        const int col0 = 32 * cb + 16 * ch;
        uint32_t px[8], py[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c0 = col0 + 2 * j;
          float p0, p1, u0, u1;
          if (VARIANT != 1) {
            p0 = ex2_approx(__uint_as_float(xs[ch][2 * j]) - L);
            p1 = ex2_approx(__uint_as_float(xs[ch][2 * j + 1]) - L);
            u0 = p0 * __uint_as_float(ys[ch][2 * j]);
            u1 = p1 * __uint_as_float(ys[ch][2 * j + 1]);
          } else {
            const float2 l2 = *reinterpret_cast<const float2*>(Ls + c0);
            const float2 d2 = *reinterpret_cast<const float2*>(Ds + c0);
            p0 = ex2_approx(__uint_as_float(xs[ch][2 * j]) - l2.x);
            p1 = ex2_approx(__uint_as_float(xs[ch][2 * j + 1]) - l2.y);
            u0 = p0 * (__uint_as_float(ys[ch][2 * j]) - d2.x);
            u1 = p1 * (__uint_as_float(ys[ch][2 * j + 1]) - d2.y);
          }
          dl += u0 + u1;
          px[j] = pack_bf16x2(p0, p1);
          py[j] = pack_bf16x2(u0, u1);
        }
        if (VARIANT < 2 || VARIANT >= 4) {
          tmem_st_32x32b_x8(xbase + 16 * ch, px);
          tmem_st_32x32b_x8(xbase + 64 + 16 * ch, py);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) asm volatile("xor.b32 %0, %0, %1;" : "+r"(px[0]) : "r"(px[j] ^ py[j]));
          dl += __uint_as_float(px[0]) * 1e-30f;
        }
      }
      if (VARIANT < 2 || VARIANT >= 4) tmem_st_wait();
      if (VARIANT == 5 || VARIANT == 6) {
        // group-wide handshake per half-item, as the publish / next-scores pair of the kernels
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&gbar[grp]);
        mbar_wait(&gbar[grp], it & 1);
        tc_fence_after();
      }
    }
    const long long t1 = clock64();
    acc = dl;
    if (tid == 0) cyc[0] = t1 - t0;
  } else if ((VARIANT == 4 || VARIANT == 6) && warp >= 17) {
    // concurrent MMA streams at roughly the kernels' rate: warp 17 scores (2 SS, N = 64), warps 18 / 19 consumers (4 TS, N = 16)
    const uint32_t id_ss = idesc_bf16(128, 64, 0, 0), id_ts = idesc_bf16(128, 16, 0, 1);
    const uint64_t da = smem_desc(smem_u32(dsm), 16, 1024, LAYOUT_SW128);
    const uint64_t db = smem_desc(smem_u32(dsm + 32768), 16, 1024, LAYOUT_SW128);
    for (int it = 0; it < 2 * iters; ++it) {
      if (warp == 17) {
        mma_ss_w(tmem + 384, da, db, id_ss, 0);
        mma_ss_w(tmem + 448, da, db, id_ss, 0);
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) mma_ts_w(tmem + 384 + 16 * (warp - 18), tmem + 128 * (it % 3) + 16 * u, desc_advance(db, u * 2048), id_ts, 1);
      }
      mma_commit_w(&mbar_done);   // count 1 barrier: phases just keep flipping
      __nanosleep(200);
    }
  }
  out[blockIdx.x * blockDim.x + tid] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int VARIANT>
void run(const char* name, float* out, long long* cyc) {
  const int iters = 4000;
  const size_t smem = 65536 + 1024;
  CK(cudaFuncSetAttribute(ew_kernel<VARIANT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ew_kernel<VARIANT><<<148, 640, smem>>>(out, cyc, 10);
  ew_kernel<VARIANT><<<148, 640, smem>>>(out, cyc, iters);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  long long h;
  CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
  printf("%-46s: %7.1f cycles per item (two half-items in parallel, 32 columns per thread)\n", name, (double)h / iters);
}

int main() {
  float* out; long long* cyc;
  CK(cudaMalloc(&out, 148 * 640 * sizeof(float)));
  CK(cudaMalloc(&cyc, 8));
  run<0>("MODE_DQ math, tcgen05.ld + st", out, cyc);
  run<1>("MODE_DKV math (smem statistics), ld + st", out, cyc);
  run<2>("MODE_DQ math, tcgen05.ld only", out, cyc);
  run<3>("MODE_DQ math, registers only", out, cyc);
  run<4>("variant 0 + concurrent MMA streams", out, cyc);
  run<5>("variant 0 + group handshake per half-item", out, cyc);
  run<6>("variant 0 + MMA streams + handshake", out, cyc);
  return 0;
}
