// MUFU.EX2 issue rate on the B200, alone and next to FMA-pipe work: the attention kernels of this repo are bound by
// exponentials (H * F^2 per edge and pass), so the roofline they are held against needs the measured rate, not a guess.
//   variant 0: ex2.approx.ftz.f32            (one result per lane and instruction)
//   variant 1: ex2.approx.ftz.bf16x2         (two results per lane and instruction)
//   variant 2: ex2.approx.f16x2
//   variant 3: f32 ex2 + 3 independent FFMA2 per ex2 (is the FMA pipe free while MUFU is saturated?)
//   variant 4: ex2_poly2 (the FMA-pipe exponential of umma.cuh) alone
// For 1..8 warps per SM sub-partition: results per clock per SM.
#include <cstdio>
#include <cstdlib>
#include "../ampnet_b200/csrc/umma.cuh"
using namespace ampconv::umma;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t ex2_f16x2(uint32_t x) {
  uint32_t y;
  asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
__device__ __forceinline__ uint32_t ex2_bf16x2_v(uint32_t x) {
  uint32_t y;
  asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}

template <int VARIANT>
__global__ void __launch_bounds__(1024, 1) mufu_kernel(float* out, long long* cyc, int iters) {
  constexpr int CH = 8;   // independent chains per thread
  float v[CH];
  uint32_t u[CH];
  float2 w[3] = {make_float2(0.1f, 0.2f), make_float2(0.3f, 0.4f), make_float2(0.5f, 0.6f)};
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    v[c] = -0.001f * (threadIdx.x + c);
    u[c] = 0xbc00bc00u + c;   // small negative halves
  }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      if (VARIANT == 0 || VARIANT == 3) v[c] = ex2_approx_v(v[c]) - 1.0f;
      if (VARIANT == 1) u[c] = ex2_bf16x2_v(u[c]) ^ 0x80008000u;
      if (VARIANT == 2) u[c] = ex2_f16x2(u[c]) ^ 0x80008000u;
      if (VARIANT == 3) {
        w[0] = f2fma(w[0], make_float2(0.999f, 0.999f), make_float2(1e-3f, 1e-3f));
        w[1] = f2fma(w[1], make_float2(0.999f, 0.999f), make_float2(1e-3f, 1e-3f));
        w[2] = f2fma(w[2], make_float2(0.999f, 0.999f), make_float2(1e-3f, 1e-3f));
      }
      if (VARIANT == 4) {
        const float2 r = ex2_poly2(make_float2(v[c], v[c] - 0.5f));
        v[c] = r.x - r.y - 0.3f;
      }
    }
  }
  const long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < CH; ++c) acc += v[c] + __uint_as_float(u[c]) * 1e-30f;
  acc += w[0].x + w[1].y + w[2].x;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int VARIANT>
void run(const char* name, float* out, long long* cyc) {
  const int iters = 2000;
  for (int wps = 1; wps <= 8; wps *= 2) {
    const int threads = wps * 4 * 32;
    mufu_kernel<VARIANT><<<148, threads>>>(out, cyc, 10);
    mufu_kernel<VARIANT><<<148, threads>>>(out, cyc, iters);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    long long h;
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    const double per_lane = (VARIANT == 1 || VARIANT == 2 || VARIANT == 4) ? 2.0 : 1.0;
    const double results = (double)iters * 8 * threads * per_lane;
    printf("%-44s warps/sub-partition=%d: %8.2f results per clock per SM\n", name, wps, results / (double)h);
  }
}

int main() {
  float* out;
  long long* cyc;
  CK(cudaMalloc(&out, 148 * 1024 * sizeof(float)));
  CK(cudaMalloc(&cyc, 8));
  run<0>("ex2.approx.ftz.f32", out, cyc);
  run<1>("ex2.approx.ftz.bf16x2", out, cyc);
  run<2>("ex2.approx.f16x2", out, cyc);
  run<3>("ex2.f32 + 3 FFMA2 per ex2 (ex2 results)", out, cyc);
  run<4>("ex2_poly2 on the FMA pipe", out, cyc);
  return 0;
}
