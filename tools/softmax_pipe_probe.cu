// How many softmax warps per SM sub-partition does it take to saturate the MUFU pipe on B200?
// Each warp repeats the register-level work of one forward item (128 scores per thread): row max (FMNMX3),
// subtract, ex2, row sum, bf16 pack, bf16x2 normalise.  No memory traffic; scores are made opaque to the
// compiler each iteration with empty asm barriers.  Prints SM-sub-partition cycles per item for 1..5 warps per
// sub-partition (the MUFU floor is 128 * 8 = 1024 cycles per item).
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>

#include "../ampnet_b200/csrc/umma.cuh"

using namespace ampconv::umma;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

template <int VARIANT, int WPS>
__global__ void __launch_bounds__(128 * WPS, 1) softmax_kernel(float* out, long long* cyc, int iters) {
  uint32_t s[128];
#pragma unroll
  for (int j = 0; j < 128; ++j) s[j] = __float_as_uint(-0.01f * (float)((threadIdx.x * 7 + j * 13) % 97));
  float accum = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 128; ++j) asm volatile("mov.b32 %0, %0;" : "+r"(s[j]));
    float mx[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) mx[u] = fmaxf(__uint_as_float(s[u]), __uint_as_float(s[8 + u]));
#pragma unroll
    for (int j = 16; j < 128; j += 16)
#pragma unroll
      for (int u = 0; u < 8; ++u) mx[u] = max3(mx[u], __uint_as_float(s[j + u]), __uint_as_float(s[j + 8 + u]));
    const float m = fmaxf(max3(mx[0], mx[1], mx[2]), max3(max3(mx[3], mx[4], mx[5]), mx[6], mx[7]));
    uint32_t pk[64];
    float l;
    if (VARIANT == 0) {
      float l0 = 0.f, l1 = 0.f;
#pragma unroll
      for (int j = 0; j < 64; ++j) {
        const float e0 = ex2_approx(__uint_as_float(s[2 * j]) - m);
        const float e1 = ex2_approx(__uint_as_float(s[2 * j + 1]) - m);
        l0 += e0;
        l1 += e1;
        pk[j] = pack_bf16x2(e0, e1);
      }
      l = l0 + l1;
    } else if (VARIANT == 1) {
      // no separate row sum / no normalisation (as if the tensor core produced the sum and O were scaled at read-back)
#pragma unroll
      for (int j = 0; j < 64; ++j) {
        const float e0 = ex2_approx(__uint_as_float(s[2 * j]) - m);
        const float e1 = ex2_approx(__uint_as_float(s[2 * j + 1]) - m);
        pk[j] = pack_bf16x2(e0, e1);
      }
      l = __uint_as_float(pk[63] ^ pk[17]);
    } else {
      // exps only (MUFU + subtract), everything else dropped: the pure pipe floor seen by this many warps
      float l0 = 0.f;
#pragma unroll
      for (int j = 0; j < 128; ++j) l0 += ex2_approx(__uint_as_float(s[j]) - m);
#pragma unroll
      for (int j = 0; j < 64; ++j) pk[j] = 0;
      l = l0;
    }
    if (VARIANT == 0) {
      const float inv_l = 1.0f / l;
      const uint32_t inv2 = pack_bf16x2(inv_l, inv_l);
#pragma unroll
      for (int j = 0; j < 64; ++j) pk[j] = mul_bf16x2(pk[j], inv2);
    }
    uint32_t x = 0;
#pragma unroll
    for (int j = 0; j < 64; ++j) asm volatile("xor.b32 %0, %0, %1;" : "+r"(x) : "r"(pk[j]));
    accum += l + __uint_as_float(x);
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = accum;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int VARIANT, int WPS>
void run1(const char* name, float* out, long long* cyc) {
  const int iters = 2000;
  const int threads = 128 * WPS;
  softmax_kernel<VARIANT, WPS><<<148, threads>>>(out, cyc, 10);
  softmax_kernel<VARIANT, WPS><<<148, threads>>>(out, cyc, iters);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  long long h;
  CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
  printf("%-28s warps/sub-partition=%d: %7.1f cycles per item per warp, %7.1f sub-partition cycles per item\n", name, WPS,
         (double)h / iters, (double)h / iters / WPS);
}
template <int VARIANT>
void run(const char* name, float* out, long long* cyc) {
  run1<VARIANT, 1>(name, out, cyc);
  run1<VARIANT, 2>(name, out, cyc);
  run1<VARIANT, 3>(name, out, cyc);
  run1<VARIANT, 4>(name, out, cyc);
}

int main() {
  float* out;
  long long* cyc;
  CK(cudaMalloc(&out, 148 * 640 * sizeof(float)));
  CK(cudaMalloc(&cyc, 8));
  run<0>("full softmax item", out, cyc);
  run<1>("no row sum / no normalise", out, cyc);
  run<2>("sub + ex2 + sum only", out, cyc);
  return 0;
}
