// tcgen05.mma issue / completion cost on B200, measured without any address arithmetic in the timed loop
// (the earlier umma_probe mode 7 had an integer modulo per iteration, which dominated its numbers).
//   mma_cost_probe            runs the whole sweep: SS|TS x N x number of distinct accumulators
// One CTA, 128 threads; warp 1 issues 256 MMAs (M=128, K=16, bf16) through the converged-warp helper
// (elect.sync), fully unrolled by 8 with compile-time operand offsets, then commits and waits.
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>

#include "../ampnet_b200/csrc/umma.cuh"

using namespace ampconv;
using namespace ampconv::umma;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

template <bool SS, int N, int NACC>
__global__ void __launch_bounds__(128) cost_kernel(long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp == 1) {
    constexpr uint32_t idesc = idesc_bf16(128, N, 0, SS ? 0 : 1);
    constexpr int STRIDE = N < 32 ? 32 : N;
    const uint64_t da = smem_desc(smem_u32(smem), 16, 1024, LAYOUT_SW128);
    const uint64_t db = smem_desc(smem_u32(smem + 32768), 16, 1024, LAYOUT_SW128);
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 32; ++it) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const uint32_t d = tmem + 256 + (u % NACC) * STRIDE;
        if (SS)
          mma_ss_w(d, desc_advance(da, (u & 3) * 32), desc_advance(db, (u & 3) * 32), idesc, 1);
        else
          mma_ts_w(d, tmem + 8 * u, desc_advance(db, u * 2048), idesc, 1);
      }
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

// Same, but the issuing warp also commits to an mbarrier after every group of G MMAs (as the attention kernels do).
template <int N, int G>
__global__ void __launch_bounds__(128) commit_kernel(long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar, bars[4];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) { mbar_init(&bar, 1); for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp == 1) {
    constexpr uint32_t idesc = idesc_bf16(128, N, 0, 1);
    const uint64_t db = smem_desc(smem_u32(smem + 32768), 16, 1024, LAYOUT_SW128);
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 256 / G; ++it) {
#pragma unroll
      for (int u = 0; u < G; ++u) mma_ts_w(tmem + 256, tmem + 8 * (u & 7), desc_advance(db, (u & 7) * 2048), idesc, 1);
      mma_commit_w(&bars[it & 3]);
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

// Mixed stream as in the backward kernels: per iteration 2 SS MMAs (N = 64, scores) + 8 TS MMAs (N = 16, consumers),
// optionally with a commit after each group.  Reveals any cost of switching instruction descriptors.
template <int PATTERN>
__global__ void __launch_bounds__(128) mixed_kernel(long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar, bars[4];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) { mbar_init(&bar, 1); for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp == 1) {
    constexpr uint32_t id_ss = idesc_bf16(128, 64, 0, 0), id_ts = idesc_bf16(128, 16, 0, 1);
    const uint64_t da = smem_desc(smem_u32(smem), 16, 1024, LAYOUT_SW128);
    const uint64_t db = smem_desc(smem_u32(smem + 32768), 16, 1024, LAYOUT_SW128);
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 32; ++it) {
      if (PATTERN == 0 || PATTERN == 2) {       // 2 SS then 8 TS
        mma_ss_w(tmem, da, db, id_ss, 0);
        mma_ss_w(tmem + 64, da, db, id_ss, 0);
        if (PATTERN == 2) mma_commit_w(&bars[0]);
#pragma unroll
        for (int u = 0; u < 8; ++u) mma_ts_w(tmem + 384 + 16 * (u >> 2), tmem + 256 + 8 * (u & 3), desc_advance(db, (u & 3) * 2048), id_ts, 1);
        if (PATTERN == 2) mma_commit_w(&bars[1]);
      } else if (PATTERN == 1) {                // fully interleaved: SS, 4 TS, SS, 4 TS
        mma_ss_w(tmem, da, db, id_ss, 0);
#pragma unroll
        for (int u = 0; u < 4; ++u) mma_ts_w(tmem + 384, tmem + 256 + 8 * u, desc_advance(db, u * 2048), id_ts, 1);
        mma_ss_w(tmem + 64, da, db, id_ss, 0);
#pragma unroll
        for (int u = 0; u < 4; ++u) mma_ts_w(tmem + 400, tmem + 256 + 8 * u, desc_advance(db, u * 2048), id_ts, 1);
      }
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

template <int PATTERN>
void run_mixed(const char* name, long long* dout) {
  const size_t smem = 65536 + 1024;
  CK(cudaFuncSetAttribute(mixed_kernel<PATTERN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int rep = 0; rep < 2; ++rep) mixed_kernel<PATTERN><<<1, 128, smem>>>(dout);
  CK(cudaDeviceSynchronize());
  long long h[2];
  CK(cudaMemcpy(h, dout, 16, cudaMemcpyDeviceToHost));
  printf("RESULT mixed %-34s: issue %7.1f clk, issue+complete %7.1f clk per group of 2 SS(N=64) + 8 TS(N=16)\n", name,
         h[0] / 32.0, h[1] / 32.0);
}

template <bool SS, int N, int NACC>
void run(long long* dout) {
  const size_t smem = 65536 + 1024;
  CK(cudaFuncSetAttribute(cost_kernel<SS, N, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int rep = 0; rep < 2; ++rep) cost_kernel<SS, N, NACC><<<1, 128, smem>>>(dout);
  CK(cudaDeviceSynchronize());
  long long h[2];
  CK(cudaMemcpy(h, dout, 16, cudaMemcpyDeviceToHost));
  printf("RESULT %s N=%3d accumulators=%d: issue %6.1f clk/MMA, issue+complete %6.1f clk/MMA\n", SS ? "SS" : "TS", N, NACC,
         h[0] / 256.0, h[1] / 256.0);
}

template <int N, int G>
void run_commit(long long* dout) {
  const size_t smem = 65536 + 1024;
  CK(cudaFuncSetAttribute(commit_kernel<N, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int rep = 0; rep < 2; ++rep) commit_kernel<N, G><<<1, 128, smem>>>(dout);
  CK(cudaDeviceSynchronize());
  long long h[2];
  CK(cudaMemcpy(h, dout, 16, cudaMemcpyDeviceToHost));
  printf("RESULT TS N=%3d commit every %d MMAs: issue %6.1f clk/MMA, issue+complete %6.1f clk/MMA\n", N, G, h[0] / 256.0,
         h[1] / 256.0);
}

int main() {
  long long* dout;
  CK(cudaMalloc(&dout, 16));
  run<false, 16, 1>(dout); run<false, 16, 2>(dout); run<false, 16, 4>(dout);
  run<false, 32, 1>(dout); run<false, 32, 2>(dout);
  run<false, 64, 1>(dout); run<false, 64, 2>(dout);
  run<false, 128, 1>(dout); run<false, 128, 2>(dout);
  run<false, 256, 1>(dout);
  run<true, 16, 1>(dout); run<true, 16, 4>(dout);
  run<true, 64, 1>(dout); run<true, 64, 2>(dout);
  run<true, 128, 1>(dout); run<true, 128, 2>(dout);
  run<true, 256, 1>(dout);
  run_mixed<0>("2 SS then 8 TS", dout);
  run_mixed<1>("SS, 4 TS, SS, 4 TS", dout);
  run_mixed<2>("2 SS, commit, 8 TS, commit", dout);
  run_commit<16, 1>(dout); run_commit<16, 2>(dout); run_commit<16, 4>(dout); run_commit<16, 8>(dout);
  run_commit<128, 1>(dout); run_commit<128, 2>(dout);
  return 0;
}
