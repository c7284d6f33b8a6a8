// TMEM <-> register bandwidth on B200: W warps (W = 4, 8, 16; lane quarter = warp % 4) each issue
// tcgen05.ld.32x32b.x16 (or .x32) back to back over their 32 lanes; prints bytes per clock per SM.
// Also the same with tcgen05.st, and ld while another warp streams SS MMAs into other TMEM columns.
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include "../ampnet_b200/csrc/umma.cuh"
using namespace ampconv;
using namespace ampconv::umma;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

template <int MODE>   // 0: ld x16, 1: ld x32, 2: st x16, 3: ld x16 with concurrent MMAs
__global__ void __launch_bounds__(640) bw_kernel(long long* out, int nwarps, int iters) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint32_t tmem_base_s;
  __shared__ uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  long long t0 = 0, t1 = 0;
  if (warp < nwarps) {
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (MODE == 1) {
        uint32_t r[32];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          tmem_ld_32x32b_x32(lane_base + 32 * ((c + (warp >> 2)) & 7), r);
          tmem_ld_wait();
          acc ^= r[0] ^ r[31];
        }
      } else if (MODE == 2) {
        uint32_t r[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) r[j] = acc + j;
#pragma unroll
        for (int c = 0; c < 8; ++c) tmem_st_32x32b_x16(lane_base + 16 * ((c + 2 * (warp >> 2)) & 15), r);
        tmem_st_wait();
      } else {
        uint32_t r[4][16];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld_32x32b_x16(lane_base + 16 * ((c + 4 * (warp >> 2)) & 15), r[c]);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 4; ++c) acc ^= r[c][0] ^ r[c][15];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld_32x32b_x16(lane_base + 16 * ((c + 4 * (warp >> 2)) & 15), r[c]);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 4; ++c) acc ^= r[c][0] ^ r[c][15];
      }
    }
    t1 = clock64();
  } else if (MODE == 3 && warp == 17) {
    const uint32_t idesc = idesc_bf16(128, 64, 0, 0);
    const uint64_t da = smem_desc(smem_u32(smem), 16, 1024, LAYOUT_SW128);
    const uint64_t db = smem_desc(smem_u32(smem + 32768), 16, 1024, LAYOUT_SW128);
    for (int it = 0; it < iters * 2; ++it) {
      mma_ss_w(tmem + 256, da, db, idesc, 0);
      mma_ss_w(tmem + 320, da, db, idesc, 0);
      mma_ss_w(tmem + 384, da, db, idesc, 0);
      mma_ss_w(tmem + 448, da, db, idesc, 0);
    }
    mma_commit_w(&bar);
    mbar_wait(&bar, 0);
  }
  if (tid == 0) { out[0] = t1 - t0; }
  if (acc == 0x12345) out[1] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int MODE>
void run(const char* name, long long* dout) {
  const size_t smem = 65536 + 1024;
  CK(cudaFuncSetAttribute(bw_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int iters = 2000;
  for (int nw : {1, 4, 8, 16}) {
    bw_kernel<MODE><<<1, 640, smem>>>(dout, nw, 10);
    bw_kernel<MODE><<<1, 640, smem>>>(dout, nw, iters);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    long long h;
    CK(cudaMemcpy(&h, dout, 8, cudaMemcpyDeviceToHost));
    const double bytes = (double)nw * iters * (MODE == 1 ? 4 * 4096.0 : 8 * 2048.0);
    printf("%-26s warps=%2d: %8.1f bytes/clk/SM  (%6.1f clk per 2 KB warp access)\n", name, nw, bytes / h, h / (bytes / nw / 2048.0));
  }
}

int main() {
  long long* dout;
  CK(cudaMalloc(&dout, 16));
  run<0>("tcgen05.ld x16", dout);
  run<1>("tcgen05.ld x32 (wait each)", dout);
  run<2>("tcgen05.st x16", dout);
  run<3>("tcgen05.ld x16 + MMA stream", dout);
  return 0;
}
