// bf16 tensor-core family, backward: flash-style recompute of the per-edge attention from the saved
// log-sum-exp (autograd of custom_multihead_attn_forward.py:4140-4186 + PyG scatter-mean, which the
// reference gets from saved [E,H,F,F] probabilities).  Two persistent warp-specialised kernels built
// from one template; both accumulate over a node's edges on chip (no atomics, deterministic):
//
//   MODE_DQ  (destination-sorted): own tiles = Q'_t, dO_t (per destination), edge tiles = K_s, V_s.
//       X = Q_h K_h^T, Y = dO_h V_h^T;  P = exp2(X - lse2_i), W = P o Y, delta_i = sum_j W_ij;
//       TX = P K_h, TY = W K_h (A from TMEM, B = K tile as MN-major operand);
//       dQ_h += TY - delta o TX.                      delta[p,h,i] is written for MODE_DKV.
//   MODE_DKV (source-sorted): own tiles = K_s, V_s (per source), edge tiles = Q'_t, dO_t plus the
//       lse2 / delta rows of the edge (bulk copies).  Everything is transposed (thread = source token):
//       X = K_h Q_h^T, Y = V_h dO_h^T;  P^T = exp2(X - lse2_col), dS^T = P^T o (Y - delta_col);
//       TX = P^T dO_h -> dV_h,  TY = dS^T Q'_h -> dK_h.
//
// Work split: an item = (edge, head).  Its two score tiles X, Y live in one of two TMEM sets (item parity:
// columns [256s, 256s+256), X at +0, Y at +128), so the score MMAs of item c+1 run while item c is being
// processed.  All 16 elementwise warps work on the same item (4 per SM sub-partition, to hide tcgen05.ld and
// MUFU latency): warp w owns TMEM lane quarter w & 3 and score columns [32g, 32g+32), g = w >> 2.  The bf16
// operands P / W (or P^T / dS^T) overwrite the first 16 columns of each group's own fp32 scores in place, so a
// write can never pass another warp's read; TX / TY land in the 16-column holes at +16 (and +48 for hd = 32).
// tcgen05 issue costs ~100 clk per instruction, so three converged warps issue: X/Y, TX and TY.
#include <cuda_bf16.h>
#include <math_constants.h>

#include "common.cuh"
#include "umma.cuh"

namespace ampconv {
namespace {

using namespace umma;

constexpr int kD = 64;
constexpr int kTileBytes = 128 * 128;
constexpr int kEwWarps = 16;     // elementwise warps: 4 TMEM lane quarters x 4 column groups of 32 score columns
constexpr int kTWarps = 4;       // T-MMA issuing warps: (TX | TY) x (K steps 0-3 | 4-7)
constexpr int kThreads = (kEwWarps + 2 + kTWarps) * 32;   // + producer warp + score-MMA warp + T warps
constexpr int MODE_DQ = 0, MODE_DKV = 1;
constexpr int kStatFloats = 4 * 128;   // H * roundup4(F) <= 512 floats per edge and statistic

struct NodeSlot {
  int node, e_begin, e_end;
};

template <int MODE>
struct BwdSmem {
  static constexpr int NS = MODE == MODE_DQ ? 3 : 2;
  static constexpr int NACC = MODE == MODE_DQ ? 1 : 2;
  uint8_t own[2][2][kTileBytes];        // [slot][tile 0/1]
  uint8_t edge[NS][2][kTileBytes];      // [stage][tile 0/1]
  float stat[NS][2][kStatFloats];       // MODE_DKV: lse2 / delta rows of the edge
  float acc[NACC * 16][512];            // [accumulator element][elementwise thread]
  uint64_t own_full[2], own_empty[2];
  uint64_t edge_full[NS], edge_empty[NS];
  uint64_t xy_full[2], xy_empty[2], u_full[2], t_full[2];
  float dl[2][4][128];                  // MODE_DQ: partial delta of [set][column group][row]
  NodeSlot slot[2];
  uint32_t tmem_base;
};

#define AMP_FAIL(code)                                                       \
  do {                                                                       \
    atomicCAS(status, 0, (int)((code) | (blockIdx.x << 16)));                \
    goto fail;                                                               \
  } while (0)
#define AMP_WAIT(bar, parity, code)                         \
  do {                                                      \
    if (!mbar_wait((bar), (parity))) AMP_FAIL(code);        \
  } while (0)

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// own0/own1: tensor maps of the per-node tiles, oth0/oth1: of the per-edge tiles.
// rowptr/nbr: CSR of the pass (by destination for MODE_DQ, by source for MODE_DKV); slot_of[e] = position of
// edge e in the statistics arrays (NULL: identity).  d_qkv: fp32 [rows, 192].
template <int HD, int MODE, bool PROF>
__global__ void __launch_bounds__(kThreads, 1)
attn_bwd_bf16_kernel(const __grid_constant__ CUtensorMap own0, const __grid_constant__ CUtensorMap own1,
                     const __grid_constant__ CUtensorMap oth0, const __grid_constant__ CUtensorMap oth1,
                     const int32_t* __restrict__ rowptr, const int32_t* __restrict__ nbr,
                     const int32_t* __restrict__ slot_of, const int32_t* __restrict__ order,
                     const float* __restrict__ lse2, float* __restrict__ delta, float* __restrict__ d_qkv, int* __restrict__ counter, int* __restrict__ status,
                     int N, int F, float out_scale0, float out_scale1, int out_ld, int out_c0, int out_c1,
                     long long* __restrict__ prof) {
  using Smem = BwdSmem<MODE>;
  constexpr int NS = Smem::NS;
  constexpr int H = kD / HD;
  extern __shared__ uint8_t smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Fs = (F + 3) & ~3;   // row stride of the statistics arrays

  if (warp == kEwWarps + 1) tmem_alloc(&sm.tmem_base, 512);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sm.own_full[i], 1);
      mbar_init(&sm.own_empty[i], 1 + kEwWarps + kTWarps);
      mbar_init(&sm.xy_full[i], 1);
      mbar_init(&sm.xy_empty[i], kEwWarps);
      mbar_init(&sm.u_full[i], kEwWarps);
      mbar_init(&sm.t_full[i], kTWarps);
    }
    for (int i = 0; i < NS; ++i) {
      mbar_init(&sm.edge_full[i], 1);
      mbar_init(&sm.edge_empty[i], MODE == MODE_DKV ? 1 + kTWarps + kEwWarps : 1 + kTWarps);
    }
    fence_barrier_init();
  }
  if (warp == kEwWarps && lane == 0) {
    prefetch_tensormap(&own0);
    prefetch_tensormap(&own1);
    prefetch_tensormap(&oth0);
    prefetch_tensormap(&oth1);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;
  const int nqk = ((F + 15) >> 4) << 4;
  const int ksteps = (F + 15) >> 4;
  const uint32_t stat_bytes = (uint32_t)(H * Fs * sizeof(float));

  if (warp == kEwWarps) {
    // ------------------------------------------------------------------ producer / scheduler
    uint32_t qi = 0, ei = 0;
    for (;;) {
      int node = -1, eb = 0, ee = 0;
      if (lane == 0) {
        const int idx = atomicAdd(counter, 1);
        node = idx < N ? (order ? order[idx] : idx) : -1;
        if (node >= 0) {
          eb = rowptr[node];
          ee = rowptr[node + 1];
        }
      }
      node = __shfl_sync(0xffffffffu, node, 0);
      eb = __shfl_sync(0xffffffffu, eb, 0);
      ee = __shfl_sync(0xffffffffu, ee, 0);
      if (node >= 0 && ee == eb) {
        // node without edges in this pass: its gradient rows are zero
        const int nblk = MODE == MODE_DQ ? 1 : 2;
        for (int i = lane; i < F * nblk * (kD / 4); i += 32) {
          const int r = i / (nblk * (kD / 4)), rem = i - r * (nblk * (kD / 4));
          const int blk = rem / (kD / 4), c4 = rem - blk * (kD / 4);
          *reinterpret_cast<float4*>(d_qkv + ((int64_t)node * F + r) * out_ld + (blk == 0 ? out_c0 : out_c1) + 4 * c4) =
              make_float4(0.f, 0.f, 0.f, 0.f);
        }
        continue;
      }
      int failed = 0;
      if (lane == 0) {
        const uint32_t qb = qi & 1;
        if (!mbar_wait(&sm.own_empty[qb], ((qi >> 1) & 1) ^ 1)) {
          failed = 101;
        } else {
          NodeSlot ns;
          ns.node = node;
          ns.e_begin = eb;
          ns.e_end = ee;
          sm.slot[qb] = ns;
          if (node < 0) {
            mbar_arrive(&sm.own_full[qb]);
          } else {
            mbar_arrive_expect_tx(&sm.own_full[qb], 2 * kTileBytes);
            tma_load_3d(sm.own[qb][0], &own0, &sm.own_full[qb], 0, 0, node);
            tma_load_3d(sm.own[qb][1], &own1, &sm.own_full[qb], 0, 0, node);
            int nb_next = nbr[eb];
            for (int e = eb; e < ee; ++e, ++ei) {
              const int nb = nb_next;
              if (e + 1 < ee) nb_next = nbr[e + 1];
              const uint32_t st = ei % NS;
              if (!mbar_wait(&sm.edge_empty[st], ((ei / NS) & 1) ^ 1)) {
                failed = 102;
                break;
              }
              if (MODE == MODE_DKV) {
                const int64_t sl = slot_of ? slot_of[e] : e;
                mbar_arrive_expect_tx(&sm.edge_full[st], 2 * kTileBytes + 2 * stat_bytes);
                bulk_load(sm.stat[st][0], lse2 + sl * H * Fs, stat_bytes, &sm.edge_full[st]);
                bulk_load(sm.stat[st][1], delta + sl * H * Fs, stat_bytes, &sm.edge_full[st]);
              } else {
                mbar_arrive_expect_tx(&sm.edge_full[st], 2 * kTileBytes);
              }
              tma_load_3d(sm.edge[st][0], &oth0, &sm.edge_full[st], 0, 0, nb);
              tma_load_3d(sm.edge[st][1], &oth1, &sm.edge_full[st], 0, 0, nb);
            }
          }
          ++qi;
        }
      }
      failed = __shfl_sync(0xffffffffu, failed, 0);
      if (failed) AMP_FAIL(failed);
      if (node < 0) break;
    }
  } else if (warp == kEwWarps + 1) {
    // ------------------------------------------------------------------ score MMAs X, Y of every item
    {
      const uint32_t idesc_xy = idesc_bf16(128, nqk, 0, 0);
      uint32_t qi = 0, edge = 0, c = 0;
      const bool do_prof = PROF && blockIdx.x == 0;
      long long pm[4] = {0, 0, 0, 0};
      long long tp = do_prof ? clock64() : 0;
#define AMP_MPHASE(i) do { if (do_prof) { const long long now_ = clock64(); pm[i] += now_ - tp; tp = now_; } } while (0)
      for (;; ++qi) {
        const uint32_t qb = qi & 1;
        AMP_WAIT(&sm.own_full[qb], (qi >> 1) & 1, 201);
        const NodeSlot ns = sm.slot[qb];
        if (ns.node < 0) break;
        AMP_MPHASE(3);
        for (int e = ns.e_begin; e < ns.e_end; ++e, ++edge) {
          const uint32_t st = edge % NS;
          AMP_WAIT(&sm.edge_full[st], (edge / NS) & 1, 202);
          AMP_MPHASE(0);
#pragma unroll
          for (int h = 0; h < H; ++h, ++c) {
            const uint32_t set = c & 1;
            AMP_WAIT(&sm.xy_empty[set], ((c >> 1) & 1) ^ 1, 203);
            AMP_MPHASE(1);
            tc_fence_after();
            const uint32_t hb = h * (HD * 2);
            const uint64_t a0 = smem_desc(smem_u32(sm.own[qb][0]) + hb, 16, 1024, LAYOUT_SW128);
            const uint64_t a1 = smem_desc(smem_u32(sm.own[qb][1]) + hb, 16, 1024, LAYOUT_SW128);
            const uint64_t b0 = smem_desc(smem_u32(sm.edge[st][0]) + hb, 16, 1024, LAYOUT_SW128);
            const uint64_t b1 = smem_desc(smem_u32(sm.edge[st][1]) + hb, 16, 1024, LAYOUT_SW128);
#pragma unroll
            for (int ks = 0; ks < HD / 16; ++ks)
              mma_ss_w(tmem + set * 256, desc_advance(a0, ks * 32), desc_advance(b0, ks * 32), idesc_xy, ks > 0);
#pragma unroll
            for (int ks = 0; ks < HD / 16; ++ks)
              mma_ss_w(tmem + set * 256 + 128, desc_advance(a1, ks * 32), desc_advance(b1, ks * 32), idesc_xy, ks > 0);
            mma_commit_w(&sm.xy_full[set]);
            AMP_MPHASE(2);
          }
          mma_commit_w(&sm.edge_empty[st]);
          if (e + 1 == ns.e_end) mma_commit_w(&sm.own_empty[qb]);
        }
      }
      if (do_prof && lane == 0) {
        for (int i = 0; i < 4; ++i) prof[32 + i] = pm[i];
        prof[36] = c;
      }
#undef AMP_MPHASE
    }
  } else if (warp >= kEwWarps + 2) {
    // ------------------------------------------------------------------ T MMAs: one warp issues TX = X' * B_tx, the other TY = Y' * B_ty
    //   (A = the bf16 operand the elementwise warps wrote back into TMEM, B = an edge tile as MN-major operand)
    {
      const uint32_t tw = warp - (kEwWarps + 2);
      const uint32_t which = tw >> 1;                          // 0: TX, 1: TY
      const int ks0 = 4 * (tw & 1);                            // this warp's K steps: [ks0, ks0 + 4)
      const uint32_t idesc_t = idesc_bf16(128, 16, 0, 1);      // N = 16 per MMA (two per K step when hd = 32)
      const int btile = (which == 0 && MODE == MODE_DKV) ? 1 : 0;   // TX of the dK/dV pass multiplies dO; all others tile 0
      uint32_t qi = 0, edge = 0, c = 0;
      const bool do_prof = PROF && blockIdx.x == 0 && tw == 0;
      long long pm[4] = {0, 0, 0, 0};
      long long tp = do_prof ? clock64() : 0;
#define AMP_TPHASE(i) do { if (do_prof) { const long long now_ = clock64(); pm[i] += now_ - tp; tp = now_; } } while (0)
      for (;; ++qi) {
        const uint32_t qb = qi & 1;
        AMP_WAIT(&sm.own_full[qb], (qi >> 1) & 1, 211);
        const NodeSlot ns = sm.slot[qb];
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.own_empty[qb]);
        if (ns.node < 0) break;
        for (int e = ns.e_begin; e < ns.e_end; ++e, ++edge) {
          const uint32_t st = edge % NS;
#pragma unroll
          for (int h = 0; h < H; ++h, ++c) {
            const uint32_t set = c & 1;
            AMP_TPHASE(3);
            AMP_WAIT(&sm.u_full[set], (c >> 1) & 1, 213);
            AMP_TPHASE(0);
            tc_fence_after();
            const uint32_t a_col = tmem + set * 256 + which * 128;
            const uint64_t bd = smem_desc(smem_u32(sm.edge[st][btile]) + h * (HD * 2), 16, 1024, LAYOUT_SW128);
            // every T MMA accumulates: the elementwise warps zeroed the T tiles, so the K steps of one tile may be
            // issued by two warps in any order
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const int ks = ks0 + kk;
              if (ks < ksteps) {
                const uint32_t a_addr = a_col + 32 * (ks >> 1) + 8 * (ks & 1);
#pragma unroll
                for (int half = 0; half < HD / 16; ++half)
                  mma_ts_w(a_col + 16 + 32 * half, a_addr, desc_advance(bd, ks * 2048 + half * 32), idesc_t, 1u);
              }
            }
            mma_commit_w(&sm.t_full[set]);
            AMP_TPHASE(1);
            if (do_prof) {
              AMP_WAIT(&sm.t_full[set], (c >> 1) & 1, 214);   // observe only: how long until all four T warps' MMAs have landed
              AMP_TPHASE(2);
            }
          }
          mma_commit_w(&sm.edge_empty[st]);
        }
      }
      if (do_prof && lane == 0) {
        for (int i = 0; i < 4; ++i) prof[40 + i] = pm[i];
        prof[44] = c;
      }
#undef AMP_TPHASE
    }
  } else {
    // ------------------------------------------------------------------ elementwise warps
    const uint32_t g = warp >> 2;                   // column group: score columns [32g, 32g+32)
    const int row = (warp & 3) * 32 + lane;
    const bool row_ok = row < F;
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    constexpr int HQ = HD / 4;                      // T columns this thread reads back per head
    uint32_t qi = 0, c = 0, ei = 0;
    float* acc = &sm.acc[0][threadIdx.x];           // element x of this thread: acc[x * 512]
    const bool do_prof = PROF && blockIdx.x == 0 && (warp == 0 || warp == kEwWarps - 1);
    long long pt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tp = do_prof ? clock64() : 0;
#define AMP_PHASE(i) do { if (do_prof) { const long long now_ = clock64(); pt[i] += now_ - tp; tp = now_; } } while (0)
    for (;;) {
      const uint32_t qb = qi & 1;
      AMP_WAIT(&sm.own_full[qb], (qi >> 1) & 1, 301);
      AMP_PHASE(6);
      const NodeSlot ns = sm.slot[qb];
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.own_empty[qb]);
      if (ns.node < 0) break;
#pragma unroll
      for (int x = 0; x < Smem::NACC * H * HQ; ++x) acc[x * 512] = 0.f;
      uint32_t t = 0;
      int e_prev = 0;
      // folds this thread's HQ columns of TX / TY of item cc (head hp, edge slot ep) into the accumulators, frees the set
      auto readback = [&](uint32_t cc, int hp, int ep) -> bool {
        const uint32_t set = cc & 1;
        if (!mbar_wait(&sm.t_full[set], (cc >> 1) & 1)) return false;
        AMP_PHASE(3);
        tc_fence_after();
        uint32_t tx[HQ], ty[HQ];
        // logical T column HQ*g maps to +16 + col (hd = 16) or to the two 16-column holes +16 / +48 (hd = 32)
        const uint32_t tcol = HQ * g;
        const uint32_t base = lane_base + set * 256 + (tcol < 16 ? 16 + tcol : 48 + (tcol - 16));
        if constexpr (HQ == 4) {
          tmem_ld_32x32b_x4(base, tx);
          tmem_ld_32x32b_x4(base + 128, ty);
        } else {
          tmem_ld_32x32b_x8(base, tx);
          tmem_ld_32x32b_x8(base + 128, ty);
        }
        tmem_ld_wait();
        float dl = 0.f;
        if (MODE == MODE_DQ) dl = (sm.dl[set][0][row] + sm.dl[set][1][row]) + (sm.dl[set][2][row] + sm.dl[set][3][row]);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.xy_empty[set]);
        float* a = acc + hp * HQ * 512;
        if (MODE == MODE_DQ) {
          if (g == 0 && row_ok) delta[((int64_t)ep * H + hp) * Fs + row] = dl;
#pragma unroll
          for (int x = 0; x < HQ; ++x) a[x * 512] += __uint_as_float(ty[x]) - dl * __uint_as_float(tx[x]);
        } else {
          float* a2 = a + H * HQ * 512;
#pragma unroll
          for (int x = 0; x < HQ; ++x) {
            a[x * 512] += __uint_as_float(ty[x]);    // dK
            a2[x * 512] += __uint_as_float(tx[x]);   // dV
          }
        }
        AMP_PHASE(4);
        return true;
      };
      for (int e = ns.e_begin; e < ns.e_end; ++e, ++ei) {
        const uint32_t st = ei % NS;
        if (MODE == MODE_DKV) AMP_WAIT(&sm.edge_full[st], (ei / NS) & 1, 306);   // acquire the bulk-copied statistics rows
#pragma unroll
        for (int h = 0; h < H; ++h) {
          const uint32_t set = c & 1;
          float L = 0.f;
          if (MODE == MODE_DQ) L = row_ok ? lse2[((int64_t)e * H + h) * Fs + row] : 0.f;
          AMP_PHASE(5);
          AMP_WAIT(&sm.xy_full[set], (c >> 1) & 1, 302);
          AMP_PHASE(0);
          tc_fence_after();
          const float* Ls = sm.stat[st][0] + h * Fs;
          const float* Ds = sm.stat[st][1] + h * Fs;
          const uint32_t xbase = lane_base + set * 256 + 32 * g;
          float dl0 = 0.f, dl1 = 0.f;
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int col0 = 32 * g + 16 * k;          // first score column of this sub-chunk
            if (col0 < nqk) {
              uint32_t xs[16], ys[16];
              tmem_ld_32x32b_x16(xbase + 16 * k, xs);
              tmem_ld_32x32b_x16(xbase + 128 + 16 * k, ys);
              tmem_ld_wait();
              uint32_t px[8], py[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int c0 = col0 + 2 * j;
                float p0, p1, u0, u1;
                if (MODE == MODE_DQ) {
                  p0 = ex2_approx(__uint_as_float(xs[2 * j]) - L);
                  p1 = ex2_approx(__uint_as_float(xs[2 * j + 1]) - L);
                  u0 = p0 * __uint_as_float(ys[2 * j]);
                  u1 = p1 * __uint_as_float(ys[2 * j + 1]);
                } else {
                  const float2 l2 = *reinterpret_cast<const float2*>(Ls + c0);
                  const float2 d2 = *reinterpret_cast<const float2*>(Ds + c0);
                  p0 = ex2_approx(__uint_as_float(xs[2 * j]) - l2.x);
                  p1 = ex2_approx(__uint_as_float(xs[2 * j + 1]) - l2.y);
                  u0 = p0 * (__uint_as_float(ys[2 * j]) - d2.x);
                  u1 = p1 * (__uint_as_float(ys[2 * j + 1]) - d2.y);
                }
                if (F < 128) {
                  if (c0 >= F) { p0 = 0.f; u0 = 0.f; }
                  if (c0 + 1 >= F) { p1 = 0.f; u1 = 0.f; }
                }
                dl0 += u0;
                dl1 += u1;
                px[j] = pack_bf16x2(p0, p1);
                py[j] = pack_bf16x2(u0, u1);
              }
              // packed columns +8k of this group's own 32 columns: always behind this thread's own reads
              tmem_st_32x32b_x8(xbase + 8 * k, px);
              tmem_st_32x32b_x8(xbase + 128 + 8 * k, py);
            }
            // Half way through the item the previous item's T tiles are complete: fold them in now and release its
            // TMEM set, so that the score MMAs of the next item are issued while the second sub-chunk is processed.
            if (k == 0 && t > 0) {
              if (!readback(c - 1, (h + H - 1) % H, e_prev)) AMP_FAIL(304);
            }
          }
          // zero the T tiles of this set (they live in the second half of this group's own, already consumed, score columns)
          if (g < HD / 16) {
            const uint32_t z[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
            tmem_st_32x32b_x16(xbase + 16, z);
            tmem_st_32x32b_x16(xbase + 128 + 16, z);
          }
          if (MODE == MODE_DQ) sm.dl[set][g][row] = dl0 + dl1;
          AMP_PHASE(1);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm.u_full[set]);
          if (MODE == MODE_DKV && h == H - 1) {
            if (lane == 0) mbar_arrive(&sm.edge_empty[st]);   // this warp no longer reads the stage's statistics rows
          }
          AMP_PHASE(2);
          e_prev = e;
          ++c;
          ++t;
        }
      }
      if (t > 0) {
        if (!readback(c - 1, H - 1, e_prev)) AMP_FAIL(305);
      }
      if (row_ok) {
        float* o = d_qkv + ((int64_t)ns.node * F + row) * out_ld;
#pragma unroll
        for (int h = 0; h < H; ++h) {
          const float* a = acc + h * HQ * 512;
          float4* o0 = reinterpret_cast<float4*>(o + out_c0 + h * HD + HQ * g);
#pragma unroll
          for (int x = 0; x < HQ; x += 4)
            o0[x >> 2] = make_float4(a[x * 512] * out_scale0, a[(x + 1) * 512] * out_scale0,
                                     a[(x + 2) * 512] * out_scale0, a[(x + 3) * 512] * out_scale0);
          if (MODE == MODE_DKV) {
            const float* a2 = a + H * HQ * 512;
            float4* o1 = reinterpret_cast<float4*>(o + out_c1 + h * HD + HQ * g);
#pragma unroll
            for (int x = 0; x < HQ; x += 4)
              o1[x >> 2] = make_float4(a2[x * 512] * out_scale1, a2[(x + 1) * 512] * out_scale1,
                                       a2[(x + 2) * 512] * out_scale1, a2[(x + 3) * 512] * out_scale1);
          }
        }
      }
      AMP_PHASE(7);
      ++qi;
    }
    if (do_prof && lane == 0) {
      long long* pr = prof + (warp == 0 ? 0 : 16);
      for (int i = 0; i < 8; ++i) pr[i] = pt[i];
      pr[8] = c;
    }
#undef AMP_PHASE
  }
fail:
  tc_fence_before();
  __syncthreads();
  if (warp == kEwWarps + 1) tmem_dealloc(tmem, 512);
}

long long* g_bwd_prof = nullptr;   // debug: set through ampconv_debug_set_bwd_profile

template <int HD, int MODE>
int launch_bwd(const CUtensorMap& own0, const CUtensorMap& own1, const CUtensorMap& oth0, const CUtensorMap& oth1,
               const int32_t* rowptr, const int32_t* nbr, const int32_t* slot_of, const int32_t* order, const float* lse2,
               float* delta, float* d_qkv, int* counter, int* status, int N, int F, float s0, float s1, int out_ld, int out_c0,
               int out_c1, cudaStream_t stream) {
  const size_t smem = sizeof(BwdSmem<MODE>) + 1024;
  const int grid = N < sm_count() ? N : sm_count();
  long long* prof = g_bwd_prof;
  if (prof) {
    AMPCONV_CUDA_TRY(cudaFuncSetAttribute(attn_bwd_bf16_kernel<HD, MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_bwd_bf16_kernel<HD, MODE, true><<<grid, kThreads, smem, stream>>>(own0, own1, oth0, oth1, rowptr, nbr, slot_of, order, lse2, delta,
                                                                          d_qkv, counter, status, N, F, s0, s1, out_ld, out_c0, out_c1, prof);
  } else {
    AMPCONV_CUDA_TRY(cudaFuncSetAttribute(attn_bwd_bf16_kernel<HD, MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_bwd_bf16_kernel<HD, MODE, false><<<grid, kThreads, smem, stream>>>(own0, own1, oth0, oth1, rowptr, nbr, slot_of, order, lse2, delta,
                                                                           d_qkv, counter, status, N, F, s0, s1, out_ld, out_c0, out_c1, nullptr);
  }
  AMPCONV_CHECK_LAUNCH();
  return AMPCONV_OK;
}

}  // namespace
}  // namespace ampconv

using namespace ampconv;

extern "C" int ampconv_attn_bf16_supported(int F, int d, int H);

// N_own: nodes the pass iterates over (destinations for MODE_DQ, sources for MODE_DKV); N_oth: nodes of the edge tiles.
static int bwd_common(int mode, const void* q, const void* k, const void* v, const void* d_agg_bf16,
                      const int32_t* rowptr, const int32_t* nbr, const int32_t* slot_of, const int32_t* order,
                      const float* lse2, float* delta, float* out, int out_ld, int out_c0, int out_c1, int64_t N_dst, int64_t N_kv, int64_t E, int F, int d,
                      int H, void* workspace, size_t workspace_bytes, void* stream_) {
  AMPCONV_REQUIRE(N_dst >= 0 && N_kv >= 0 && E >= 0 && F > 0 && d > 0 && H > 0 && d % H == 0);
  if (!ampconv_attn_bf16_supported(F, d, H)) return AMPCONV_ERR_UNSUPPORTED;
  const int64_t N_own = mode == MODE_DQ ? N_dst : N_kv;
  if (N_own == 0) return AMPCONV_OK;
  AMPCONV_REQUIRE(q && k && v && d_agg_bf16 && rowptr && out && workspace);
  AMPCONV_REQUIRE(E == 0 || (nbr && lse2 && delta));
  if (workspace_bytes < 256) return AMPCONV_ERR_WORKSPACE;
  cudaStream_t stream = as_stream(stream_);
  CUtensorMap mq, mk, mv, mg;
  const int64_t nd = N_dst > 0 ? N_dst : 1, nk = N_kv > 0 ? N_kv : 1;
  if (!make_tensor_map_bf16_3d(&mq, q, kD, F, nd, kD, 128) || !make_tensor_map_bf16_3d(&mk, k, kD, F, nk, kD, 128) ||
      !make_tensor_map_bf16_3d(&mv, v, kD, F, nk, kD, 128) || !make_tensor_map_bf16_3d(&mg, d_agg_bf16, kD, F, nd, kD, 128))
    return AMPCONV_ERR_CUDA;
  int* counter = reinterpret_cast<int*>(workspace) + (mode == MODE_DQ ? 2 : 4);
  int* status = reinterpret_cast<int*>(workspace) + 1;
  AMPCONV_CUDA_TRY(cudaMemsetAsync(counter, 0, sizeof(int), stream));
  const int hd = d / H;
  const float inv_sqrt_hd = 1.0f / sqrtf((float)hd);
  const float ln2 = 0.6931471805599453f;
  if (mode == MODE_DQ) {
    // dQ = hd^-1/2 * (dS K)
    if (hd == 16)
      return launch_bwd<16, MODE_DQ>(mq, mg, mk, mv, rowptr, nbr, slot_of, order, lse2, delta, out, counter, status, (int)N_own, F,
                                    inv_sqrt_hd, 0.f, out_ld, out_c0, out_c1, stream);
    return launch_bwd<32, MODE_DQ>(mq, mg, mk, mv, rowptr, nbr, slot_of, order, lse2, delta, out, counter, status, (int)N_own, F,
                                  inv_sqrt_hd, 0.f, out_ld, out_c0, out_c1, stream);
  }
  // dK = hd^-1/2 * dS^T Q = ln2 * dS^T Q'  (Q' = Q * log2e / sqrt(hd)),  dV = P^T dO
  if (hd == 16)
    return launch_bwd<16, MODE_DKV>(mk, mv, mq, mg, rowptr, nbr, slot_of, order, lse2, delta, out, counter, status, (int)N_own, F,
                                   ln2, 1.f, out_ld, out_c0, out_c1, stream);
  return launch_bwd<32, MODE_DKV>(mk, mv, mq, mg, rowptr, nbr, slot_of, order, lse2, delta, out, counter, status, (int)N_own, F, ln2,
                                 1.f, out_ld, out_c0, out_c1, stream);
}

extern "C" int ampconv_attn_bwd_dq_bf16(const void* q, const void* k, const void* v, const void* d_agg_bf16,
                                        const float* lse2, const int32_t* dst_rowptr, const int32_t* dst_src,
                                        const int32_t* order, float* d_qkv, float* delta, int64_t N, int64_t E, int F,
                                        int d, int H, void* workspace, size_t workspace_bytes, void* stream) {
  return bwd_common(MODE_DQ, q, k, v, d_agg_bf16, dst_rowptr, dst_src, nullptr, order, lse2, delta, d_qkv, 3 * kD, 0, 0, N, N, E, F,
                    d, H, workspace, workspace_bytes, stream);
}

extern "C" int ampconv_attn_bwd_dkv_bf16(const void* q, const void* k, const void* v, const void* d_agg_bf16,
                                         const float* lse2, const float* delta, const int32_t* src_rowptr,
                                         const int32_t* src_dst, const int32_t* src_pos, const int32_t* order,
                                         float* d_qkv, int64_t N, int64_t E, int F, int d, int H,
                                         void* workspace, size_t workspace_bytes, void* stream) {
  return bwd_common(MODE_DKV, q, k, v, d_agg_bf16, src_rowptr, src_dst, src_pos, order, lse2, const_cast<float*>(delta), d_qkv,
                    3 * kD, kD, 2 * kD, N, N, E, F, d, H, workspace, workspace_bytes, stream);
}

// Destination-partitioned variants (multi-GPU): q / d_agg cover the num_nodes local destinations, k / v the
// num_kv_nodes rows of the all-gathered tensors.  _dq writes d_q fp32 [num_nodes*F, 64]; _dkv writes the partial
// d_k | d_v fp32 [num_kv_nodes*F, 128] of the local edges (to be reduce-scattered to the owners).
extern "C" int ampconv_attn_bwd_dq_bf16_part(const void* q, const void* k, const void* v, const void* d_agg_bf16,
                                             const float* lse2, const int32_t* dst_rowptr, const int32_t* dst_src,
                                             const int32_t* order, float* d_q, float* delta, int64_t num_nodes,
                                             int64_t num_kv_nodes, int64_t E, int F, int d, int H, void* workspace,
                                             size_t workspace_bytes, void* stream) {
  return bwd_common(MODE_DQ, q, k, v, d_agg_bf16, dst_rowptr, dst_src, nullptr, order, lse2, delta, d_q, kD, 0, 0, num_nodes,
                    num_kv_nodes, E, F, d, H, workspace, workspace_bytes, stream);
}

extern "C" int ampconv_attn_bwd_dkv_bf16_part(const void* q, const void* k, const void* v, const void* d_agg_bf16,
                                              const float* lse2, const float* delta, const int32_t* src_rowptr,
                                              const int32_t* src_dst, const int32_t* src_pos, const int32_t* order,
                                              float* d_kv, int64_t num_nodes, int64_t num_kv_nodes, int64_t E, int F, int d,
                                              int H, void* workspace, size_t workspace_bytes, void* stream) {
  return bwd_common(MODE_DKV, q, k, v, d_agg_bf16, src_rowptr, src_dst, src_pos, order, lse2, const_cast<float*>(delta), d_kv,
                    2 * kD, 0, kD, num_nodes, num_kv_nodes, E, F, d, H, workspace, workspace_bytes, stream);
}

// Debug: when set to a device buffer of 64 int64, the next backward launches run the instrumented kernel and fill it
// with the cycles one elementwise warp spent per phase (wait X/Y, chunks, publish, wait T, fold T, stats load,
// node wait, node epilogue) and its item count.  NULL restores the product kernels.
extern "C" int ampconv_debug_set_bwd_profile(long long* prof) {
  ampconv::g_bwd_prof = prof;
  return AMPCONV_OK;
}
