// bf16 tensor-core family, backward: flash-style recompute of the per-edge attention from the saved
// log-sum-exp (autograd of custom_multihead_attn_forward.py:4140-4186 + PyG scatter-mean, which the
// reference gets from saved [E,H,F,F] probabilities).  Two persistent warp-specialised kernels built
// from one template; both accumulate over a node's edges IN TENSOR MEMORY (no atomics, deterministic):
//
//   MODE_DQ  (destination-sorted): own tiles = Q'_t, dO_t (per destination), edge tiles = K_s, V_s.
//       X = Q_h K_h^T, Y = dO_h V_h^T;  P = exp2(X - lse2_i), W = P o Y, delta_i = sum_j W_ij;
//       dQ_h = sum_e W_e K_h  -  sum_e delta_e o (P_e K_h):
//       the first sum accumulates in TMEM over the node's edges; P_e K_h (16 columns per edge and head) goes to a
//       small TMEM ring and is folded into registers two items later, when delta_e is known, so that no MMA
//       operand ever waits for a row sum.  delta[p,h,i] is written for MODE_DKV.
//   MODE_DKV (source-sorted): own tiles = K_s, V_s (per source), edge tiles = Q'_t, dO_t plus the
//       lse2 / delta rows of the edge (bulk copies).  Everything is transposed (thread = source token):
//       X = K_h Q_h^T - 1 lse2^T, Y = V_h dO_h^T - 1 delta^T;  P^T = exp2(X), dS^T = P^T o Y;
//       dV_h += P^T dO_h,  dK_h += dS^T Q'_h, both accumulated in TMEM over the node's edges.
//       The column statistics are subtracted BY THE TENSOR CORE: four statistics warps turn the edge's lse2 / delta rows
//       into a bf16 tile S (row = destination token, head h at columns [16h, 16h+4) = -lse_hi, -lse_lo, -delta_hi,
//       -delta_lo with x_hi = bf16(x), x_lo = bf16(x - x_hi): 16 mantissa bits) and every score MMA gets one more
//       K = 16 step against a constant tile of ones (columns 0, 1 for X; 18, 19 for Y).  The elementwise warps are left
//       with exp2, one multiply and the two bf16 packs per score element.
//
// Work split: a half-item = (edge, head, half) covers 64 of the 128 score columns.  Its two fp32 score tiles
// X | Y (64 + 64 columns) live in one of three TMEM sets (columns [128 s, 128 s + 128)); columns [384, 512) hold
// the accumulators (MODE_DKV: dV | dK; MODE_DQ: dQ | the P K ring).  The 16 elementwise warps form two groups of
// 8 that work on alternate half-items, so that one group's tcgen05.ld / st / barrier latencies overlap the other
// group's exponentials (MUFU is the binding pipe: profiles/r01_softmax_pipe.log).  Warp w owns TMEM lane quarter
// w & 3 and 32 score columns of its group's half-item; the bf16 operands overwrite the first 8 of every 16 fp32
// score columns in place, so a write can never pass another warp's read.  Three converged warps issue
// tcgen05.mma (~20 clk per instruction, profiles/r01_mma_cost.log): scores, consumers of the X-side operand,
// consumers of the Y-side operand; a set is recycled as soon as both consumers have committed.
#include <cuda_bf16.h>
#include <math_constants.h>

#include "common.cuh"
#include "umma.cuh"

namespace ampconv {
namespace {

using namespace umma;

constexpr int kD = 64;
constexpr int kTileBytes = 128 * 128;
constexpr int kEwWarps = 16;     // elementwise warps: 4 TMEM lane quarters x 4 column groups of 16 score columns
constexpr int kFoldWarps = 4;    // MODE_DQ: one warp per TMEM lane quarter folds delta o (P K) and writes dQ;
                                 // MODE_DKV: the same four warps build the edge's statistics tile (see below)
// elementwise warps + producer warp + score-MMA warp + 2 consumer-MMA warps + fold / statistics warps
__host__ __device__ constexpr int threads_of(int) { return (kEwWarps + 4 + kFoldWarps) * 32; }
constexpr int kSets = 3;         // TMEM score sets of 128 columns
constexpr int kAccCol = 384;     // accumulator block 0 at [384, 448), block 1 at [448, 512)
constexpr int MODE_DQ = 0, MODE_DKV = 1;
constexpr int kStatFloats = 4 * 128;   // H * roundup4(F) <= 512 floats per edge and statistic
// Share of the exponentials evaluated on the FMA pipe (umma.cuh: ex2_poly2) instead of MUFU: every fourth pair of score
// columns when enabled.  Measured at C4 (profiles/r02_poly_share_ab.log): dK/dV 28.61 -> 28.07 ms with it, dQ 28.56 -> 29.32 ms
// (its elementwise warps have no issue slots to spare), forward at a half share 26.2 -> 28.4 ms: these kernels are bound by
// latency between the pipeline stages, not by MUFU throughput.
#ifndef AMP_POLY_DQ
#define AMP_POLY_DQ 0
#endif
#ifndef AMP_POLY_DKV
#define AMP_POLY_DKV 1
#endif
constexpr bool kPolyShareDq = AMP_POLY_DQ != 0, kPolyShareDkv = AMP_POLY_DKV != 0;

struct NodeSlot {
  int node, e_begin, e_end;
  int grp;     // head group of this work item (GROUPS == 2: head_dim 8 as zero-padded head_dim-16 tiles, see attn_bf16.cu)
};
// Reads a NodeSlot so that the compiler KNOWS the fields are warp-uniform (shfl from lane 0): loop counters and the
// descriptor / TMEM address arithmetic derived from them can then live in uniform registers, and a tcgen05.mma
// issue costs a handful of instructions instead of a register-to-uniform broadcast per operand.
__device__ __forceinline__ NodeSlot uniform_slot(const NodeSlot& s) {
  NodeSlot u;
  u.node = __shfl_sync(0xffffffffu, s.node, 0);
  u.e_begin = __shfl_sync(0xffffffffu, s.e_begin, 0);
  u.e_end = __shfl_sync(0xffffffffu, s.e_end, 0);
  u.grp = __shfl_sync(0xffffffffu, s.grp, 0);
  return u;
}
struct TrueTag { static constexpr bool value = true; };
struct FalseTag { static constexpr bool value = false; };

template <int MODE>
struct BwdSmem {
  static constexpr int NS = 3;
  uint8_t own[2][2][kTileBytes];        // [slot][tile 0/1]
  uint8_t edge[NS][2][kTileBytes];      // [stage][tile 0/1]
  uint8_t cst[MODE == MODE_DKV ? kTileBytes : 1024];         // MODE_DKV: constant selector tile (ones at columns 0,1 and 18,19)
  uint8_t stile[2][MODE == MODE_DKV ? kTileBytes : 1024];    // MODE_DKV: statistics tile of the edge, double-buffered
  static constexpr int NSTAT = MODE == MODE_DKV ? 2 : 1;
  float stat[NS][NSTAT][kStatFloats];   // lse2 (both modes) / delta (MODE_DKV) rows of the edge, bulk-copied by the producer
  float dl[MODE == MODE_DQ ? 8 : 1][4][128];   // MODE_DQ: partial delta of [item & 7][group * 2 + column block][row]
  float racc[MODE == MODE_DQ ? 64 * 128 : 4];   // MODE_DQ: - sum_e delta_e o (P_e K) of the node, [column][row]
  uint64_t own_full[2], own_empty[2];
  uint64_t edge_full[NS], edge_empty[NS];
  uint64_t xy_full[kSets], set_empty[kSets], op_full[kSets];
  uint64_t acc_full, acc_empty;
  uint64_t dl_bar[8][4];                // MODE_DQ: [item & 7][lane quarter]: the quarter's four warps wrote their partial deltas
  uint64_t ko_full[4], ko_empty[4];     // MODE_DQ: ring of P K results (slot = item % (64 / HD))
  uint64_t st_full[2], st_empty[2];     // MODE_DKV: statistics tile written / no longer read by the score MMAs
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
// same, barrier given by its shared-space address (sm32 + constant offset: no address arithmetic at the use)
#define AMP_WAIT_A(addr, parity, code)                      \
  do {                                                      \
    if (!mbar_wait_a((addr), (parity))) AMP_FAIL(code);     \
  } while (0)
#define SM_OFF(member) (sm32 + (uint32_t)offsetof(Smem, member))

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Debug (PROF kernels only): wait accounting into a local uint32_t wtx[] of the calling role
#define AMP_XWAIT(i, bar, parity, code)                                       \
  do {                                                                        \
    if (PROF) {                                                               \
      const uint32_t t0_ = (uint32_t)clock();                                 \
      AMP_WAIT_A(bar, parity, code);                                          \
      wtx[i] += (uint32_t)clock() - t0_;                                      \
    } else {                                                                  \
      AMP_WAIT_A(bar, parity, code);                                          \
    }                                                                         \
  } while (0)

// own0/own1: tensor maps of the per-node tiles, oth0/oth1: of the per-edge tiles.
// rowptr/nbr: CSR of the pass (by destination for MODE_DQ, by source for MODE_DKV); slot_of[e] = position of
// edge e in the statistics arrays (NULL: identity).  d_qkv: fp32 [rows, out_ld].
// GROUPS == 2 (head_dim 8): a work item is (node, head group); tiles arrive zero-padded to 16 columns per head through the
// 4-D tensor maps of umma.cuh (nothing padded in HBM), statistics rows are those of the group's four heads, and the epilogues
// write the 8 real columns of every head to their true place.
template <int HD, int MODE, int NH, int GROUPS, bool PROF>
__global__ void __launch_bounds__(threads_of(MODE), 1)
attn_bwd_bf16_kernel(const __grid_constant__ CUtensorMap own0, const __grid_constant__ CUtensorMap own1,
                     const __grid_constant__ CUtensorMap oth0, const __grid_constant__ CUtensorMap oth1,
                     const int32_t* __restrict__ rowptr, const int32_t* __restrict__ nbr,
                     const int32_t* __restrict__ slot_of, const int32_t* __restrict__ lse_map,
                     const int32_t* __restrict__ order, const float* __restrict__ lse2, float* __restrict__ delta, float* __restrict__ d_qkv, int* __restrict__ counter, int* __restrict__ status,
                     int N, int F, float out_scale0, float out_scale1, int out_ld, int out_c0, int out_c1,
                     uint16_t* __restrict__ halo_bf16, int halo_from, int accumulate, int out_bf16,
                     const uint8_t* __restrict__ gown0, const uint8_t* __restrict__ gown1,
                     const uint8_t* __restrict__ goth0, const uint8_t* __restrict__ goth1, long long* __restrict__ prof) {
  using Smem = BwdSmem<MODE>;
  constexpr int NS = Smem::NS;
  constexpr int H = kD / HD;          // heads of a work item
  constexpr int HT = H * GROUPS;      // heads of the layer (row count of an edge's statistics block)
  extern __shared__ uint8_t smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sm32 = smem_base_opaque(&sm);   // shared-space address of sm, held in a register by the hot loops
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Fs = (F + 3) & ~3;   // row stride of the statistics arrays
  constexpr int W_SCORE = kEwWarps + 1, W_TX = kEwWarps + 2, W_TY = kEwWarps + 3;

  if (warp == W_SCORE) tmem_alloc(&sm.tmem_base, 512);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sm.own_full[i], 1);
      mbar_init(&sm.own_empty[i], 1 + kEwWarps + 2 + kFoldWarps);
      mbar_init(&sm.st_full[i], kFoldWarps);
      mbar_init(&sm.st_empty[i], 1);
    }
    for (int i = 0; i < kSets; ++i) {
      mbar_init(&sm.xy_full[i], 1);
      mbar_init(&sm.set_empty[i], 2);
      mbar_init(&sm.op_full[i], kEwWarps / 2);
    }
    for (int i = 0; i < NS; ++i) {
      mbar_init(&sm.edge_full[i], GROUPS == 2 ? 2 : 1);   // head_dim 8: bulk copies of the statistics + the cp.async unit
      mbar_init(&sm.edge_empty[i], 3 + (MODE == MODE_DQ ? kEwWarps : kFoldWarps));   // the readers of the statistics rows
    }
    for (int i = 0; i < 32; ++i) mbar_init(&sm.dl_bar[i >> 2][i & 3], 4);
    for (int i = 0; i < 4; ++i) {
      mbar_init(&sm.ko_full[i], 1);
      mbar_init(&sm.ko_empty[i], kFoldWarps);
    }
    mbar_init(&sm.acc_full, 2);
    mbar_init(&sm.acc_empty, MODE == MODE_DQ ? kFoldWarps : kEwWarps);
    fence_barrier_init();
  }
  if (MODE == MODE_DKV) {
    // selector tile: row r, 16-byte chunk 0 = (1, 1, 0, ...), chunk 2 = (0, 0, 1, 1, 0, ...) (128B swizzle: chunk c of row r
    // sits at chunk position c ^ (r & 7)); statistics tiles start as zeros, only the even chunks are ever rewritten
    for (int i = threadIdx.x; i < kTileBytes / 16; i += blockDim.x) {
      const int r = i >> 3, c = (i & 7) ^ (r & 7);
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (c == 0) v.x = 0x3f803f80u;      // bf16 1.0 | 1.0
      if (c == 2) v.y = 0x3f803f80u;
      reinterpret_cast<uint4*>(sm.cst)[i] = v;
      reinterpret_cast<uint4*>(sm.stile[0])[i] = make_uint4(0u, 0u, 0u, 0u);
      reinterpret_cast<uint4*>(sm.stile[1])[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async_smem();
  }
  if (GROUPS == 2) {
    // the pad half of every head slot and the rows >= F are written here once and never again
    uint4* z = reinterpret_cast<uint4*>(sm.own[0][0]);
    for (int i = threadIdx.x; i < (4 + 2 * NS) * kTileBytes / 16; i += blockDim.x) z[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
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
  constexpr int nhalf = NH;                          // halves with at least one valid score column (2 when F > 64)
  const uint32_t stat_bytes = (uint32_t)(H * Fs * sizeof(float));

  // register re-partitioning: the four single-lane control warps give registers to the 16 elementwise warps
  if (warp == kEwWarps) {
    // ------------------------------------------------------------------ producer / scheduler
    if (MODE == MODE_DQ) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    else asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
    if constexpr (GROUPS == 1) {
    uint32_t qi = 0, ei = 0;
    for (;;) {
      int node = -1, eb = 0, ee = 0;
      const int grp = 0;
      if (lane == 0) {
        const int ni = atomicAdd(counter, 1);
        node = ni < N ? (order ? order[ni] : ni) : -1;
        if (node >= 0) {
          eb = rowptr[node];
          ee = rowptr[node + 1];
        }
      }
      node = __shfl_sync(0xffffffffu, node, 0);
      eb = __shfl_sync(0xffffffffu, eb, 0);
      ee = __shfl_sync(0xffffffffu, ee, 0);
      if (node >= 0 && ee == eb) {
        // node without edges in this pass: its gradient rows are zero (halo sources always have an edge)
        if (halo_bf16 != nullptr && node >= halo_from) continue;
        if (accumulate) continue;   // a later ring phase: the node's rows already hold the earlier phases' sum
        const int nblk = MODE == MODE_DQ ? 1 : 2;
        if (out_bf16) {
          uint16_t* ob = reinterpret_cast<uint16_t*>(d_qkv);
          for (int i = lane; i < F * nblk * (kD / 8); i += 32) {
            const int r = i / (nblk * (kD / 8)), rem = i - r * (nblk * (kD / 8));
            const int blk = rem / (kD / 8), c8 = rem - blk * (kD / 8);
            *reinterpret_cast<uint4*>(ob + ((int64_t)node * F + r) * out_ld + (blk == 0 ? out_c0 : out_c1) + 8 * c8) =
                make_uint4(0u, 0u, 0u, 0u);
          }
          continue;
        }
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
          ns.grp = grp;
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
              // statistics of the edge: delta lives at the pass's slot of the edge, lse2 where the FORWARD wrote it
              // (lse_map: slot -> forward index; the forward may have run finer ring phases than this pass)
              if (MODE == MODE_DKV) {
                const int64_t sl = slot_of ? slot_of[e] : e;
                const int64_t le = lse_map ? lse_map[sl] : sl;
                mbar_arrive_expect_tx(&sm.edge_full[st], 2 * kTileBytes + 2 * stat_bytes);
                bulk_load(sm.stat[st][0], lse2 + (le * HT + grp * H) * Fs, stat_bytes, &sm.edge_full[st]);
                bulk_load(sm.stat[st][1], delta + (sl * HT + grp * H) * Fs, stat_bytes, &sm.edge_full[st]);
              } else {
                const int64_t le = lse_map ? lse_map[e] : e;
                mbar_arrive_expect_tx(&sm.edge_full[st], 2 * kTileBytes + stat_bytes);
                bulk_load(sm.stat[st][0], lse2 + (le * HT + grp * H) * Fs, stat_bytes, &sm.edge_full[st]);
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
    } else {
    // head_dim 8: tiles are built by the whole warp with cp.async (umma.cuh: load_padded_tile); a load unit (a work item's own
    // tiles, or an edge's tiles) is one commit group whose full barrier is raised once the NEXT unit has been issued
    // (wait_group 1, proxy fence, arrive).  The statistics rows still travel as bulk copies on the same barrier, which
    // therefore counts two arrivals per phase.
    uint32_t qi = 0, ei = 0, pend = 0;
    auto retire = [&](uint32_t next_bar) {
      cp_async_commit();
      if (pend) {
        cp_async_wait<1>();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_a(pend);
      }
      pend = next_bar;
    };
    for (;;) {
      int node = -1, eb = 0, ee = 0, grp = 0;
      if (lane == 0) {
        const int idx = atomicAdd(counter, 1);
        const int ni = idx / GROUPS;
        grp = idx - ni * GROUPS;
        node = ni < N ? (order ? order[ni] : ni) : -1;
        if (node >= 0) {
          eb = rowptr[node];
          ee = rowptr[node + 1];
        }
      }
      node = __shfl_sync(0xffffffffu, node, 0);
      eb = __shfl_sync(0xffffffffu, eb, 0);
      ee = __shfl_sync(0xffffffffu, ee, 0);
      grp = __shfl_sync(0xffffffffu, grp, 0);
      if (node >= 0 && ee == eb) {
        if (halo_bf16 != nullptr && node >= halo_from) continue;
        if (accumulate || grp != 0) continue;   // later ring phase: rows hold the earlier sum; group 0 zero-fills whole rows
        const int nblk = MODE == MODE_DQ ? 1 : 2;
        if (out_bf16) {
          uint16_t* ob = reinterpret_cast<uint16_t*>(d_qkv);
          for (int i = lane; i < F * nblk * (kD / 8); i += 32) {
            const int r = i / (nblk * (kD / 8)), rem = i - r * (nblk * (kD / 8));
            const int blk = rem / (kD / 8), c8 = rem - blk * (kD / 8);
            *reinterpret_cast<uint4*>(ob + ((int64_t)node * F + r) * out_ld + (blk == 0 ? out_c0 : out_c1) + 8 * c8) =
                make_uint4(0u, 0u, 0u, 0u);
          }
          continue;
        }
        for (int i = lane; i < F * nblk * (kD / 4); i += 32) {
          const int r = i / (nblk * (kD / 4)), rem = i - r * (nblk * (kD / 4));
          const int blk = rem / (kD / 4), c4 = rem - blk * (kD / 4);
          *reinterpret_cast<float4*>(d_qkv + ((int64_t)node * F + r) * out_ld + (blk == 0 ? out_c0 : out_c1) + 4 * c4) =
              make_float4(0.f, 0.f, 0.f, 0.f);
        }
        continue;
      }
      const uint32_t qb = qi & 1;
      int failed = 0;
      if (lane == 0) {
        if (!mbar_wait(&sm.own_empty[qb], ((qi >> 1) & 1) ^ 1)) {
          failed = 101;
        } else {
          NodeSlot ns;
          ns.node = node;
          ns.e_begin = eb;
          ns.e_end = ee;
          ns.grp = grp;
          sm.slot[qb] = ns;
          if (node < 0) mbar_arrive(&sm.own_full[qb]);
        }
      }
      failed = __shfl_sync(0xffffffffu, failed, 0);
      if (failed) AMP_FAIL(failed);
      if (node < 0) break;
      load_padded_tile(smem_u32(sm.own[qb][0]), gown0, node, F, grp, lane);
      load_padded_tile(smem_u32(sm.own[qb][1]), gown1, node, F, grp, lane);
      retire(smem_u32(&sm.own_full[qb]));
      for (int e = eb; e < ee; ++e, ++ei) {
        const int nb = nbr[e];
        const uint32_t st = ei % NS;
        if (lane == 0) {
          if (!mbar_wait(&sm.edge_empty[st], ((ei / NS) & 1) ^ 1)) {
            failed = 102;
          } else if (MODE == MODE_DKV) {
            const int64_t sl = slot_of ? slot_of[e] : e;
            const int64_t le = lse_map ? lse_map[sl] : sl;
            mbar_arrive_expect_tx(&sm.edge_full[st], 2 * stat_bytes);
            bulk_load(sm.stat[st][0], lse2 + (le * HT + grp * H) * Fs, stat_bytes, &sm.edge_full[st]);
            bulk_load(sm.stat[st][1], delta + (sl * HT + grp * H) * Fs, stat_bytes, &sm.edge_full[st]);
          } else {
            const int64_t le = lse_map ? lse_map[e] : e;
            mbar_arrive_expect_tx(&sm.edge_full[st], stat_bytes);
            bulk_load(sm.stat[st][0], lse2 + (le * HT + grp * H) * Fs, stat_bytes, &sm.edge_full[st]);
          }
        }
        failed = __shfl_sync(0xffffffffu, failed, 0);
        if (failed) AMP_FAIL(failed);
        load_padded_tile(smem_u32(sm.edge[st][0]), goth0, nb, F, grp, lane);
        load_padded_tile(smem_u32(sm.edge[st][1]), goth1, nb, F, grp, lane);
        retire(smem_u32(&sm.edge_full[st]));
      }
      ++qi;
    }
    retire(0u);
    cp_async_wait<0>();
    }
  } else if (warp == W_SCORE) {
    // ------------------------------------------------------------------ score MMAs X, Y of every half-item
    if (MODE == MODE_DQ) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    else asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
    {
      const uint32_t idesc_xy = idesc_bf16(128, 64, 0, 0);
      uint32_t qi = 0, edge = 0, k = 0;
      for (;; ++qi) {
        const uint32_t qb = qi & 1;
        AMP_WAIT(&sm.own_full[qb], (qi >> 1) & 1, 201);
        const NodeSlot ns = uniform_slot(sm.slot[qb]);
        if (ns.node < 0) break;
        for (int e = ns.e_begin; e < ns.e_end; ++e, ++edge) {
          const uint32_t st = edge % NS;
          AMP_WAIT(&sm.edge_full[st], (edge / NS) & 1, 202);
          const uint32_t sb = edge & 1;
          if (MODE == MODE_DKV) AMP_WAIT(&sm.st_full[sb], (edge >> 1) & 1, 204);   // the edge's statistics tile is written
#pragma unroll
          for (int h = 0; h < H; ++h) {
            for (int half = 0; half < nhalf; ++half, ++k) {
              const uint32_t set = k % kSets;
              AMP_WAIT(&sm.set_empty[set], ((k / kSets) & 1) ^ 1, 203);
              tc_fence_after();
              const uint32_t hb = h * (HD * 2);
              const uint64_t a0 = smem_desc(smem_u32(sm.own[qb][0]) + hb, 16, 1024, LAYOUT_SW128);
              const uint64_t a1 = smem_desc(smem_u32(sm.own[qb][1]) + hb, 16, 1024, LAYOUT_SW128);
              // B = rows [64 half, 64 half + 64) of the edge tiles (8 swizzle atoms of 8 rows = 8192 bytes)
              const uint64_t b0 = smem_desc(smem_u32(sm.edge[st][0]) + half * 8192 + hb, 16, 1024, LAYOUT_SW128);
              const uint64_t b1 = smem_desc(smem_u32(sm.edge[st][1]) + half * 8192 + hb, 16, 1024, LAYOUT_SW128);
#pragma unroll
              for (int ks = 0; ks < HD / 16; ++ks)
                mma_ss_w(tmem + set * 128, desc_advance(a0, ks * 32), desc_advance(b0, ks * 32), idesc_xy, ks > 0);
#pragma unroll
              for (int ks = 0; ks < HD / 16; ++ks)
                mma_ss_w(tmem + set * 128 + 64, desc_advance(a1, ks * 32), desc_advance(b1, ks * 32), idesc_xy, ks > 0);
              if (MODE == MODE_DKV) {
                // one more K step each: X -= 1 lse2^T, Y -= 1 delta^T (selector tile x the edge's statistics tile, head slice h)
                const uint64_t ac = smem_desc(smem_u32(sm.cst), 16, 1024, LAYOUT_SW128);
                const uint64_t bs = smem_desc(smem_u32(sm.stile[sb]) + half * 8192 + h * 32, 16, 1024, LAYOUT_SW128);
                mma_ss_w(tmem + set * 128, ac, bs, idesc_xy, 1u);
                mma_ss_w(tmem + set * 128 + 64, desc_advance(ac, 32), bs, idesc_xy, 1u);
              }
              mma_commit_w(&sm.xy_full[set]);
            }
          }
          mma_commit_w(&sm.edge_empty[st]);
          if (MODE == MODE_DKV) mma_commit_w(&sm.st_empty[sb]);
          if (e + 1 == ns.e_end) mma_commit_w(&sm.own_empty[qb]);
        }
      }
    }
  } else if (warp == W_TX || warp == W_TY) {
    // ------------------------------------------------------------------ consumer MMAs
    //   which = 0: X-side operand (MODE_DQ: P, MODE_DKV: P^T), which = 1: Y-side operand (W / dS^T)
    //   B = an edge tile as MN-major operand: K / K (MODE_DQ), dO / Q' (MODE_DKV)
    //   destination: MODE_DKV block `which` of the node accumulators; MODE_DQ which = 1 the dQ accumulator,
    //   which = 0 the P K ring slot of the item (fresh per item, folded by the elementwise warps).
    if (MODE == MODE_DQ) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    else asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
    {
      const uint32_t which = warp == W_TY ? 1u : 0u;
      const uint32_t idesc_t = idesc_bf16(128, 16, 0, 1);      // N = 16 per MMA (two per K step when hd = 32)
      const int btile = (which == 0 && MODE == MODE_DKV) ? 1 : 0;
      const bool ring = MODE == MODE_DQ && which == 0;
      uint32_t wtx[4] = {0, 0, 0, 0};
      const uint32_t tx_begin = PROF ? (uint32_t)clock() : 0u;
      constexpr int R = 64 / HD;                                // ring slots
      uint32_t qi = 0, edge = 0, k = 0, item = 0;
      for (;; ++qi) {
        const uint32_t qb = qi & 1;
        AMP_WAIT(&sm.own_full[qb], (qi >> 1) & 1, 211);
        const NodeSlot ns = uniform_slot(sm.slot[qb]);
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.own_empty[qb]);
        if (ns.node < 0) break;
        // the accumulators of the previous node must have been read back
        AMP_XWAIT(0, SM_OFF(acc_empty), (qi & 1) ^ 1, 212);
        for (int e = ns.e_begin; e < ns.e_end; ++e, ++edge) {
          const uint32_t st = edge % NS;
          const uint32_t keep = ring ? 0u : (e != ns.e_begin ? 1u : 0u);
#pragma unroll
          for (int h = 0; h < H; ++h, ++item) {
            const uint64_t bd = smem_desc(smem_u32(sm.edge[st][btile]) + h * (HD * 2), 16, 1024, LAYOUT_SW128);
            const uint32_t slot = item % R;
            const uint32_t d_col = ring ? tmem + kAccCol + 64 + slot * HD
                                        : tmem + kAccCol + (MODE == MODE_DKV ? which * 64 : 0) + h * HD;
            if (ring) AMP_XWAIT(1, SM_OFF(ko_empty) + 8 * slot, ((item / R) & 1) ^ 1, 214);
            for (int half = 0; half < nhalf; ++half, ++k) {
              const uint32_t set = k % kSets;
              AMP_XWAIT(2, SM_OFF(op_full) + 8 * set, (k / kSets) & 1, 213);
              tc_fence_after();
              const uint32_t a_col = tmem + set * 128 + which * 64;
              // all four K steps, always: score columns >= F carry zero operands and meet zero-filled tile rows
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                const uint32_t acc = (half | ks) ? 1u : keep;
#pragma unroll
                for (int part = 0; part < HD / 16; ++part)
                  mma_ts_w(d_col + 16 * part, a_col + 16 * ks, desc_advance(bd, (4 * half + ks) * 2048 + part * 32), idesc_t, acc);
              }
              mma_commit_w(&sm.set_empty[set]);
            }
            if (ring) mma_commit_w(&sm.ko_full[slot]);
          }
          mma_commit_w(&sm.edge_empty[st]);
        }
        mma_commit_w(&sm.acc_full);
      }
      if (PROF && blockIdx.x == 0 && lane == 0) {
        long long* pr = prof + 32 + 8 * which;      // [0] accumulators free, [1] ring slot free, [2] operands, [6] total, [7] items
        for (int i = 0; i < 3; ++i) pr[i] = wtx[i];
        pr[6] = (uint32_t)clock() - tx_begin;
        pr[7] = item;
      }
    }
  } else if (MODE == MODE_DQ && warp >= kEwWarps + 4) {
    // ------------------------------------------------------------------ fold warps (MODE_DQ): one per TMEM lane quarter
    //   per item: delta = sum of the four elementwise partials; racc -= delta o (P K) read from the ring slot; delta is
    //   written for the dK/dV pass.  Per node: dQ = (TMEM accumulator + racc) * scale.  racc lives in shared memory
    //   ([column][row], conflict-free) so that these warps run on 56 registers.
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    {
      const uint32_t q4 = warp & 3;
      const int row = q4 * 32 + lane;
      const bool row_ok = row < F;
      const uint32_t lane_base = tmem + ((uint32_t)(q4 * 32) << 16);
      constexpr int R = 64 / HD;
      const uint32_t racc = SM_OFF(racc) + 4 * row;   // column c at racc + 512 c (shared-space addresses: LDS / STS)
#pragma unroll 8
      for (int c = 0; c < 64; ++c) sts_f32(racc + 512 * c, 0.f);
      uint32_t qi = 0, item = 0;
      uint32_t wtx[4] = {0, 0, 0, 0};
      const uint32_t tf_begin = PROF ? (uint32_t)clock() : 0u;
      for (;; ++qi) {
        const uint32_t qb = qi & 1;
        AMP_XWAIT(3, SM_OFF(own_full) + 8 * qb, (qi >> 1) & 1, 401);
        const NodeSlot ns = uniform_slot(sm.slot[qb]);
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.own_empty[qb]);
        if (ns.node < 0) break;
        // statistics rows of an edge's items are consecutive; with two head groups an edge's block has HT rows, this item's
        // group owns H of them
        float* delta_row = delta + ((int64_t)ns.e_begin * HT + ns.grp * H) * Fs + row;
        for (int e = ns.e_begin; e < ns.e_end; ++e) {
#pragma unroll 1
          for (int h = 0; h < H; ++h, ++item) {
            const uint32_t slot = item % R;
            AMP_XWAIT(0, SM_OFF(dl_bar) + 32 * (item & 7) + 8 * q4, (item >> 3) & 1, 402);
            const uint32_t dls = SM_OFF(dl) + 2048 * (item & 7) + 4 * row;
            const float dsum = (lds_f32(dls) + lds_f32(dls + 512)) + (lds_f32(dls + 1024) + lds_f32(dls + 1536));
            AMP_XWAIT(1, SM_OFF(ko_full) + 8 * slot, (item / R) & 1, 403);
            tc_fence_after();
            const uint32_t ra = racc + h * HD * 512;
#pragma unroll
            for (int part = 0; part < HD / 16; ++part) {
              uint32_t ko[16];
              tmem_ld_32x32b_x16(lane_base + kAccCol + 64 + slot * HD + 16 * part, ko);
              tmem_ld_wait();
#pragma unroll
              for (int x = 0; x < 16; ++x) {
                const uint32_t a = ra + (16 * part + x) * 512;
                sts_f32(a, fmaf(-dsum, __uint_as_float(ko[x]), lds_f32(a)));
              }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_a(SM_OFF(ko_empty) + 8 * slot);
            if (row_ok) delta_row[0] = dsum;
            delta_row += Fs;
          }
          if (GROUPS > 1) delta_row += (HT - H) * Fs;
        }
        // node end: all consumer MMAs of the node have landed in the dQ accumulator
        AMP_XWAIT(2, SM_OFF(acc_full), qi & 1, 404);
        tc_fence_after();
        float* o = d_qkv + ((int64_t)ns.node * F + row) * out_ld + out_c0;
#pragma unroll 1
        for (int c4 = 0; c4 < 4; ++c4) {
          uint32_t a[16];
          tmem_ld_32x32b_x16(lane_base + kAccCol + 16 * c4, a);
          tmem_ld_wait();
          float v[16];
#pragma unroll
          for (int x = 0; x < 16; ++x) {
            v[x] = (__uint_as_float(a[x]) + lds_f32(racc + (16 * c4 + x) * 512)) * out_scale0;
            sts_f32(racc + (16 * c4 + x) * 512, 0.f);
          }
          if (row_ok && out_bf16) {
            // gradient rows as bf16 (single-GPU path: the dX projection and the weight-gradient kernel feed bf16 operands)
            uint16_t* ob = reinterpret_cast<uint16_t*>(d_qkv) + ((int64_t)ns.node * F + row) * out_ld + out_c0 +
                           (GROUPS == 1 ? 16 * c4 : 32 * ns.grp + 8 * c4);
#pragma unroll
            for (int x = 0; x < (GROUPS == 1 ? 16 : 8); x += 8)
              *reinterpret_cast<uint4*>(ob + x) = make_uint4(pack_bf16x2(v[x], v[x + 1]), pack_bf16x2(v[x + 2], v[x + 3]),
                                                             pack_bf16x2(v[x + 4], v[x + 5]), pack_bf16x2(v[x + 6], v[x + 7]));
          } else if (row_ok) {
            // GROUPS == 2 (HD == 16): TMEM columns [16 c4, 16 c4 + 8) are the real columns of padded head c4
            float* oc = GROUPS == 1 ? o + 16 * c4 : o + 32 * ns.grp + 8 * c4;
#pragma unroll
            for (int x = 0; x < (GROUPS == 1 ? 16 : 8); x += 4) {
              float4 r = make_float4(v[x], v[x + 1], v[x + 2], v[x + 3]);
              if (accumulate) {   // ring phases (multi-GPU): dQ is a sum over all phases' edges
                const float4 pr = *reinterpret_cast<const float4*>(oc + x);
                r.x += pr.x; r.y += pr.y; r.z += pr.z; r.w += pr.w;
              }
              *reinterpret_cast<float4*>(oc + x) = r;
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.acc_empty);
      }
      if (PROF && blockIdx.x == 0 && warp == kEwWarps + 4 && lane == 0) {
        long long* pr = prof + 48;                  // [0] delta partials, [1] P K slot, [2] accumulators, [3] own tiles, [6] total, [7] items
        for (int i = 0; i < 4; ++i) pr[i] = wtx[i];
        pr[6] = (uint32_t)clock() - tf_begin;
        pr[7] = item;
      }
    }
  } else if (MODE == MODE_DKV && warp >= kEwWarps + 4) {
    // ------------------------------------------------------------------ statistics warps (MODE_DKV): thread = destination token
    //   per edge: the bulk-copied fp32 rows lse2[h][i], delta[h][i] become the bf16 tile S[i][16h .. 16h+3] =
    //   (-lse_hi, -lse_lo, -delta_hi, -delta_lo); rows >= F and the odd 16-byte chunks stay zero.
    // register pool of the CTA: 768 threads x 80 at launch; the elementwise warps take 16 x 32 x (96 - 80) = 8192, which the
    // control warps (80 -> 48) and these warps (80 -> 48) release: 2 x 128 x 32 = 8192 (setmaxnreg.inc blocks until then)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
    {
      const int row = (warp & 3) * 32 + lane;
      const bool row_ok = row < F;
      uint32_t qi = 0, edge = 0;
      for (;; ++qi) {
        const uint32_t qb = qi & 1;
        AMP_WAIT_A(SM_OFF(own_full) + 8 * qb, (qi >> 1) & 1, 501);
        const NodeSlot ns = uniform_slot(sm.slot[qb]);
        __syncwarp();
        if (lane == 0) mbar_arrive_a(SM_OFF(own_empty) + 8 * qb);
        if (ns.node < 0) break;
        for (int e = ns.e_begin; e < ns.e_end; ++e, ++edge) {
          const uint32_t st = edge % NS, sb = edge & 1;
          AMP_WAIT_A(SM_OFF(edge_full) + 8 * st, (edge / NS) & 1, 502);
          AMP_WAIT_A(SM_OFF(st_empty) + 8 * sb, ((edge >> 1) & 1) ^ 1, 503);
          if (row_ok) {
            const uint32_t src = SM_OFF(stat) + (uint32_t)(st * 2 * kStatFloats * 4) + 4 * row;
            const uint32_t dst = SM_OFF(stile) + sb * kTileBytes + row * 128;
#pragma unroll
            for (int h = 0; h < H; ++h) {
              const float l = -lds_f32(src + h * Fs * 4);
              const float dd = -lds_f32(src + kStatFloats * 4 + h * Fs * 4);
              // x = hi + lo with hi = bf16(x), lo = bf16(x - hi)
              const uint32_t lh = pack_bf16x2(l, 0.f) & 0xffffu, dh = pack_bf16x2(dd, 0.f) & 0xffffu;
              const float l_lo = l - __uint_as_float(lh << 16), d_lo = dd - __uint_as_float(dh << 16);
              const uint32_t w0 = lh | (pack_bf16x2(0.f, l_lo) & 0xffff0000u);
              const uint32_t w1 = dh | (pack_bf16x2(0.f, d_lo) & 0xffff0000u);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (((2 * h) ^ (row & 7)) << 4)), "r"(w0), "r"(w1),
                           "r"(0u), "r"(0u)
                           : "memory");
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive_a(SM_OFF(st_full) + 8 * sb);
            mbar_arrive_a(SM_OFF(edge_empty) + 8 * st);
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ elementwise warps
    asm volatile("setmaxnreg.inc.sync.aligned.u32 96;");
    const uint32_t q4 = warp & 3;                   // TMEM lane quarter
    const uint32_t grp = warp >> 3;                 // group: works on half-items with k & 1 == grp
    const uint32_t cb = (warp >> 2) & 1;            // column block: score columns [32cb, 32cb+32) of the half
    const uint32_t j4 = grp * 2 + cb;               // 0..3 within the lane quarter (delta partial slot / epilogue columns)
    const int row = q4 * 32 + lane;
    const bool row_ok = row < F;
    const uint32_t lane_base = tmem + ((uint32_t)(q4 * 32) << 16);
    uint32_t qi = 0, k = 0, ei = 0, item = 0;
    // light-weight wait accounting (debug entry point only): cycles this warp spent blocked on each kind of barrier
    const bool do_prof = PROF && blockIdx.x == 0 && (warp == 0 || warp == 8);
    uint32_t wt[6] = {0, 0, 0, 0, 0, 0};           // [4], [5] unused since the fold moved to its own warps
    const uint32_t t_begin = do_prof ? (uint32_t)clock() : 0u;
#define AMP_PHASE(i)
#define AMP_TWAIT(i, bar, parity, code)                                       \
  do {                                                                        \
    if (PROF) {                                                               \
      const uint32_t t0_ = (uint32_t)clock();                                 \
      AMP_WAIT_A(bar, parity, code);                                          \
      wt[i] += (uint32_t)clock() - t0_;                                       \
    } else {                                                                  \
      AMP_WAIT_A(bar, parity, code);                                          \
    }                                                                         \
  } while (0)
    for (;;) {
      const uint32_t qb = qi & 1;
      AMP_TWAIT(0, SM_OFF(own_full) + 8 * qb, (qi >> 1) & 1, 301);
      const NodeSlot ns = sm.slot[qb];
      __syncwarp();
      if (lane == 0) mbar_arrive_a(SM_OFF(own_empty) + 8 * qb);
      if (ns.node < 0) break;
      for (int e = ns.e_begin; e < ns.e_end; ++e, ++ei) {
        const uint32_t st = ei % NS;
        if (MODE == MODE_DQ) AMP_TWAIT(1, SM_OFF(edge_full) + 8 * st, (ei / NS) & 1, 306);   // acquire the bulk-copied statistics rows
        // NOT unrolled: one copy of the body keeps the elementwise loop inside the instruction cache
#pragma unroll 1
        for (int h = 0; h < H; ++h, ++item) {
          const uint32_t ls_addr = SM_OFF(stat) + (uint32_t)((st * Smem::NSTAT * kStatFloats + h * Fs) * 4);
          // MODE_DQ: this thread's row statistic (rows >= F read stale shared memory: they only reach discarded output rows)
          const float L = MODE == MODE_DQ ? lds_f32(ls_addr + 4 * row) : 0.f;
          float2 dl2 = make_float2(0.f, 0.f);
          AMP_PHASE(5);
          // this group's half-item of the item: with two halves per item the item starts at an even k and group g owns
          // half g; with one half per item the groups alternate items
          const int half = nhalf == 2 ? (int)grp : 0;
          const uint32_t kk = k + half;
          k += nhalf;
          if (nhalf == 2 || (kk & 1) == grp) {
            const uint32_t set = kk % kSets;
            AMP_TWAIT(2, SM_OFF(xy_full) + 8 * set, (kk / kSets) & 1, 302);
            tc_fence_after();
            const uint32_t xbase = lane_base + set * 128 + 32 * cb;
            uint32_t xs[2][16], ys[2][16];
            tmem_ld_32x32b_x16(xbase, xs[0]);
            tmem_ld_32x32b_x16(xbase + 64, ys[0]);
            tmem_ld_32x32b_x16(xbase + 16, xs[1]);
            tmem_ld_32x32b_x16(xbase + 64 + 16, ys[1]);
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
              const int col0 = 64 * half + 32 * cb + 16 * ch;     // first score column of this chunk
              uint32_t px[8], py[8];
              if (ch == 0) tmem_ld_wait();
              // one pair of score columns; TAIL adds the selects that keep columns >= F (zero-filled tile rows, undefined
              // statistics) out of delta and out of the MMA operands
              auto pair = [&](int j, auto tail_tag) {
                constexpr bool TAIL = decltype(tail_tag)::value;
                const int c0 = col0 + 2 * j;
                const float2 x2 = make_float2(__uint_as_float(xs[ch][2 * j]), __uint_as_float(xs[ch][2 * j + 1]));
                const float2 y2 = make_float2(__uint_as_float(ys[ch][2 * j]), __uint_as_float(ys[ch][2 * j + 1]));
                float2 p2;
                const bool poly = (MODE == MODE_DQ ? kPolyShareDq : kPolyShareDkv) && (j & 3) == 3;
                if (MODE == MODE_DQ) {
                  const float2 e2 = f2add(x2, make_float2(-L, -L));
                  p2 = poly ? ex2_poly2(e2) : make_float2(ex2_approx(e2.x), ex2_approx(e2.y));
                } else {
                  // the column statistics were subtracted by the score MMAs (statistics tile)
                  p2 = poly ? ex2_poly2(x2) : make_float2(ex2_approx(x2.x), ex2_approx(x2.y));
                }
                float2 u2 = f2mul(p2, y2);
                if (TAIL) {
                  if (c0 >= F) { p2.x = 0.f; u2.x = 0.f; }
                  if (c0 + 1 >= F) { p2.y = 0.f; u2.y = 0.f; }
                }
                if (MODE == MODE_DQ) dl2 = f2add(dl2, u2);
                px[j] = pack_bf16x2(p2.x, p2.y);
                py[j] = pack_bf16x2(u2.x, u2.y);
              };
              if (col0 + 16 > F) {     // warp-uniform: only the chunk that straddles F and the ones behind it
#pragma unroll
                for (int j = 0; j < 8; ++j) pair(j, TrueTag{});
              } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) pair(j, FalseTag{});
              }
              // packed operands go to the first 8 of the chunk's own 16 columns: always behind this thread's own reads
              tmem_st_32x32b_x8(xbase + 16 * ch, px);
              tmem_st_32x32b_x8(xbase + 64 + 16 * ch, py);
            }
            AMP_PHASE(1);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_a(SM_OFF(op_full) + 8 * set);
            AMP_PHASE(2);
          }
          if (MODE == MODE_DQ) {
            // partial delta of this warp's columns (zero when the item had no half-item for this group)
            sts_f32(SM_OFF(dl) + 2048 * (item & 7) + 512 * j4 + 4 * row, dl2.x + dl2.y);
            __syncwarp();
            if (lane == 0) mbar_arrive_a(SM_OFF(dl_bar) + 32 * (item & 7) + 8 * q4);
            AMP_PHASE(3);
          }
        }
        if (MODE == MODE_DQ) {
          __syncwarp();
          if (lane == 0) mbar_arrive_a(SM_OFF(edge_empty) + 8 * st);   // this warp no longer reads the stage's statistics rows
        }
      }
      // node epilogue.  MODE_DQ: the fold warps own the accumulators (they add the delta term and write dQ); the
      // elementwise warps go straight on to the next node.
      if (MODE == MODE_DQ) {
      } else {
        // all consumer MMAs of the node have landed in the accumulators
        AMP_TWAIT(3, SM_OFF(acc_full), qi & 1, 304);
        tc_fence_after();
        // block 0 = dV (X-side operand P^T), block 1 = dK (Y-side operand dS^T); this warp owns 32 of the 128 columns
        uint32_t a[32];
        tmem_ld_32x32b_x32(lane_base + kAccCol + 32 * j4, a);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_a(SM_OFF(acc_empty));
        if (row_ok) {
          const bool is_dv = j4 < 2;
          const float sc = is_dv ? out_scale1 : out_scale0;
          if (GROUPS == 1) {
            const bool to_halo = halo_bf16 != nullptr && ns.node >= halo_from;
            if (to_halo || out_bf16) {
              // halo source (multi-GPU): this partial row travels to its owner, written as bf16 [node - halo_from][row][out_ld];
              // out_bf16 (single GPU): every row is written as bf16 for the bf16-fed projection / weight-gradient kernels
              uint16_t* ob = to_halo ? halo_bf16 + ((int64_t)(ns.node - halo_from) * F + row) * out_ld
                                     : reinterpret_cast<uint16_t*>(d_qkv) + ((int64_t)ns.node * F + row) * out_ld;
              uint4* o = reinterpret_cast<uint4*>(ob + (is_dv ? out_c1 : out_c0) + 32 * (j4 & 1));
#pragma unroll
              for (int x = 0; x < 32; x += 8)
                o[x >> 3] = make_uint4(pack_bf16x2(__uint_as_float(a[x]) * sc, __uint_as_float(a[x + 1]) * sc),
                                       pack_bf16x2(__uint_as_float(a[x + 2]) * sc, __uint_as_float(a[x + 3]) * sc),
                                       pack_bf16x2(__uint_as_float(a[x + 4]) * sc, __uint_as_float(a[x + 5]) * sc),
                                       pack_bf16x2(__uint_as_float(a[x + 6]) * sc, __uint_as_float(a[x + 7]) * sc));
            } else {
              float4* o = reinterpret_cast<float4*>(d_qkv + ((int64_t)ns.node * F + row) * out_ld +
                                                    (is_dv ? out_c1 : out_c0) + 32 * (j4 & 1));
#pragma unroll
              for (int x = 0; x < 32; x += 4)
                o[x >> 2] = make_float4(__uint_as_float(a[x]) * sc, __uint_as_float(a[x + 1]) * sc,
                                        __uint_as_float(a[x + 2]) * sc, __uint_as_float(a[x + 3]) * sc);
            }
          } else {
            // head_dim 8: this warp's 32 accumulator columns are padded heads 2 (j4 & 1) and 2 (j4 & 1) + 1 (16 columns each,
            // 8 real); head lh of group grp lives in columns [32 grp + 8 lh, + 8) of the 64-wide dK (dV) row
            const int ns_grp = ns.grp;   // (the slot itself may already hold the next node: this warp released it at the top)
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int col = (is_dv ? out_c1 : out_c0) + 32 * ns_grp + 8 * (2 * (int)(j4 & 1) + u);
              const uint32_t* au = a + 16 * u;
              const bool to_halo = halo_bf16 != nullptr && ns.node >= halo_from;
              if (to_halo || out_bf16) {
                uint16_t* ob = to_halo ? halo_bf16 + ((int64_t)(ns.node - halo_from) * F + row) * out_ld
                                       : reinterpret_cast<uint16_t*>(d_qkv) + ((int64_t)ns.node * F + row) * out_ld;
                *reinterpret_cast<uint4*>(ob + col) =
                    make_uint4(pack_bf16x2(__uint_as_float(au[0]) * sc, __uint_as_float(au[1]) * sc),
                               pack_bf16x2(__uint_as_float(au[2]) * sc, __uint_as_float(au[3]) * sc),
                               pack_bf16x2(__uint_as_float(au[4]) * sc, __uint_as_float(au[5]) * sc),
                               pack_bf16x2(__uint_as_float(au[6]) * sc, __uint_as_float(au[7]) * sc));
              } else {
                float4* o = reinterpret_cast<float4*>(d_qkv + ((int64_t)ns.node * F + row) * out_ld + col);
                o[0] = make_float4(__uint_as_float(au[0]) * sc, __uint_as_float(au[1]) * sc, __uint_as_float(au[2]) * sc,
                                   __uint_as_float(au[3]) * sc);
                o[1] = make_float4(__uint_as_float(au[4]) * sc, __uint_as_float(au[5]) * sc, __uint_as_float(au[6]) * sc,
                                   __uint_as_float(au[7]) * sc);
              }
            }
          }
        }
      }
      AMP_PHASE(7);
      ++qi;
    }
    if (do_prof && lane == 0) {
      long long* pr = prof + (warp == 0 ? 0 : 16);
      for (int i = 0; i < 6; ++i) pr[i] = wt[i];
      pr[6] = (uint32_t)clock() - t_begin;
      pr[8] = item;
    }
#undef AMP_PHASE
#undef AMP_TWAIT
  }
fail:
  if (GROUPS == 2) cp_async_wait<0>();   // no copy into this CTA's shared memory may outlive it
  tc_fence_before();
  __syncthreads();
  if (warp == W_SCORE) tmem_dealloc(tmem, 512);
}

long long* g_bwd_prof = nullptr;   // debug: set through ampconv_debug_set_bwd_profile

template <int HD, int MODE, int GROUPS>
int launch_bwd(const CUtensorMap& own0, const CUtensorMap& own1, const CUtensorMap& oth0, const CUtensorMap& oth1,
               const void* gown0, const void* gown1, const void* goth0, const void* goth1,
               const int32_t* rowptr, const int32_t* nbr, const int32_t* slot_of, const int32_t* lse_map, const int32_t* order,
               const float* lse2, float* delta, float* d_qkv, int* counter, int* status, int N, int F, float s0, float s1, int out_ld, int out_c0,
               int out_c1, uint16_t* halo_bf16, int halo_from, int accumulate, int out_bf16, cudaStream_t stream) {
  const size_t smem = sizeof(BwdSmem<MODE>) + 1024;
  const int items = N * GROUPS;
  const int grid = items < sm_count() ? items : sm_count();
  long long* prof = g_bwd_prof;
#define AMP_LAUNCH_BWD(NH_, PROF_, prof_)                                                                                          \
  do {                                                                                                                           \
    AMPCONV_CUDA_TRY(cudaFuncSetAttribute(attn_bwd_bf16_kernel<HD, MODE, NH_, GROUPS, PROF_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                          (int)smem));                                                                           \
    attn_bwd_bf16_kernel<HD, MODE, NH_, GROUPS, PROF_><<<grid, threads_of(MODE), smem, stream>>>(                                         \
        own0, own1, oth0, oth1, rowptr, nbr, slot_of, lse_map, order, lse2, delta, d_qkv, counter, status, N, F, s0, s1, out_ld, out_c0,   \
        out_c1, halo_bf16, halo_from, accumulate, out_bf16, reinterpret_cast<const uint8_t*>(gown0),                                     \
        reinterpret_cast<const uint8_t*>(gown1), reinterpret_cast<const uint8_t*>(goth0),                                       \
        reinterpret_cast<const uint8_t*>(goth1), prof_);                                                                                  \
  } while (0)
  if (prof && GROUPS == 1) {
    if (F > 64) AMP_LAUNCH_BWD(2, true, prof);
    else AMP_LAUNCH_BWD(1, true, prof);
  } else {
    if (F > 64) AMP_LAUNCH_BWD(2, false, nullptr);
    else AMP_LAUNCH_BWD(1, false, nullptr);
  }
#undef AMP_LAUNCH_BWD
  AMPCONV_CHECK_LAUNCH();
  return AMPCONV_OK;
}

}  // namespace
}  // namespace ampconv

using namespace ampconv;

extern "C" int ampconv_attn_bf16_supported(int F, int d, int H);

// N_own: nodes the pass's tensor maps cover (destinations for MODE_DQ, sources for MODE_DKV); N_oth: nodes of the edge tiles.
// n_work: length of the `order` work list (-1: all N_own nodes); accumulate: MODE_DQ only, see the *_phase entry points.
static int bwd_common(int mode, const void* q, const void* k, const void* v, const void* d_agg_bf16,
                      const int32_t* rowptr, const int32_t* nbr, const int32_t* slot_of, const int32_t* order,
                      const float* lse2, float* delta, float* out, int out_ld, int out_c0, int out_c1, int64_t N_dst, int64_t N_kv, int64_t E, int F, int d,
                      int H, void* workspace, size_t workspace_bytes, void* stream_, void* halo_bf16 = nullptr,
                      int64_t halo_from = 0, int64_t n_work = -1, int accumulate = 0, const int32_t* lse_map = nullptr,
                      int out_bf16 = 0) {
  AMPCONV_REQUIRE(N_dst >= 0 && N_kv >= 0 && E >= 0 && F > 0 && d > 0 && H > 0 && d % H == 0);
  if (!ampconv_attn_bf16_supported(F, d, H)) return AMPCONV_ERR_UNSUPPORTED;
  const int64_t N_own = mode == MODE_DQ ? N_dst : N_kv;
  if (n_work < 0) n_work = N_own;
  if (N_own == 0 || n_work == 0) return AMPCONV_OK;
  AMPCONV_REQUIRE(n_work <= N_own && (n_work == N_own || order != nullptr));
  AMPCONV_REQUIRE(q && k && v && d_agg_bf16 && rowptr && workspace && (out || halo_bf16));
  AMPCONV_REQUIRE(E == 0 || (nbr && lse2 && delta));
  if (workspace_bytes < 256) return AMPCONV_ERR_WORKSPACE;
  cudaStream_t stream = as_stream(stream_);
  CUtensorMap mq, mk, mv, mg;
  const int64_t nd = N_dst > 0 ? N_dst : 1, nk = N_kv > 0 ? N_kv : 1;
  if (!make_tensor_map_bf16_3d(&mq, q, kD, F, nd, kD, 128) || !make_tensor_map_bf16_3d(&mk, k, kD, F, nk, kD, 128) ||
      !make_tensor_map_bf16_3d(&mv, v, kD, F, nk, kD, 128) || !make_tensor_map_bf16_3d(&mg, d_agg_bf16, kD, F, nd, kD, 128))
    return AMPCONV_ERR_CUDA;   // (head_dim 8 loads its tiles with cp.async from the raw pointers instead)
  int* counter = reinterpret_cast<int*>(workspace) + (mode == MODE_DQ ? 2 : 4);
  int* status = reinterpret_cast<int*>(workspace) + 1;
  AMPCONV_CUDA_TRY(cudaMemsetAsync(counter, 0, sizeof(int), stream));
  const int hd = d / H;
  const float inv_sqrt_hd = 1.0f / sqrtf((float)hd);
  const float ln2 = 0.6931471805599453f;
  if (mode == MODE_DQ) {
    // dQ = hd^-1/2 * (dS K)
    if (hd == 8)
      return launch_bwd<16, MODE_DQ, 2>(mq, mg, mk, mv, q, d_agg_bf16, k, v, rowptr, nbr, slot_of, lse_map, order, lse2, delta, out, counter, status, (int)n_work, F,
                                       inv_sqrt_hd, 0.f, out_ld, out_c0, out_c1, nullptr, 0, accumulate, out_bf16, stream);
    if (hd == 16)
      return launch_bwd<16, MODE_DQ, 1>(mq, mg, mk, mv, q, d_agg_bf16, k, v, rowptr, nbr, slot_of, lse_map, order, lse2, delta, out, counter, status, (int)n_work, F,
                                    inv_sqrt_hd, 0.f, out_ld, out_c0, out_c1, nullptr, 0, accumulate, out_bf16, stream);
    return launch_bwd<32, MODE_DQ, 1>(mq, mg, mk, mv, q, d_agg_bf16, k, v, rowptr, nbr, slot_of, lse_map, order, lse2, delta, out, counter, status, (int)n_work, F,
                                  inv_sqrt_hd, 0.f, out_ld, out_c0, out_c1, nullptr, 0, accumulate, out_bf16, stream);
  }
  // dK = hd^-1/2 * dS^T Q = ln2 * dS^T Q'  (Q' = Q * log2e / sqrt(hd)),  dV = P^T dO
  if (hd == 8)
    return launch_bwd<16, MODE_DKV, 2>(mk, mv, mq, mg, k, v, q, d_agg_bf16, rowptr, nbr, slot_of, lse_map, order, lse2, delta, out, counter, status, (int)n_work, F,
                                      ln2, 1.f, out_ld, out_c0, out_c1, reinterpret_cast<uint16_t*>(halo_bf16), (int)halo_from, 0, out_bf16, stream);
  if (hd == 16)
    return launch_bwd<16, MODE_DKV, 1>(mk, mv, mq, mg, k, v, q, d_agg_bf16, rowptr, nbr, slot_of, lse_map, order, lse2, delta, out, counter, status, (int)n_work, F,
                                   ln2, 1.f, out_ld, out_c0, out_c1, reinterpret_cast<uint16_t*>(halo_bf16), (int)halo_from, 0, out_bf16, stream);
  return launch_bwd<32, MODE_DKV, 1>(mk, mv, mq, mg, k, v, q, d_agg_bf16, rowptr, nbr, slot_of, lse_map, order, lse2, delta, out, counter, status, (int)n_work, F, ln2,
                                 1.f, out_ld, out_c0, out_c1, reinterpret_cast<uint16_t*>(halo_bf16), (int)halo_from, 0, out_bf16, stream);
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

// Same two passes with the gradient rows written as bf16 (d_qkv_bf16 [rows, 192]): the consumers -- the dX projection and the
// in_proj weight-gradient kernel -- feed bf16 operands to the tensor cores anyway, so nothing is lost and 25 GB of HBM traffic
// per step at the ogbn-arxiv shape are saved (fp32 rows: 16.6 GB written once and read twice).
extern "C" int ampconv_attn_bwd_dq_bf16_h(const void* q, const void* k, const void* v, const void* d_agg_bf16,
                                          const float* lse2, const int32_t* dst_rowptr, const int32_t* dst_src,
                                          const int32_t* order, void* d_qkv_bf16, float* delta, int64_t N, int64_t E, int F,
                                          int d, int H, void* workspace, size_t workspace_bytes, void* stream) {
  return bwd_common(MODE_DQ, q, k, v, d_agg_bf16, dst_rowptr, dst_src, nullptr, order, lse2, delta,
                    reinterpret_cast<float*>(d_qkv_bf16), 3 * kD, 0, 0, N, N, E, F, d, H, workspace, workspace_bytes, stream, nullptr, 0,
                    -1, 0, nullptr, 1);
}

extern "C" int ampconv_attn_bwd_dkv_bf16_h(const void* q, const void* k, const void* v, const void* d_agg_bf16,
                                           const float* lse2, const float* delta, const int32_t* src_rowptr,
                                           const int32_t* src_dst, const int32_t* src_pos, const int32_t* order,
                                           void* d_qkv_bf16, int64_t N, int64_t E, int F, int d, int H,
                                           void* workspace, size_t workspace_bytes, void* stream) {
  return bwd_common(MODE_DKV, q, k, v, d_agg_bf16, src_rowptr, src_dst, src_pos, order, lse2, const_cast<float*>(delta),
                    reinterpret_cast<float*>(d_qkv_bf16), 3 * kD, kD, 2 * kD, N, N, E, F, d, H, workspace, workspace_bytes, stream,
                    nullptr, 0, -1, 0, nullptr, 1);
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

// As ampconv_attn_bwd_dkv_bf16_part, for the halo exchange: sources [0, num_own) are this rank's nodes, their rows go to
// d_kv_own fp32 [num_own*F, 128]; sources [num_own, num_kv_nodes) are halo nodes, their partial rows are written as bf16
// into d_kv_halo [(num_kv_nodes - num_own)*F, 128], ready to be sent to their owners.
extern "C" int ampconv_attn_bwd_dkv_bf16_halo(const void* q, const void* k, const void* v, const void* d_agg_bf16,
                                              const float* lse2, const float* delta, const int32_t* src_rowptr,
                                              const int32_t* src_dst, const int32_t* src_pos, const int32_t* order,
                                              float* d_kv_own, void* d_kv_halo, int64_t num_nodes, int64_t num_own,
                                              int64_t num_kv_nodes, int64_t E, int F, int d, int H, void* workspace,
                                              size_t workspace_bytes, void* stream) {
  AMPCONV_REQUIRE(num_own >= 0 && num_own <= num_kv_nodes && (num_own == num_kv_nodes || d_kv_halo != nullptr));
  return bwd_common(MODE_DKV, q, k, v, d_agg_bf16, src_rowptr, src_dst, src_pos, order, lse2, const_cast<float*>(delta), d_kv_own,
                    2 * kD, 0, kD, num_nodes, num_kv_nodes, E, F, d, H, workspace, workspace_bytes, stream, d_kv_halo, num_own);
}

// Ring-phase variants (multi-GPU, ampnet_b200/distributed.py): one launch per source-owner phase.
//   _dq_phase : the `n_work` destinations listed in `order` (those with an edge in the phase); accumulate = 0 overwrites d_q
//               (and zero-fills destinations without an edge), accumulate = 1 adds this phase's contribution; delta is
//               indexed by the PHASE's destination-sorted slots and consumed by the same phase's _dkv launch.
//   _dkv_phase: the `n_work` sources listed in `order`, all inside the phase's compact-id range.  Own sources
//               (d_kv_halo NULL) go to d_kv_own fp32: row r at d_kv_own[r * own_ld], dK at column own_dk_col, dV at
//               own_dv_col (a dense [rows, 128] tensor or the dK | dV columns of d_qkv [rows, 192]); the halo sources of the
//               phase's owner, compact ids [halo_from, ...), go as bf16 rows to d_kv_halo[(id - halo_from)*F, 128]: the
//               block that travels to that owner.
extern "C" int ampconv_attn_bwd_dq_bf16_phase(const void* q, const void* k, const void* v, const void* d_agg_bf16,
                                              const float* lse2, const int32_t* lse_map, const int32_t* dst_rowptr,
                                              const int32_t* dst_src,
                                              const int32_t* order, int64_t n_work, int accumulate, float* d_q, int64_t d_q_ld,
                                              float* delta,
                                              int64_t num_nodes, int64_t num_kv_nodes, int64_t E, int F, int d, int H,
                                              void* workspace, size_t workspace_bytes, void* stream) {
  AMPCONV_REQUIRE(n_work >= 0 && (order != nullptr || n_work == 0) && d_q_ld >= kD && d_q_ld % 4 == 0);
  return bwd_common(MODE_DQ, q, k, v, d_agg_bf16, dst_rowptr, dst_src, nullptr, order, lse2, delta, d_q, (int)d_q_ld, 0, 0, num_nodes,
                    num_kv_nodes, E, F, d, H, workspace, workspace_bytes, stream, nullptr, 0, n_work, accumulate, lse_map);
}

extern "C" int ampconv_attn_bwd_dkv_bf16_phase(const void* q, const void* k, const void* v, const void* d_agg_bf16,
                                               const float* lse2, const int32_t* lse_map, const float* delta,
                                               const int32_t* src_rowptr,
                                               const int32_t* src_dst, const int32_t* src_pos, const int32_t* order,
                                               int64_t n_work, float* d_kv_own, int64_t own_ld, int64_t own_dk_col,
                                               int64_t own_dv_col, void* d_kv_halo, int64_t halo_from,
                                               int64_t num_nodes, int64_t num_kv_nodes, int64_t E, int F, int d, int H,
                                               void* workspace, size_t workspace_bytes, void* stream) {
  if (n_work == 0) return AMPCONV_OK;
  AMPCONV_REQUIRE(n_work > 0 && order != nullptr && halo_from >= 0 && ((d_kv_own != nullptr) != (d_kv_halo != nullptr)));
  if (d_kv_halo != nullptr)      // a halo phase: only the owner's bf16 block [*, 128] = dK | dV is written
    return bwd_common(MODE_DKV, q, k, v, d_agg_bf16, src_rowptr, src_dst, src_pos, order, lse2, const_cast<float*>(delta), nullptr,
                      2 * kD, 0, kD, num_nodes, num_kv_nodes, E, F, d, H, workspace, workspace_bytes, stream, d_kv_halo,
                      halo_from, n_work, 0, lse_map);
  AMPCONV_REQUIRE(own_ld >= 2 * kD && own_ld % 4 == 0 && own_dk_col % 4 == 0 && own_dv_col % 4 == 0 && own_dk_col + kD <= own_ld &&
                  own_dv_col + kD <= own_ld);
  return bwd_common(MODE_DKV, q, k, v, d_agg_bf16, src_rowptr, src_dst, src_pos, order, lse2, const_cast<float*>(delta), d_kv_own,
                    (int)own_ld, (int)own_dk_col, (int)own_dv_col, num_nodes, num_kv_nodes, E, F, d, H, workspace, workspace_bytes,
                    stream, nullptr, num_kv_nodes, n_work, 0, lse_map);
}

// Debug: when set to a device buffer of 64 int64, the next backward launches run the instrumented kernel and fill it
// with the cycles elementwise warp 0 of CTA 0 spent per phase (wait X/Y, compute, publish, delta exchange, second
// operand publish, statistics load, node / accumulator wait, node epilogue) and its item count.  NULL restores the
// product kernels.
extern "C" int ampconv_debug_set_bwd_profile(long long* prof) {
  ampconv::g_bwd_prof = prof;
  return AMPCONV_OK;
}
