// bf16 tensor-core family: fused per-edge multi-head attention + mean aggregation on tcgen05.
//
// Replaces the reference's per-edge chain (src/ampnet/conv/amp_conv.py:24-51 ->
// torch.nn.MultiheadAttention: head split, q*hd^-1/2, bmm, softmax, bmm --
// src/ampnet/conv/custom_multihead_attn_forward.py:4140-4186, 4376-4387) and PyG's scatter-mean
// (amp_conv.py:11) with ONE persistent, warp-specialised kernel per direction:
//
//   producer warp : dynamic node scheduler + TMA (cp.async.bulk.tensor, 128B swizzle) of the
//                   destination's Q tile (once per node) and each in-edge's K and V tiles
//                   ("gather by edge index" = a TMA box at row src*F of the node-major tensors);
//   3 MMA warps   : tcgen05.mma is issued by one elected lane of a converged warp.  Warp 9 issues the
//                   scores S_h = Q_h K_h^T (SS, K-major operands, head = 32-byte slice of the 128-byte
//                   swizzled row); warps 10 / 11 issue O_h += P_h V_h for the heads of softmax
//                   warpgroup 0 / 1 (TS: P read from TMEM, V as MN-major operand straight from the
//                   row-major tile).  Issue cost (~100 clk per tcgen05 op) is what the split hides;
//   2 softmax warpgroups: thread = one destination token (TMEM lane); row max / exp2 / row sum in
//                   registers, the NORMALISED probabilities written back to TMEM as bf16, so that the
//                   P V MMAs of all in-edges of the destination accumulate straight into one TMEM
//                   tile (the mean aggregation: no atomics, no per-edge read-back, no [E,F,d]
//                   message tensor, no [E,H,F,F] probabilities in HBM).
//
// Layouts: Q', K, V are bf16 [N, F, 64] node-major, rows of 128 bytes; Q' is pre-multiplied by
// log2(e)/sqrt(hd) so that scores are in the log2 domain.  lse2[p,h,i] = max + log2(sum) per
// destination-sorted edge slot p is saved for the backward recompute.
#include <cuda_bf16.h>
#include <math_constants.h>

#include "common.cuh"
#include "umma.cuh"

namespace ampconv {
namespace {

using namespace umma;

constexpr int kD = 64;                 // embed dim of this kernel family (one 128-byte swizzle atom per row)
constexpr int kTileBytes = 128 * 128;  // 128 tokens x 64 bf16
constexpr int kStages = 5;             // K/V ring depth
constexpr int kFwdThreads = 384;       // 2 softmax warpgroups + producer warp + 3 MMA-issuing warps
#ifndef AMP_POLY_FWD
#define AMP_POLY_FWD 1
#endif
// share of the exponentials evaluated on the FMA pipe (umma.cuh: ex2_poly2): 0 = none, 1 = a quarter, 2 = half
constexpr int kPolyShare = AMP_POLY_FWD;

struct NodeSlot {
  int node, p_begin, p_end;
  float inv_deg;
  int grp;     // head group of this work item (GROUPS == 2: head_dim 8, heads 4 grp .. 4 grp + 3 as padded head_dim-16 tiles)
};

// Reads a NodeSlot so that the compiler KNOWS the fields are warp-uniform (shfl from lane 0): the MMA warps' loop counters
// and descriptor arithmetic then live in uniform registers (a tcgen05.mma issue costs ~10 instructions instead of ~25).
__device__ __forceinline__ NodeSlot uniform_slot(const NodeSlot& s) {
  NodeSlot u;
  u.node = __shfl_sync(0xffffffffu, s.node, 0);
  u.p_begin = __shfl_sync(0xffffffffu, s.p_begin, 0);
  u.p_end = __shfl_sync(0xffffffffu, s.p_end, 0);
  u.inv_deg = __shfl_sync(0xffffffffu, s.inv_deg, 0);
  u.grp = __shfl_sync(0xffffffffu, s.grp, 0);
  return u;
}

struct FwdSmem {
  // tiles first (1024-byte aligned), then bookkeeping
  uint8_t q[2][kTileBytes];
  uint8_t kv[kStages][2][kTileBytes];
  uint64_t q_full[2], q_empty[2];
  uint64_t kv_full[kStages], kv_empty[kStages];
  uint64_t s_full[2], s_empty[2], p_full[2], p_empty[2];
  uint64_t o_full[2], o_empty[2];     // per node parity: O tile complete / read back
  NodeSlot slot[2];
  uint32_t tmem_base;
};

// first failure wins: status = code | blockIdx << 16
#define AMP_FAIL(code)                                                       \
  do {                                                                       \
    atomicCAS(status, 0, (int)((code) | (blockIdx.x << 16)));                \
    goto fail;                                                               \
  } while (0)
#define AMP_WAIT(bar, parity, code)                         \
  do {                                                      \
    if (!mbar_wait((bar), (parity))) AMP_FAIL(code);        \
  } while (0)

// GROUPS == 2 serves head_dim 8 (embed 64, 8 heads: the ogbn-products shape) with the head_dim-16 pipeline: a work item is
// (node, head group g); the producer warp builds its tiles with cp.async (umma.cuh: load_padded_tile): the 8 real columns of
// every head land in the first half of a 16-column slot whose second half stays zero, so Q_h K_h^T and P_h V_h see 8 real +
// 8 zero columns per head and nothing padded ever exists in HBM; the epilogue writes the 8 real columns of every head to
// their true place.
template <int HD, int GROUPS, bool PROF>
__global__ void __launch_bounds__(kFwdThreads, 1)
attn_fwd_bf16_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                     const __grid_constant__ CUtensorMap mapV, const int32_t* __restrict__ rowptr,
                     const int32_t* __restrict__ dst_src, const float* __restrict__ inv_deg,
                     const int32_t* __restrict__ order, int* __restrict__ counter, int* __restrict__ status,
                     float* __restrict__ agg, float* __restrict__ lse2, int N, int F, int accumulate,
                     const uint8_t* __restrict__ gq, const uint8_t* __restrict__ gk, const uint8_t* __restrict__ gv,
                     long long* __restrict__ prof) {
  constexpr int H = kD / HD;        // heads
  constexpr int HL = H / 2;         // heads per softmax warpgroup (head h belongs to warpgroup h & 1)
  static_assert(H % 2 == 0, "this kernel splits heads between two warpgroups");
  extern __shared__ uint8_t smem_raw[];
  FwdSmem& sm = *reinterpret_cast<FwdSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 9) tmem_alloc(&sm.tmem_base, 512);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sm.q_full[i], 1);
      mbar_init(&sm.q_empty[i], 1 + 8 + 2);
      mbar_init(&sm.s_full[i], 1);
      mbar_init(&sm.s_empty[i], 4);
      mbar_init(&sm.p_full[i], 4);
      mbar_init(&sm.p_empty[i], 1);
      mbar_init(&sm.o_full[i], 2);
      mbar_init(&sm.o_empty[i], 8);
    }
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&sm.kv_full[i], 1);
      mbar_init(&sm.kv_empty[i], 3);
    }
    fence_barrier_init();
  }
  if (warp == 8 && lane == 0) {
    prefetch_tensormap(&mapQ);
    prefetch_tensormap(&mapK);
    prefetch_tensormap(&mapV);
  }
  if (GROUPS == 2) {
    // the pad half of every head slot and the rows >= F are written here once and never again
    uint4* z = reinterpret_cast<uint4*>(sm.q[0]);
    for (int i = threadIdx.x; i < (2 + 2 * kStages) * kTileBytes / 16; i += blockDim.x) z[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;
  // TMEM columns: S[b] = b*128 (fp32 scores), P[b] = 256 + b*64 (bf16 pairs), O[node parity] = 384 + 64*par + h*HD
  const int nqk = ((F + 15) >> 4) << 4;   // MMA N of the score tile

  if (warp == 8) {
    // ------------------------------------------------------------------ producer / scheduler
    // The whole warp walks the node list (dynamic scheduler); lane 0 drives the mbarriers and TMA.
    // Isolated nodes never enter the pipeline: the warp writes their zero rows directly.
    if constexpr (GROUPS == 1) {
    uint32_t qi = 0, ei = 0;
    for (;;) {
      int node = -1, pb = 0, pe = 0;
      const int grp = 0;
      if (lane == 0) {
        const int ni = atomicAdd(counter, 1);
        node = ni < N ? (order ? order[ni] : ni) : -1;
        if (node >= 0) {
          pb = rowptr[node];
          pe = rowptr[node + 1];
        }
      }
      node = __shfl_sync(0xffffffffu, node, 0);
      pb = __shfl_sync(0xffffffffu, pb, 0);
      pe = __shfl_sync(0xffffffffu, pe, 0);
      if (node >= 0 && pe == pb) {
        if (accumulate) continue;   // a later ring phase: the node's rows already hold the earlier phases' sum
        float4* z = reinterpret_cast<float4*>(agg + (int64_t)node * F * kD);
        for (int i = lane; i < F * (kD / 4); i += 32) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        continue;
      }
      int failed = 0;
      if (lane == 0) {
        const uint32_t qb = qi & 1;
        if (!mbar_wait(&sm.q_empty[qb], ((qi >> 1) & 1) ^ 1)) {
          failed = 101;
        } else {
          NodeSlot ns;
          ns.node = node;
          ns.p_begin = pb;
          ns.p_end = pe;
          ns.inv_deg = node >= 0 ? inv_deg[node] : 0.f;
          ns.grp = grp;
          sm.slot[qb] = ns;
          if (node < 0) {
            mbar_arrive(&sm.q_full[qb]);
          } else {
            mbar_arrive_expect_tx(&sm.q_full[qb], kTileBytes);
            tma_load_3d(sm.q[qb], &mapQ, &sm.q_full[qb], 0, 0, node);
            int src_next = dst_src[pb];
            for (int p = pb; p < pe; ++p, ++ei) {
              const int src = src_next;
              if (p + 1 < pe) src_next = dst_src[p + 1];
              const uint32_t st = ei % kStages;
              if (!mbar_wait(&sm.kv_empty[st], ((ei / kStages) & 1) ^ 1)) {
                failed = 102;
                break;
              }
              mbar_arrive_expect_tx(&sm.kv_full[st], 2 * kTileBytes);
              tma_load_3d(sm.kv[st][0], &mapK, &sm.kv_full[st], 0, 0, src);
              tma_load_3d(sm.kv[st][1], &mapV, &sm.kv_full[st], 0, 0, src);
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
    // head_dim 8: the whole warp copies (cp.async); a load unit (the Q tile of a work item, or an edge's K and V tiles) is one
    // commit group, and the full barrier of unit u is raised when unit u + 1 has been issued (wait_group 1, proxy fence,
    // arrive): every wait of this warp only depends on units older than the newest one, so nothing can deadlock.
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
      int node = -1, pb = 0, pe = 0, grp = 0;
      if (lane == 0) {
        const int idx = atomicAdd(counter, 1);
        const int ni = idx / GROUPS;
        grp = idx - ni * GROUPS;
        node = ni < N ? (order ? order[ni] : ni) : -1;
        if (node >= 0) {
          pb = rowptr[node];
          pe = rowptr[node + 1];
        }
      }
      node = __shfl_sync(0xffffffffu, node, 0);
      pb = __shfl_sync(0xffffffffu, pb, 0);
      pe = __shfl_sync(0xffffffffu, pe, 0);
      grp = __shfl_sync(0xffffffffu, grp, 0);
      if (node >= 0 && pe == pb) {
        if (accumulate || grp != 0) continue;   // later ring phase: rows hold the earlier sum; group 0 zero-fills the whole row
        float4* z = reinterpret_cast<float4*>(agg + (int64_t)node * F * kD);
        for (int i = lane; i < F * (kD / 4); i += 32) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        continue;
      }
      const uint32_t qb = qi & 1;
      int failed = 0;
      if (lane == 0) {
        if (!mbar_wait(&sm.q_empty[qb], ((qi >> 1) & 1) ^ 1)) {
          failed = 101;
        } else {
          NodeSlot ns;
          ns.node = node;
          ns.p_begin = pb;
          ns.p_end = pe;
          ns.inv_deg = node >= 0 ? inv_deg[node] : 0.f;
          ns.grp = grp;
          sm.slot[qb] = ns;
          if (node < 0) mbar_arrive(&sm.q_full[qb]);
        }
      }
      failed = __shfl_sync(0xffffffffu, failed, 0);
      if (failed) AMP_FAIL(failed);
      if (node < 0) break;
      load_padded_tile(smem_u32(sm.q[qb]), gq, node, F, grp, lane);
      retire(smem_u32(&sm.q_full[qb]));
      for (int p = pb; p < pe; ++p, ++ei) {
        const int src = dst_src[p];
        const uint32_t st = ei % kStages;
        if (lane == 0 && !mbar_wait(&sm.kv_empty[st], ((ei / kStages) & 1) ^ 1)) failed = 102;
        failed = __shfl_sync(0xffffffffu, failed, 0);
        if (failed) AMP_FAIL(failed);
        load_padded_tile(smem_u32(sm.kv[st][0]), gk, src, F, grp, lane);
        load_padded_tile(smem_u32(sm.kv[st][1]), gv, src, F, grp, lane);
        retire(smem_u32(&sm.kv_full[st]));
      }
      ++qi;
    }
    retire(0u);        // raises the last unit's barrier (the empty group it commits completes at once)
    cp_async_wait<0>();
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------------ score MMAs (all heads)
    {
      const uint32_t idesc_qk = idesc_bf16(128, nqk, 0, 0);
      uint32_t qi = 0, edge = 0;
      for (;; ++qi) {
        const uint32_t qb = qi & 1;
        AMP_WAIT(&sm.q_full[qb], (qi >> 1) & 1, 201);
        const NodeSlot ns = uniform_slot(sm.slot[qb]);
        if (ns.node < 0) break;
        for (int p = ns.p_begin; p < ns.p_end; ++p, ++edge) {
          const uint32_t st = edge % kStages;
          AMP_WAIT(&sm.kv_full[st], (edge / kStages) & 1, 202);
#pragma unroll
          for (int h = 0; h < H; ++h) {
            const uint32_t b = h & 1;
            const uint32_t cc = edge * HL + (h >> 1);
            AMP_WAIT(&sm.s_empty[b], (cc & 1) ^ 1, 203);
            tc_fence_after();
            const uint64_t qd = smem_desc(smem_u32(sm.q[qb]) + h * (HD * 2), 16, 1024, LAYOUT_SW128);
            const uint64_t kd = smem_desc(smem_u32(sm.kv[st][0]) + h * (HD * 2), 16, 1024, LAYOUT_SW128);
#pragma unroll
            for (int ks = 0; ks < HD / 16; ++ks)
              mma_ss_w(tmem + b * 128, desc_advance(qd, ks * 32), desc_advance(kd, ks * 32), idesc_qk, ks > 0);
            mma_commit_w(&sm.s_full[b]);
          }
          mma_commit_w(&sm.kv_empty[st]);                       // K tile consumed (1 of 3 arrivals)
          if (p + 1 == ns.p_end) mma_commit_w(&sm.q_empty[qb]);  // Q tile consumed
        }
      }
    }
  } else if (warp >= 10) {
    // ------------------------------------------------------------------ P V MMAs for softmax warpgroup b = warp - 10
    {
      const uint32_t b = warp - 10;
      const uint32_t idesc_pv = idesc_bf16(128, HD, 0, 1);
      uint32_t qi = 0, edge = 0, c = 0;
      for (;; ++qi) {
        const uint32_t qb = qi & 1;
        AMP_WAIT(&sm.q_full[qb], (qi >> 1) & 1, 211);
        const NodeSlot ns = uniform_slot(sm.slot[qb]);
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.q_empty[qb]);
        if (ns.node < 0) break;
        const uint32_t par = qi & 1;
        // the first P V of this node overwrites the O tile of node qi - 2: wait until it has been read back
        AMP_WAIT(&sm.o_empty[par], ((qi >> 1) & 1) ^ 1, 212);
        for (int p = ns.p_begin; p < ns.p_end; ++p, ++edge) {
          const uint32_t st = edge % kStages;
          const uint32_t keep = p != ns.p_begin;                 // accumulate over the node's edges
#pragma unroll
          for (int hl = 0; hl < HL; ++hl, ++c) {
            const int h = 2 * hl + b;
            AMP_WAIT(&sm.p_full[b], c & 1, 213);
            tc_fence_after();
            const uint64_t vdesc = smem_desc(smem_u32(sm.kv[st][1]) + h * (HD * 2), 16, 1024, LAYOUT_SW128);
            const uint32_t o_col = tmem + 384 + par * 64 + h * HD;
            const uint32_t p_col = tmem + 256 + b * 64;
            // all eight K steps, always: P columns >= F are exact zeros (masked scores) and meet zero-filled V rows
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              mma_ts_w(o_col, p_col + 8 * ks, desc_advance(vdesc, ks * 2048), idesc_pv, ks > 0 ? 1u : keep);
            mma_commit_w(&sm.p_empty[b]);
          }
          mma_commit_w(&sm.kv_empty[st]);                         // V tile consumed by this warp's heads
        }
        mma_commit_w(&sm.o_full[par]);                            // this warp's heads of the node are complete
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax warpgroups
    const uint32_t b = warp >> 2;
    const int row = (warp & 3) * 32 + lane;
    const bool row_ok = row < F;
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t qi = 0, c = 0;
    // optional phase timers (debug entry point only): cycles spent by warp 0 of CTA 0 in each phase
    const bool do_prof = PROF && blockIdx.x == 0;
    long long pt[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long tp = do_prof ? clock64() : 0;
#define AMP_PHASE(i) do { if (do_prof) { const long long now_ = clock64(); pt[i] += now_ - tp; tp = now_; } } while (0)
    for (;;) {
      const uint32_t qb = qi & 1;
      AMP_WAIT(&sm.q_full[qb], (qi >> 1) & 1, 301);
      AMP_PHASE(8);
      // only the edge range lives in registers across the item loop (the softmax threads hold 192 values); node, inv_deg and
      // grp are re-read from the slot in the node epilogue, so this warp releases the slot (q_empty) only after that
      const int ns_node = sm.slot[qb].node, ns_pb = sm.slot[qb].p_begin, ns_pe = sm.slot[qb].p_end;
      const int ns_grp = GROUPS == 1 ? 0 : sm.slot[qb].grp;
      if (ns_node < 0) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.q_empty[qb]);
        break;
      }
      for (int p = ns_pb; p < ns_pe; ++p) {
#pragma unroll
        for (int hl = 0; hl < HL; ++hl) {
          const int h = 2 * hl + b;
          AMP_WAIT(&sm.s_full[b], c & 1, 302);
          AMP_PHASE(0);
          tc_fence_after();
          uint32_t s[128];
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tmem_ld_32x32b_x32(lane_base + b * 128 + 32 * k, *reinterpret_cast<uint32_t(*)[32]>(&s[32 * k]));
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm.s_empty[b]);
          AMP_PHASE(1);
          // mask source tokens >= F (their K rows are TMA zero fill, the score columns may be stale)
          if (F < 128) {
#pragma unroll
            for (int j = 0; j < 128; ++j)
              if (j >= F) s[j] = __float_as_uint(-CUDART_INF_F);
          }
          // row max: 8 chains of 3-input max (FMNMX3)
          float mx[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) mx[u] = fmaxf(__uint_as_float(s[u]), __uint_as_float(s[8 + u]));
#pragma unroll
          for (int j = 16; j < 128; j += 16)
#pragma unroll
            for (int u = 0; u < 8; ++u) mx[u] = max3(mx[u], __uint_as_float(s[j + u]), __uint_as_float(s[j + 8 + u]));
          const float m = fmaxf(max3(mx[0], mx[1], mx[2]), max3(max3(mx[3], mx[4], mx[5]), mx[6], mx[7]));
          AMP_PHASE(2);
          // exponentials: subtract and row sum as packed fp32 pairs (FADD2: half the issue slots of scalar FADD)
          float2 la = make_float2(0.f, 0.f), lb = la;
          const float2 nm = make_float2(-m, -m);
          uint32_t pk[64];
#pragma unroll
          for (int j = 0; j < 64; j += 2) {
            const float2 a2 = f2add(make_float2(__uint_as_float(s[2 * j]), __uint_as_float(s[2 * j + 1])), nm);
            const float2 b2 = f2add(make_float2(__uint_as_float(s[2 * j + 2]), __uint_as_float(s[2 * j + 3])), nm);
            const float2 ea = make_float2(ex2_approx(a2.x), ex2_approx(a2.y));
            // one pair in four goes to the FMA pipe instead of MUFU (kPolyShare, see ex2_poly2)
            const float2 eb = (kPolyShare == 2 || (kPolyShare == 1 && (j & 2))) ? ex2_poly2(b2)
                                                                                 : make_float2(ex2_approx(b2.x), ex2_approx(b2.y));
            la = f2add(la, ea);
            lb = f2add(lb, eb);
            pk[j] = pack_bf16x2(ea.x, ea.y);
            pk[j + 1] = pack_bf16x2(eb.x, eb.y);
          }
          const float l0 = la.x + la.y, l1 = lb.x + lb.y;
          const float l = l0 + l1;
          // normalise in bf16x2 so that the P V MMAs of all in-edges can accumulate into one TMEM tile
          const float inv_l = 1.0f / l;
          const uint32_t inv2 = pack_bf16x2(inv_l, inv_l);
#pragma unroll
          for (int j = 0; j < 64; ++j) pk[j] = mul_bf16x2(pk[j], inv2);
          AMP_PHASE(4);
          AMP_WAIT(&sm.p_empty[b], (c & 1) ^ 1, 303);
          AMP_PHASE(3);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tmem_st_32x32b_x16(lane_base + 256 + b * 64 + 16 * k, *reinterpret_cast<uint32_t(*)[16]>(&pk[16 * k]));
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm.p_full[b]);
          if (row_ok) lse2[((int64_t)p * (H * GROUPS) + ns_grp * H + h) * ((F + 3) & ~3) + row] = m + __log2f(l);
          AMP_PHASE(5);
          ++c;
        }
      }
      // node epilogue: all P V MMAs of the node have landed in O[parity]; read this warpgroup's heads
      {
        const uint32_t par = qi & 1;
        AMP_WAIT(&sm.o_full[par], (qi >> 1) & 1, 304);
        AMP_PHASE(6);
        tc_fence_after();
        uint32_t o[HL][HD];
#pragma unroll
        for (int hl = 0; hl < HL; ++hl) {
          if constexpr (HD == 16) {
            tmem_ld_32x32b_x16(lane_base + 384 + par * 64 + (2 * hl + b) * HD, o[hl]);
          } else {
            tmem_ld_32x32b_x32(lane_base + 384 + par * 64 + (2 * hl + b) * HD, o[hl]);
          }
        }
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.o_empty[par]);
        NodeSlot ns;
        ns.node = sm.slot[qb].node;
        ns.inv_deg = sm.slot[qb].inv_deg;
        ns.grp = GROUPS == 1 ? 0 : sm.slot[qb].grp;
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.q_empty[qb]);   // slot contents are in registers (8 of the 11 arrivals)
        if (row_ok) {
          float* dst = agg + ((int64_t)ns.node * F + row) * kD;
#pragma unroll
          for (int hl = 0; hl < HL; ++hl) {
            // GROUPS == 2: the 8 real columns of padded head 2 hl + b go to columns [32 grp + 8 (2 hl + b), + 8)
            float4* o4 = reinterpret_cast<float4*>(dst + (GROUPS == 1 ? (2 * hl + b) * HD : 32 * ns.grp + 8 * (2 * hl + b)));
#pragma unroll
            for (int x = 0; x < (GROUPS == 1 ? HD : 8); x += 4) {
              float4 r = make_float4(__uint_as_float(o[hl][x]) * ns.inv_deg, __uint_as_float(o[hl][x + 1]) * ns.inv_deg,
                                     __uint_as_float(o[hl][x + 2]) * ns.inv_deg, __uint_as_float(o[hl][x + 3]) * ns.inv_deg);
              if (accumulate) {   // ring phases (multi-GPU): the mean is a sum over all phases' edges, inv_deg is the full one
                const float4 a = o4[x >> 2];
                r.x += a.x; r.y += a.y; r.z += a.z; r.w += a.w;
              }
              o4[x >> 2] = r;
            }
          }
        }
      }
      AMP_PHASE(9);
      ++qi;
    }
    if (do_prof && threadIdx.x == 0) {
      for (int i = 0; i < 10; ++i) prof[i] = pt[i];
      prof[10] = c;
    }
#undef AMP_PHASE
  }
fail:
  if (GROUPS == 2) cp_async_wait<0>();   // no copy into this CTA's shared memory may outlive it
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem, 512);
}

}  // namespace
}  // namespace ampconv

using namespace ampconv;

extern "C" int ampconv_attn_bf16_supported(int F, int d, int H) {
  if (d != kD || H <= 0 || d % H) return 0;
  const int hd = d / H;
  return (hd == 16 || hd == 32 || hd == 8) && F >= 1 && F <= 128;   // hd 8: two head groups through zero-padding TMA boxes
}

static int attn_fwd_bf16_impl(const void* q, const void* k, const void* v,
                              const int32_t* dst_rowptr, const int32_t* dst_src, const float* inv_deg,
                              const int32_t* order, float* agg, float* lse2,
                              int64_t N, int64_t N_kv, int64_t E, int F, int d, int H,
                              void* workspace, size_t workspace_bytes, void* stream_, long long* prof,
                              int64_t n_work = -1, int accumulate = 0) {
  AMPCONV_REQUIRE(N >= 0 && E >= 0 && F > 0 && d > 0 && H > 0 && d % H == 0);
  if (!ampconv_attn_bf16_supported(F, d, H)) return AMPCONV_ERR_UNSUPPORTED;
  if (n_work < 0) n_work = N;        // length of the `order` work list (all destinations unless a ring phase lists fewer)
  if (N == 0 || n_work == 0) return AMPCONV_OK;
  AMPCONV_REQUIRE(n_work <= N && (n_work == N || order != nullptr));
  AMPCONV_REQUIRE(q && k && v && dst_rowptr && inv_deg && agg && workspace && (E == 0 || (dst_src && lse2)));
  if (workspace_bytes < 256) return AMPCONV_ERR_WORKSPACE;
  cudaStream_t stream = as_stream(stream_);
  CUtensorMap mq, mk, mv;
  const int hd = d / H;
  if (!make_tensor_map_bf16_3d(&mq, q, kD, F, N, kD, 128) || !make_tensor_map_bf16_3d(&mk, k, kD, F, N_kv, kD, 128) ||
      !make_tensor_map_bf16_3d(&mv, v, kD, F, N_kv, kD, 128))
    return AMPCONV_ERR_CUDA;   // (head_dim 8 loads its tiles with cp.async from the raw pointers instead)
  int* counter = reinterpret_cast<int*>(workspace);
  int* status = counter + 1;
  AMPCONV_CUDA_TRY(cudaMemsetAsync(counter, 0, sizeof(int), stream));   // the status word (counter + 1) is the caller's: zeroed once per layer call
  const size_t smem = sizeof(FwdSmem) + 1024;
  const int64_t items = n_work * (hd == 8 ? 2 : 1);
  const int grid = (int)(items < sm_count() ? items : sm_count());
#define AMP_LAUNCH_FWD(HDV, GRP, PROFV)                                                                                    \
  do {                                                                                                                \
    AMPCONV_CUDA_TRY(cudaFuncSetAttribute(attn_fwd_bf16_kernel<HDV, GRP, PROFV>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                          (int)smem));                                                                \
    attn_fwd_bf16_kernel<HDV, GRP, PROFV><<<grid, kFwdThreads, smem, stream>>>(mq, mk, mv, dst_rowptr, dst_src, inv_deg, order, \
                                                                         counter, status, agg, lse2, (int)n_work, F, accumulate,         \
                                                                         reinterpret_cast<const uint8_t*>(q),                \
                                                                         reinterpret_cast<const uint8_t*>(k),                \
                                                                         reinterpret_cast<const uint8_t*>(v), prof);         \
  } while (0)
  if (hd == 8) {
    AMP_LAUNCH_FWD(16, 2, false);
  } else if (hd == 16) {
    if (prof) AMP_LAUNCH_FWD(16, 1, true); else AMP_LAUNCH_FWD(16, 1, false);
  } else {
    if (prof) AMP_LAUNCH_FWD(32, 1, true); else AMP_LAUNCH_FWD(32, 1, false);
  }
#undef AMP_LAUNCH_FWD
  AMPCONV_CHECK_LAUNCH();
  return AMPCONV_OK;
}

extern "C" int ampconv_attn_fwd_bf16(const void* q, const void* k, const void* v,
                                     const int32_t* dst_rowptr, const int32_t* dst_src, const float* inv_deg,
                                     const int32_t* order, float* agg, float* lse2,
                                     int64_t N, int64_t E, int F, int d, int H,
                                     void* workspace, size_t workspace_bytes, void* stream_) {
  return attn_fwd_bf16_impl(q, k, v, dst_rowptr, dst_src, inv_deg, order, agg, lse2, N, N, E, F, d, H, workspace,
                            workspace_bytes, stream_, nullptr);
}

// Debug variant: additionally fills prof[0..10] (device, 11 x int64) with the cycles warp 0 of CTA 0 spent per
// softmax phase (wait S, load S, max, wait P slot, exp+store, publish, wait O, accumulate, node wait, node epilogue)
// and its item count.  Not part of the product path.
extern "C" int ampconv_attn_fwd_bf16_profile(const void* q, const void* k, const void* v,
                                             const int32_t* dst_rowptr, const int32_t* dst_src, const float* inv_deg,
                                             const int32_t* order, float* agg, float* lse2,
                                             int64_t N, int64_t E, int F, int d, int H,
                                             void* workspace, size_t workspace_bytes, void* stream_, long long* prof) {
  return attn_fwd_bf16_impl(q, k, v, dst_rowptr, dst_src, inv_deg, order, agg, lse2, N, N, E, F, d, H, workspace,
                            workspace_bytes, stream_, prof);
}

// Destination-partitioned variant (multi-GPU): q covers the num_nodes local destinations, k / v the num_kv_nodes rows of
// the all-gathered tensors; dst_src holds ids into k / v.
extern "C" int ampconv_attn_fwd_bf16_part(const void* q, const void* k, const void* v,
                                          const int32_t* dst_rowptr, const int32_t* dst_src, const float* inv_deg,
                                          const int32_t* order, float* agg, float* lse2, int64_t num_nodes,
                                          int64_t num_kv_nodes, int64_t E,
                                          int F, int d, int H, void* workspace, size_t workspace_bytes, void* stream_) {
  AMPCONV_REQUIRE(num_kv_nodes > 0 || E == 0);
  return attn_fwd_bf16_impl(q, k, v, dst_rowptr, dst_src, inv_deg, order, agg, lse2, num_nodes,
                            num_kv_nodes > 0 ? num_kv_nodes : 1, E, F, d, H, workspace, workspace_bytes, stream_, nullptr);
}

// Ring-phase variant (multi-GPU, ampnet_b200/distributed.py): the rank's edges are split by the owner of their source; one
// launch per phase over the `n_work` destinations listed in `order` that have an edge in the phase.  accumulate = 0: first
// phase (rows of destinations without an edge are zero-filled, results overwrite agg); accumulate = 1: agg += this phase.
// rowptr / dst_src / lse2 are the phase's own (lse2 indexed by the phase's destination-sorted slots).
extern "C" int ampconv_attn_fwd_bf16_phase(const void* q, const void* k, const void* v,
                                           const int32_t* dst_rowptr, const int32_t* dst_src, const float* inv_deg,
                                           const int32_t* order, int64_t n_work, int accumulate, float* agg, float* lse2,
                                           int64_t num_nodes, int64_t num_kv_nodes, int64_t E,
                                           int F, int d, int H, void* workspace, size_t workspace_bytes, void* stream_) {
  AMPCONV_REQUIRE(num_kv_nodes > 0 || E == 0);
  AMPCONV_REQUIRE(n_work >= 0 && (order != nullptr || n_work == 0));
  return attn_fwd_bf16_impl(q, k, v, dst_rowptr, dst_src, inv_deg, order, agg, lse2, num_nodes,
                            num_kv_nodes > 0 ? num_kv_nodes : 1, E, F, d, H, workspace, workspace_bytes, stream_, nullptr,
                            n_work, accumulate);
}

// Reads back the protocol status word written by the bf16 kernels (0 = ok).  Synchronises the stream.
extern "C" int ampconv_bf16_status(const void* workspace, int* status_host, void* stream_) {
  AMPCONV_REQUIRE(workspace && status_host);
  cudaStream_t stream = as_stream(stream_);
  AMPCONV_CUDA_TRY(cudaMemcpyAsync(status_host, reinterpret_cast<const int*>(workspace) + 1, sizeof(int),
                                   cudaMemcpyDeviceToHost, stream));
  AMPCONV_CUDA_TRY(cudaStreamSynchronize(stream));
  return AMPCONV_OK;
}
