/*
 * ampconv.h -- C ABI of the B200-native AMPConv hot path (libampconv.so).
 *
 * The reference (HarryL-Git/ampnet) has no FFI: its "operator API" for this path is the
 * PyTorch module AMPConv (src/ampnet/conv/amp_conv.py:9-51), which dispatches to PyG's
 * MessagePassing.propagate (gather x[src], x[dst]; scatter-mean at dst) and to
 * torch.nn.MultiheadAttention.  This header is what a host -- the Python mirror in
 * ampnet_b200/, or any other language with a C FFI -- binds instead of that eager chain.
 * Every entry point cites the reference step it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *   - every call returns 0 on success or a negative ampconv_status; nothing throws; the
 *     caller owns every buffer (the two opt-in side-output calls, ampconv_attn_weights_* and
 *     ampconv_edge_output_*, use stream-ordered temporaries); nothing synchronises the device
 *     except ampconv_graph_build (which must report out-of-range indices);
 *   - tokens per node F = width / d, head_dim hd = d / H, rows = N * F;
 *   - "position" p in [0,E) is an edge's slot in DESTINATION-sorted order (stable).
 */
#ifndef AMPCONV_H_
#define AMPCONV_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AMPCONV_ABI_VERSION 1

#if defined(__GNUC__)
#define AMPCONV_API __attribute__((visibility("default")))
#else
#define AMPCONV_API
#endif

typedef enum ampconv_status {
  AMPCONV_OK = 0,
  AMPCONV_ERR_INVALID_ARGUMENT = -1, /* null pointer, negative size, d % H != 0 ...           */
  AMPCONV_ERR_UNSUPPORTED = -2,      /* shape outside what the requested kernel family covers  */
  AMPCONV_ERR_INDEX_RANGE = -3,      /* edge_index holds a node id outside [0, N)              */
  AMPCONV_ERR_WORKSPACE = -4,        /* workspace too small                                    */
  AMPCONV_ERR_CUDA = -5,             /* a CUDA runtime call or launch failed (see last_cuda)   */
  AMPCONV_ERR_NO_DEVICE = -6         /* no sm_100 device                                       */
} ampconv_status;

AMPCONV_API int ampconv_abi_version(void);
AMPCONV_API const char* ampconv_strerror(int status);
/* cudaError_t of the most recent AMPCONV_ERR_CUDA on this thread (0 if none). */
AMPCONV_API int ampconv_last_cuda_error(void);
/* Number of kernels this library has launched so far in this process (monotonic). */
AMPCONV_API uint64_t ampconv_launch_count(void);
/* Fills SM count and compute capability of the current device. */
AMPCONV_API int ampconv_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------
 * Graph preparation.  Replaces the per-call bookkeeping of PyG propagate
 * (amp_conv.py:24-26: x_j = x[edge_index[0]], x_i = x[edge_index[1]], scatter-mean at
 * edge_index[1]) by two sorted views that are built once per edge_index and reused by both
 * layers and by forward and backward:
 *   by destination: dst_rowptr[N+1], dst_src[E] (source of the edge in slot p),
 *                   dst_eid[E]  (column of edge_index the slot came from);
 *   by source:      src_rowptr[N+1], src_dst[E], src_pos[E] (destination-sorted slot of the
 *                   same edge, to find its saved softmax statistics);
 *   inv_deg[N] = 1 / max(in_degree, 1), has_in[N] = in_degree > 0 ? 1 : 0.
 * Duplicate edges and self loops are ordinary edges.  Synchronises `stream`.
 * ------------------------------------------------------------------------------------------ */
AMPCONV_API int ampconv_graph_workspace_bytes(int64_t num_edges, int64_t num_nodes, size_t* bytes);
AMPCONV_API int ampconv_graph_build(const int64_t* edge_index /* [2,E] row 0 = src, row 1 = dst */,
                        int64_t num_edges, int64_t num_nodes,
                        int32_t* dst_rowptr, int32_t* dst_src, int32_t* dst_eid,
                        int32_t* src_rowptr, int32_t* src_dst, int32_t* src_pos,
                        float* inv_deg, float* has_in,
                        void* workspace, size_t workspace_bytes, void* stream);

/* Same for a destination-partitioned graph (multi-GPU): destinations are local ids in [0, num_nodes), sources
 * are ids in [0, num_src_nodes) of the all-gathered K/V tensors; src_rowptr has num_src_nodes + 1 entries. */
AMPCONV_API int ampconv_graph_build_bipartite(const int64_t* edge_index, int64_t num_edges, int64_t num_nodes,
                                  int64_t num_src_nodes,
                                  int32_t* dst_rowptr, int32_t* dst_src, int32_t* dst_eid,
                                  int32_t* src_rowptr, int32_t* src_dst, int32_t* src_pos,
                                  float* inv_deg, float* has_in,
                                  void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Strict fp32 family (CUDA cores, any F / d / H with hd <= 128).  Parity bar 1e-4 relative.
 * ------------------------------------------------------------------------------------------ */

/* qkv[r, 0:3d] = x[r, 0:d] @ in_proj_weight^T + in_proj_bias, once per NODE token instead of
 * once per edge (replaces F._in_projection_packed on [E,F,d] gathers;
 * custom_multihead_attn_forward.py:4031-4084). */
AMPCONV_API int ampconv_qkv_proj_f32(const float* x, const float* in_proj_weight, const float* in_proj_bias,
                         float* qkv, int64_t rows, int d, void* stream);

/* Fused per-edge multi-head attention + mean aggregation, destination-sorted, no atomics
 * (replaces head split, q*hd^-1/2, bmm, softmax, bmm of custom_multihead_attn_forward.py
 * :4140-4186,4376-4387 and PyG's scatter-mean, amp_conv.py:11).
 *   agg[n,i,:]  = inv_deg[n] * sum over in-edges of softmax_j(q_i k_j / sqrt(hd)) v_j
 *   lse[p,h,i]  = log-sum-exp of the scaled scores of edge slot p (saved for backward). */
AMPCONV_API int ampconv_attn_fwd_f32(const float* qkv, const int32_t* dst_rowptr, const int32_t* dst_src,
                         const float* inv_deg, float* agg, float* lse,
                         int64_t num_nodes, int64_t num_edges, int F, int d, int H, void* stream);

/* out[r,:] = agg[r,:] @ out_proj_weight^T + out_proj_bias * has_in[node(r)]
 * (out_proj of custom_multihead_attn_forward.py:4436-4437, moved after the mean). */
AMPCONV_API int ampconv_out_proj_f32(const float* agg, const float* out_proj_weight, const float* out_proj_bias,
                         const float* has_in, float* out, int64_t num_nodes, int F, int d, void* stream);

/* Head-averaged attention coefficients in ORIGINAL edge order (the attn_output_weights side
 * output, amp_conv.py:39; custom_multihead_attn_forward.py:4441-4442):
 *   weights[dst_eid[p], i, j] = mean_h exp(q_i k_j / sqrt(hd) - lse[p,h,i]). */
AMPCONV_API int ampconv_attn_weights_f32(const float* qkv, const float* lse, const int32_t* dst_rowptr,
                             const int32_t* dst_src, const int32_t* dst_eid, float* weights,
                             int64_t num_nodes, int64_t num_edges, int F, int d, int H, void* stream);

/* Chunked form (at the ogbn-arxiv shape all of [E,F,F] is 76 GB): weights[m,:,:] for the M listed
 * destination-sorted slots; slot_dst[p] = destination node of slot p, for every slot ([E]).  The callers
 * (synthetic_benchmark/visualize_attention_coefficients.py:221-232) index the result by original edge id; the host
 * maps ids to slots and walks them in slices. */
AMPCONV_API int ampconv_attn_weights_slots_f32(const float* qkv, const float* lse, const int32_t* slot_dst,
                                   const int32_t* dst_src, const int32_t* slots, int64_t M, float* weights,
                                   int F, int d, int H, void* stream);

/* Per-edge attention output after out_proj in ORIGINAL edge order (the attn_output side output,
 * amp_conv.py:39): edge_out[dst_eid[p], i, :] = (softmax(q k^T) v)[i,:] @ Wo^T + bo. */
AMPCONV_API int ampconv_edge_output_f32(const float* qkv, const float* lse, const int32_t* dst_rowptr,
                            const int32_t* dst_src, const int32_t* dst_eid,
                            const float* out_proj_weight, const float* out_proj_bias, float* edge_out,
                            int64_t num_nodes, int64_t num_edges, int F, int d, int H, void* stream);

/* Backward of ampconv_out_proj_f32.  d_agg is additionally multiplied by inv_deg[node], i.e. it
 * is the gradient w.r.t. every in-edge's un-normalised attention output.
 * d_w [d,d], d_b [d] are overwritten.  workspace: ampconv_param_grad_workspace_bytes(d, d). */
AMPCONV_API int ampconv_out_proj_bwd_f32(const float* d_out, const float* agg, const float* out_proj_weight,
                             const float* inv_deg, const float* has_in,
                             float* d_agg, float* d_w, float* d_b,
                             int64_t num_nodes, int F, int d,
                             void* workspace, size_t workspace_bytes, void* stream);

/* Backward of ampconv_attn_fwd_f32 (flash-style recompute from lse; autograd of
 * custom_multihead_attn_forward.py:4140-4186 + scatter-mean).  Two kernels, no atomics:
 *   _dq : destination-sorted; writes d_qkv[:, 0:d] and delta[p,h,i] = sum_j P_ij dP_ij;
 *   _dkv: source-sorted; reads delta; writes d_qkv[:, d:3d].
 * ampconv_attn_bwd_f32 runs both.  d_qkv [rows,3d] is overwritten; delta has the size of lse. */
AMPCONV_API int ampconv_attn_bwd_dq_f32(const float* qkv, const float* d_agg, const float* lse,
                            const int32_t* dst_rowptr, const int32_t* dst_src,
                            float* d_qkv, float* delta,
                            int64_t num_nodes, int64_t num_edges, int F, int d, int H, void* stream);
AMPCONV_API int ampconv_attn_bwd_dkv_f32(const float* qkv, const float* d_agg, const float* lse, const float* delta,
                             const int32_t* src_rowptr, const int32_t* src_dst, const int32_t* src_pos,
                             float* d_qkv,
                             int64_t num_nodes, int64_t num_edges, int F, int d, int H, void* stream);
AMPCONV_API int ampconv_attn_bwd_f32(const float* qkv, const float* d_agg, const float* lse,
                         const int32_t* dst_rowptr, const int32_t* dst_src,
                         const int32_t* src_rowptr, const int32_t* src_dst, const int32_t* src_pos,
                         float* d_qkv, float* delta,
                         int64_t num_nodes, int64_t num_edges, int F, int d, int H, void* stream);

/* Backward of ampconv_qkv_proj_f32: d_x [rows,d], d_w [3d,d], d_b [3d] are overwritten.
 * workspace: ampconv_param_grad_workspace_bytes(3d, d). */
AMPCONV_API int ampconv_qkv_proj_bwd_f32(const float* x, const float* d_qkv, const float* in_proj_weight,
                             float* d_x, float* d_w, float* d_b, int64_t rows, int d,
                             void* workspace, size_t workspace_bytes, void* stream);

AMPCONV_API int ampconv_param_grad_workspace_bytes(int out_dim, int in_dim, size_t* bytes);

/* ------------------------------------------------------------------------------------------
 * bf16 tensor-core family (tcgen05 + TMEM + TMA; d = 64, head_dim 16 or 32, F <= 128).
 * Parity bar 2e-2 relative (BASELINE.json north_star, "bf16 mode").
 * ------------------------------------------------------------------------------------------ */

/* 1 if the tcgen05 kernels cover this shape, else 0 (callers then use the fp32 family). */
AMPCONV_API int ampconv_attn_bf16_supported(int F, int d, int H);

/* Node-level in-projection (custom_multihead_attn_forward.py:4031-4084) emitting three bf16
 * [rows, d] tensors; q is additionally multiplied by q_scale (= log2(e)/sqrt(hd): the q*hd^-1/2
 * of :4173 folded together with the base-2 softmax). */
AMPCONV_API int ampconv_qkv_proj_bf16(const float* x, const float* in_proj_weight, const float* in_proj_bias,
                          void* q, void* k, void* v, int64_t rows, int d, float q_scale, void* stream);

/* Fused attention + mean aggregation on tensor cores (same contract as ampconv_attn_fwd_f32).
 * q/k/v: bf16 [N,F,d] from ampconv_qkv_proj_bf16; agg: fp32 [N,F,d];
 * lse2[p,h,i] = log2-sum-exp2 of the (log2-domain) scores of edge slot p, shape [E, H, roundup4(F)];
 * order: optional node processing order [N] (NULL = 0..N-1); workspace >= 256 bytes, ZEROED BY THE CALLER once per
 * layer call: int[0], [2], [4] are scheduler counters (reset by each launch), int[1] is the family's status word --
 * the first pipeline time-out of any launch that shares the workspace sticks there (see ampconv_bf16_status). */
AMPCONV_API int ampconv_attn_fwd_bf16(const void* q, const void* k, const void* v,
                          const int32_t* dst_rowptr, const int32_t* dst_src, const float* inv_deg,
                          const int32_t* order, float* agg, float* lse2,
                          int64_t num_nodes, int64_t num_edges, int F, int d, int H,
                          void* workspace, size_t workspace_bytes, void* stream);

/* Debug variant of ampconv_attn_fwd_bf16: also fills prof[0..10] (device int64) with per-phase cycle counts of
 * one softmax warp (see csrc/attn_bf16.cu).  Used by tools/phase_profile.py only. */
AMPCONV_API int ampconv_attn_fwd_bf16_profile(const void* q, const void* k, const void* v,
                                  const int32_t* dst_rowptr, const int32_t* dst_src, const float* inv_deg,
                                  const int32_t* order, float* agg, float* lse2,
                                  int64_t num_nodes, int64_t num_edges, int F, int d, int H,
                                  void* workspace, size_t workspace_bytes, void* stream, long long* prof);

/* Debug: device buffer of 64 int64 (or NULL) that makes the tcgen05 backward kernels record per-phase cycles. */
AMPCONV_API int ampconv_debug_set_bwd_profile(long long* prof);

/* Backward of ampconv_out_proj_f32 for the bf16 family: identical, except that d_agg (already
 * multiplied by inv_deg) is emitted as bf16 [N,F,d], the dO tile the tcgen05 backward kernels load. */
AMPCONV_API int ampconv_out_proj_bwd_bf16(const float* d_out, const float* agg, const float* out_proj_weight,
                              const float* inv_deg, const float* has_in,
                              void* d_agg_bf16, float* d_w, float* d_b,
                              int64_t num_nodes, int F, int d,
                              void* workspace, size_t workspace_bytes, void* stream);

/* Backward of ampconv_attn_fwd_bf16 on tensor cores (flash-style recompute from lse2), two kernels:
 *   _dq : destination-sorted; writes d_qkv[:, 0:d] (fp32 [rows,3d]) and delta[p,h,i];
 *   _dkv: source-sorted; reads delta; writes d_qkv[:, d:3d].
 * lse2 / delta use a row stride of roundup4(F): shape [E, H, roundup4(F)].
 * order: optional processing order of the pass's nodes (longest edge list first balances the persistent CTAs); NULL = 0..N-1.
 * Same workspace as the forward call (>= 256 bytes). */
AMPCONV_API int ampconv_attn_bwd_dq_bf16(const void* q, const void* k, const void* v, const void* d_agg_bf16,
                             const float* lse2, const int32_t* dst_rowptr, const int32_t* dst_src,
                             const int32_t* order, float* d_qkv, float* delta, int64_t num_nodes, int64_t num_edges, int F, int d, int H,
                             void* workspace, size_t workspace_bytes, void* stream);
AMPCONV_API int ampconv_attn_bwd_dkv_bf16(const void* q, const void* k, const void* v, const void* d_agg_bf16,
                              const float* lse2, const float* delta, const int32_t* src_rowptr,
                              const int32_t* src_dst, const int32_t* src_pos, const int32_t* order, float* d_qkv,
                              int64_t num_nodes, int64_t num_edges, int F, int d, int H,
                              void* workspace, size_t workspace_bytes, void* stream);

/* tcgen05 node-level projections (d = 64): same contracts as ampconv_qkv_proj_bf16, ampconv_out_proj_f32 and
 * the input-gradient halves of ampconv_out_proj_bwd_bf16 / ampconv_qkv_proj_bwd_f32, as HBM-bound persistent
 * kernels (fp32 rows are converted to bf16 on the way into shared memory, fp32 accumulation in TMEM).
 * `workspace` is the family's >= 256-byte workspace. */
AMPCONV_API int ampconv_qkv_proj_tc(const float* x, const float* in_proj_weight, const float* in_proj_bias,
                        void* q, void* k, void* v, int64_t rows, int d, float q_scale, void* workspace, void* stream);
AMPCONV_API int ampconv_out_proj_tc(const float* agg, const float* out_proj_weight, const float* out_proj_bias,
                        const float* has_in, float* out, int64_t num_nodes, int F, int d, void* workspace, void* stream);
AMPCONV_API int ampconv_out_proj_bwd_input_tc(const float* d_out, const float* out_proj_weight, const float* inv_deg,
                                  void* d_agg_bf16, int64_t num_nodes, int F, int d, void* workspace, void* stream);
AMPCONV_API int ampconv_qkv_proj_bwd_input_tc(const float* d_qkv, const float* in_proj_weight, float* d_x,
                                  int64_t rows, int d, void* workspace, void* stream);
/* Parameter-gradient halves (d_w, d_b) of the two projection backwards; workspace as for the combined calls. */
AMPCONV_API int ampconv_out_proj_bwd_params_f32(const float* d_out, const float* agg, const float* has_in,
                                    float* d_w, float* d_b, int64_t num_nodes, int F, int d,
                                    void* workspace, size_t workspace_bytes, void* stream);
AMPCONV_API int ampconv_qkv_proj_bwd_params_f32(const float* x, const float* d_qkv, float* d_w, float* d_b,
                                    int64_t rows, int d, void* workspace, size_t workspace_bytes, void* stream);

/* tcgen05 versions of the two parameter-gradient reductions (d = 64): both operands are MN-major bf16 tiles,
 * the bias gradient comes out of the same MMA through a gate column; per-CTA partials + deterministic reduce.
 * scratch: ampconv_param_grad_workspace_bytes(3d, d) bytes; workspace: the family's >= 256-byte workspace. */
AMPCONV_API int ampconv_out_proj_bwd_params_tc(const float* d_out, const float* agg, const float* has_in,
                                   float* d_w, float* d_b, int64_t num_nodes, int F, int d,
                                   void* scratch, size_t scratch_bytes, void* workspace, void* stream);
AMPCONV_API int ampconv_qkv_proj_bwd_params_tc(const float* x, const float* d_qkv, float* d_w, float* d_b,
                                   int64_t rows, int d, void* scratch, size_t scratch_bytes, void* workspace,
                                   void* stream);

/* The backward with bf16 gradient rows (the single-GPU product path): _dq_h / _dkv_h write d_qkv_bf16 [rows, 3d] bf16 instead
 * of fp32, _bwd_input_tc_h / _bwd_params_tc_h consume it.  The projection and weight-gradient kernels feed bf16 operands to
 * the tensor cores in either case, so the results are those of the fp32-row entry points; the rows cost a third of the HBM
 * traffic (written once, read twice). */
AMPCONV_API int ampconv_attn_bwd_dq_bf16_h(const void* q, const void* k, const void* v, const void* d_agg_bf16,
                               const float* lse2, const int32_t* dst_rowptr, const int32_t* dst_src,
                               const int32_t* order, void* d_qkv_bf16, float* delta, int64_t num_nodes, int64_t num_edges,
                               int F, int d, int H, void* workspace, size_t workspace_bytes, void* stream);
AMPCONV_API int ampconv_attn_bwd_dkv_bf16_h(const void* q, const void* k, const void* v, const void* d_agg_bf16,
                                const float* lse2, const float* delta, const int32_t* src_rowptr,
                                const int32_t* src_dst, const int32_t* src_pos, const int32_t* order, void* d_qkv_bf16,
                                int64_t num_nodes, int64_t num_edges, int F, int d, int H,
                                void* workspace, size_t workspace_bytes, void* stream);
AMPCONV_API int ampconv_qkv_proj_bwd_input_tc_h(const void* d_qkv_bf16, const float* in_proj_weight, float* d_x,
                                    int64_t rows, int d, void* workspace, void* stream);
AMPCONV_API int ampconv_qkv_proj_bwd_params_tc_h(const float* x, const void* d_qkv_bf16, float* d_w, float* d_b,
                                     int64_t rows, int d, void* scratch, size_t scratch_bytes, void* workspace,
                                     void* stream);

/* Destination-partitioned variants of the three attention kernels (multi-GPU): q / d_agg cover the
 * num_nodes LOCAL destinations, k / v the num_kv_nodes rows of the all-gathered K / V; graph views come from
 * ampconv_graph_build_bipartite.  _dq writes d_q fp32 [num_nodes*F, d]; _dkv writes the partial d_k | d_v
 * fp32 [num_kv_nodes*F, 2d] contributed by the local edges (the host reduce-scatters it to the owners). */
AMPCONV_API int ampconv_attn_fwd_bf16_part(const void* q, const void* k, const void* v,
                               const int32_t* dst_rowptr, const int32_t* dst_src, const float* inv_deg,
                               const int32_t* order, float* agg, float* lse2, int64_t num_nodes, int64_t num_kv_nodes,
                               int64_t num_edges,
                               int F, int d, int H, void* workspace, size_t workspace_bytes, void* stream);
AMPCONV_API int ampconv_attn_bwd_dq_bf16_part(const void* q, const void* k, const void* v, const void* d_agg_bf16,
                                  const float* lse2, const int32_t* dst_rowptr, const int32_t* dst_src,
                                  const int32_t* order, float* d_q, float* delta, int64_t num_nodes, int64_t num_kv_nodes,
                                  int64_t num_edges,
                                  int F, int d, int H, void* workspace, size_t workspace_bytes, void* stream);
AMPCONV_API int ampconv_attn_bwd_dkv_bf16_part(const void* q, const void* k, const void* v, const void* d_agg_bf16,
                                   const float* lse2, const float* delta, const int32_t* src_rowptr,
                                   const int32_t* src_dst, const int32_t* src_pos, const int32_t* order, float* d_kv,
                                   int64_t num_nodes, int64_t num_kv_nodes, int64_t num_edges, int F, int d, int H,
                                   void* workspace, size_t workspace_bytes, void* stream);

/* Halo-exchange variants (destination-partitioned multi-GPU, ampnet_b200/distributed.py; no counterpart in the reference,
 * whose only multi-process code is the unsynchronised gloo demo experiments/cora_benchmark_graphsaint_distributed.py:63,83).
 * _dkv_bf16_halo: sources [0, num_own) are this rank's nodes -> d_kv_own fp32 [num_own*F, 128]; sources
 * [num_own, num_kv_nodes) are halo nodes -> partial rows as bf16 in d_kv_halo [(num_kv_nodes-num_own)*F, 128].
 * ampconv_halo_add_bf16: acc[tgt[i], :] += sum over j in [rowptr[i], rowptr[i+1]) of recv[pos[j], :] (bf16 rows of
 * row_elems elements, fp32 accumulator rows), in list order: deterministic, no atomics. */
AMPCONV_API int ampconv_attn_bwd_dkv_bf16_halo(const void* q, const void* k, const void* v, const void* d_agg_bf16,
                                   const float* lse2, const float* delta, const int32_t* src_rowptr,
                                   const int32_t* src_dst, const int32_t* src_pos, const int32_t* order,
                                   float* d_kv_own, void* d_kv_halo, int64_t num_nodes, int64_t num_own,
                                   int64_t num_kv_nodes, int64_t num_edges, int F, int d, int H,
                                   void* workspace, size_t workspace_bytes, void* stream);
AMPCONV_API int ampconv_halo_add_bf16(const void* recv_bf16, const int32_t* tgt, const int32_t* rowptr, const int32_t* pos,
                          float* acc, int64_t n_tgt, int64_t row_elems, void* stream);

/* Ring-phase variants of the three attention kernels (multi-GPU overlap, ampnet_b200/distributed.py).  The rank's edges are
 * split by the OWNER of their source ("phase": 0 = own sources, t = the t-th owner in ring order); each phase has its own
 * destination- / source-sorted views (ampconv_graph_build_bipartite over the phase's edges, compact source ids unchanged)
 * and is one launch over the `n_work` nodes listed in `order`.  lse2 / delta are indexed by the phase's own slots.
 *   fwd / dq : accumulate = 0 -> first phase executed (overwrites; destinations without an edge are zero-filled),
 *              accumulate = 1 -> adds to agg / d_q (destinations without an edge in the phase are not in `order`);
 *   lse_map  : optional (NULL = identity): the backward may run over a COARSER edge set than the forward's phases (one dQ
 *              launch over all halo edges, dK|dV per owner from the same views); lse_map[slot of this pass] = index of the
 *              edge's statistics block in the forward's lse2; delta is always indexed by this pass's slots;
 *   dq       : row r of the result at d_q[r * d_q_ld] (a dense [rows, 64] tensor or the first columns of d_qkv [rows, 192]);
 *   dkv      : own sources (d_kv_halo NULL) -> d_kv_own fp32, row r at d_kv_own[r * own_ld], dK at column own_dk_col and dV
 *              at own_dv_col; the phase owner's halo sources, compact ids [halo_from, ...) (d_kv_own NULL) -> bf16 rows
 *              d_kv_halo[(id - halo_from)*F, 128] = dK | dV, the block that travels to that owner. */
AMPCONV_API int ampconv_attn_fwd_bf16_phase(const void* q, const void* k, const void* v,
                                const int32_t* dst_rowptr, const int32_t* dst_src, const float* inv_deg,
                                const int32_t* order, int64_t n_work, int accumulate, float* agg, float* lse2,
                                int64_t num_nodes, int64_t num_kv_nodes, int64_t num_edges,
                                int F, int d, int H, void* workspace, size_t workspace_bytes, void* stream);
AMPCONV_API int ampconv_attn_bwd_dq_bf16_phase(const void* q, const void* k, const void* v, const void* d_agg_bf16,
                                   const float* lse2, const int32_t* lse_map, const int32_t* dst_rowptr, const int32_t* dst_src,
                                   const int32_t* order, int64_t n_work, int accumulate, float* d_q, int64_t d_q_ld,
                                   float* delta,
                                   int64_t num_nodes, int64_t num_kv_nodes, int64_t num_edges, int F, int d, int H,
                                   void* workspace, size_t workspace_bytes, void* stream);
AMPCONV_API int ampconv_attn_bwd_dkv_bf16_phase(const void* q, const void* k, const void* v, const void* d_agg_bf16,
                                    const float* lse2, const int32_t* lse_map, const float* delta, const int32_t* src_rowptr,
                                    const int32_t* src_dst, const int32_t* src_pos, const int32_t* order,
                                    int64_t n_work, float* d_kv_own, int64_t own_ld, int64_t own_dk_col, int64_t own_dv_col,
                                    void* d_kv_halo, int64_t halo_from,
                                    int64_t num_nodes, int64_t num_kv_nodes, int64_t num_edges, int F, int d, int H,
                                    void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Peer memory over NVLink / NVSwitch (one process per GPU; csrc/peer.cu).  The exchange step of the partitioned path
 * is not a collective: every rank maps its peers' receive windows (CUDA IPC) and PUSHES rows into them with
 * stream-ordered device-to-device copies (copy engines: no SM is taken from the persistent attention kernels), then
 * raises a flag word in the receiver's window; the receiver's compute stream waits for the flag right before the
 * kernel that consumes the rows.
 *   _alloc / _free : a zeroed cudaMalloc window (IPC handles need a whole allocation);
 *   _export        : 64-byte IPC handle of a window; _open / _close: map / unmap a peer's window in this process;
 *   _copy          : dst (local or peer) <- src, `bytes` bytes, on `stream`;
 *   _ramp          : fills ramp[i] = i (the device-resident source of flag values); _signal: *peer_flag = value
 *                    (value < ramp length), ordered after everything enqueued on `stream` before it;
 *   _wait          : one-thread kernel on `stream` that returns when *flag == expected (ld.acquire.sys); after
 *                    budget_seconds it records 601 | expected << 16 in the status word of `workspace` and gives up;
 *   ampconv_gather_rows: dst[i, :] = src[idx[i], :] for rows of row_bytes (multiple of 16) bytes: packs the rows a
 *                    receiver needs into one contiguous block.
 * ------------------------------------------------------------------------------------------ */
AMPCONV_API int ampconv_peer_alloc(size_t bytes, void** ptr);
AMPCONV_API int ampconv_peer_free(void* ptr);
AMPCONV_API int ampconv_peer_export(void* ptr, void* handle_out /* 64 bytes, host */);
AMPCONV_API int ampconv_peer_open(const void* handle /* 64 bytes, host */, void** ptr);
AMPCONV_API int ampconv_peer_close(void* ptr);
AMPCONV_API int ampconv_peer_copy(void* dst, const void* src, size_t bytes, void* stream);
AMPCONV_API int ampconv_peer_ramp(int32_t* ramp, int n, void* stream);
AMPCONV_API int ampconv_peer_signal(int32_t* peer_flag, const int32_t* ramp, int32_t value, void* stream);
AMPCONV_API int ampconv_peer_wait(const int32_t* flag, int32_t expected, void* workspace, double budget_seconds, void* stream);
AMPCONV_API int ampconv_gather_rows(const void* src, const int64_t* idx, void* dst, int64_t n_rows, int64_t row_bytes,
                        void* stream);

/* ampconv_halo_add_bf16 into a strided accumulator: a received row holds row_elems / tok_elems token rows; token row t of node
 * n is acc[(n * tokens + t) * acc_ld + acc_col ...] -- e.g. the dK | dV columns [d, 3d) of one d_qkv [rows, 3d] tensor. */
AMPCONV_API int ampconv_halo_add_bf16_strided(const void* recv_bf16, const int32_t* tgt, const int32_t* rowptr, const int32_t* pos,
                                  float* acc, int64_t n_tgt, int64_t row_elems, int64_t tok_elems, int64_t acc_ld,
                                  int64_t acc_col, void* stream);

/* Copies the status word of the last bf16 kernel that used `workspace` to the host (0 = ok, otherwise
 * the id of the pipeline wait that timed out).  Synchronises `stream`; meant for tests and debugging. */
AMPCONV_API int ampconv_bf16_status(const void* workspace, int* status_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AMPCONV_H_ */
