"""Destination-partitioned AMPConv over one process per GPU.

The reference has no working multi-GPU path (its gloo demo never synchronises,
``experiments/cora_benchmark_graphsaint_distributed.py:63,83``).  Every edge is an independent attention
problem and the aggregation is a sum per destination, so the path shards by destination:

* destinations are range-partitioned into ``world`` contiguous ranges balanced by in-edge count;
  rank r owns the rows of x / Q / out / dX of its range and projects K, V for its own nodes;
* exchange step, forward: a HALO exchange of the projected K and V (bf16) -- every rank receives exactly the
  remote source rows its in-edges reference (``all_to_all_single`` with per-owner counts; the index lists are
  exchanged once per graph) and addresses them by compact id: own rows first, halo rows behind them;
* exchange step, backward: each rank computes the partial dK | dV its local edges contribute to every source it
  references; the halo rows travel back to their owners (bf16 by default) and are added there in a fixed order
  (deterministic); dQ never leaves the rank;
* the four parameter gradients are all-reduced.

Host logic only: the kernels are the ``*_part`` entry points of include/ampconv.h.  The collectives go through
``torch.distributed`` (NCCL over NVLink on the GPU box; the CPU tests run the same host logic over gloo).
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib
from . import functional as F_


# ------------------------------------------------------------------------------------------ partitioning (pure torch, CPU or GPU)
def partition_ranges(in_degree, world):
    """Boundaries [world + 1] of contiguous destination ranges with (nearly) equal in-edge counts."""
    n = in_degree.numel()
    csum = torch.cumsum(in_degree.to(torch.int64), 0)
    total = int(csum[-1]) if n > 0 else 0
    bounds = [0]
    for r in range(1, world):
        target = total * r // world
        # first node index whose prefix sum exceeds the target (ties keep ranges non-empty where possible)
        idx = int(torch.searchsorted(csum, torch.tensor(target, dtype=torch.int64, device=csum.device), right=True))
        idx = max(idx, bounds[-1])
        bounds.append(min(idx, n))
    bounds.append(n)
    return torch.tensor(bounds, dtype=torch.int64)


class PartitionedGraph:
    """Local view of rank ``rank``: the edges whose destination it owns, destinations as local ids, sources as
    compact ids (``[0, n_local)`` = own nodes, ``n_local + j`` = j-th halo node; halo nodes are sorted by global id,
    hence grouped by owner)."""

    def __init__(self, edge_index, num_nodes, world, rank, bounds=None):
        src, dst = edge_index[0], edge_index[1]
        if bounds is None:
            bounds = partition_ranges(torch.bincount(dst, minlength=num_nodes).cpu(), world)
        self.bounds = bounds
        self.world, self.rank, self.num_nodes = world, rank, int(num_nodes)
        self.lo, self.hi = int(bounds[rank]), int(bounds[rank + 1])
        self.n_local = self.hi - self.lo
        mine = (dst >= self.lo) & (dst < self.hi)
        self.edge_ids = torch.nonzero(mine, as_tuple=False).squeeze(1)       # columns of the global edge_index
        s = src[mine]
        own = (s >= self.lo) & (s < self.hi)
        self.halo_ids = torch.unique(s[~own])                                # sorted global ids of the remote sources
        self.n_halo = int(self.halo_ids.numel())
        halo_pos = torch.searchsorted(self.halo_ids, s) if self.n_halo > 0 else torch.zeros_like(s)
        compact = torch.where(own, s - self.lo, self.n_local + halo_pos)
        self.local_edge_index = torch.stack([compact, dst[mine] - self.lo]).contiguous()
        self.num_kv_nodes = self.n_local + self.n_halo
        owner = torch.searchsorted(bounds.to(self.halo_ids.device), self.halo_ids, right=True) - 1
        self.recv_counts = torch.bincount(owner, minlength=world).cpu().tolist()   # halo rows per owner (0 for this rank)
        self.send_counts = None     # rows this rank sends to every other rank       } filled by build_plan()
        self.send_idx = None        # local node ids of those rows, grouped by receiver }
        self.graph = None           # device CSR, built lazily

    def global_id(self, compact):
        """Global node id of a compact source id."""
        compact = compact.to(self.halo_ids.device)
        halo = self.halo_ids[(compact - self.n_local).clamp_min(0)] if self.n_halo > 0 else compact
        return torch.where(compact < self.n_local, compact + self.lo, halo)

    def build_plan(self, group=None):
        """One-time exchange of the halo index lists: afterwards every rank knows which of its rows each peer needs."""
        if self.send_counts is not None:
            return self
        dev = self.halo_ids.device
        recv = torch.tensor(self.recv_counts, dtype=torch.int64, device=dev)
        send = torch.empty_like(recv)
        dist.all_to_all_single(send, recv, group=group)
        self.send_counts = send.cpu().tolist()
        want = torch.empty(int(sum(self.send_counts)), dtype=torch.int64, device=dev)
        dist.all_to_all_single(want, self.halo_ids.contiguous(), output_split_sizes=self.send_counts,
                               input_split_sizes=self.recv_counts, group=group)
        self.send_idx = (want - self.lo).contiguous()
        # receive-side plan of the backward exchange: per own node, the positions of its partial rows in the receive
        # buffer, in sender order (stable sort) -> deterministic accumulation
        order = torch.sort(self.send_idx, stable=True)
        self.add_pos = order.indices.to(torch.int32).contiguous()
        tgt, counts = torch.unique_consecutive(order.values, return_counts=True)
        self.add_tgt = tgt.to(torch.int32).contiguous()
        rowptr = torch.zeros(tgt.numel() + 1, dtype=torch.int64, device=dev)
        rowptr[1:] = torch.cumsum(counts, 0)
        self.add_rowptr = rowptr.to(torch.int32).contiguous()
        return self

    def device_graph(self):
        if self.graph is None:
            self.graph = BipartiteGraph(self.local_edge_index, self.n_local, self.num_kv_nodes)
        return self.graph


class BipartiteGraph:
    """CSR by local destination and by padded source (ampconv_graph_build_bipartite)."""

    def __init__(self, edge_index, num_dst, num_src):
        if not edge_index.is_cuda:
            raise TypeError("edge_index must be a CUDA tensor")
        dev = edge_index.device
        e = edge_index.size(1)
        self.num_edges, self.num_nodes, self.num_src = e, int(num_dst), int(num_src)
        i32 = dict(dtype=torch.int32, device=dev)
        self.dst_rowptr = torch.empty(num_dst + 1, **i32)
        self.dst_src = torch.empty(e, **i32)
        self.dst_eid = torch.empty(e, **i32)
        self.src_rowptr = torch.empty(num_src + 1, **i32)
        self.src_dst = torch.empty(e, **i32)
        self.src_pos = torch.empty(e, **i32)
        self.inv_deg = torch.empty(num_dst, dtype=torch.float32, device=dev)
        self.has_in = torch.empty(num_dst, dtype=torch.float32, device=dev)
        nbytes = ctypes.c_size_t(0)
        with torch.cuda.device(dev):
            _lib.call("ampconv_graph_workspace_bytes", _lib.i64(e), _lib.i64(max(num_dst, num_src)), ctypes.byref(nbytes))
            ws = torch.empty(max(nbytes.value, 1), dtype=torch.uint8, device=dev)
            _lib.call("ampconv_graph_build_bipartite", edge_index.contiguous(), _lib.i64(e), _lib.i64(num_dst),
                      _lib.i64(num_src), self.dst_rowptr, self.dst_src, self.dst_eid, self.src_rowptr, self.src_dst,
                      self.src_pos, self.inv_deg, self.has_in, ws, _lib.size_t(ws.numel()),
                      _lib.stream_ptr(torch.cuda.current_stream(dev)))
        self.order_dst = torch.argsort(self.dst_rowptr[1:] - self.dst_rowptr[:-1], descending=True).to(torch.int32)
        self.order_src = torch.argsort(self.src_rowptr[1:] - self.src_rowptr[:-1], descending=True).to(torch.int32)


# ------------------------------------------------------------------------------------------ collectives (NCCL or gloo)
def halo_gather(local_rows, halo_out, pg, group=None):
    """Forward exchange: local_rows [n_local, C] -> halo_out [n_halo, C], the rows of the remote sources this rank's
    edges reference (row j = node pg.halo_ids[j]).  Requires pg.build_plan()."""
    send = local_rows.index_select(0, pg.send_idx)
    dist.all_to_all_single(halo_out, send, output_split_sizes=pg.recv_counts, input_split_sizes=pg.send_counts, group=group)
    return halo_out


def halo_scatter_add(halo_rows, local_acc, pg, group=None):
    """Backward exchange: halo_rows [n_halo, C] (partial sums for remote sources) travel to their owners and are added
    into local_acc [n_local, C] (fp32), one sender block after the other (unique ids per block: deterministic)."""
    recv = halo_rows.new_empty((int(sum(pg.send_counts)),) + tuple(halo_rows.shape[1:]))
    dist.all_to_all_single(recv, halo_rows.contiguous(), output_split_sizes=pg.send_counts, input_split_sizes=pg.recv_counts,
                           group=group)
    o = 0
    for cnt in pg.send_counts:
        if cnt:
            local_acc.index_add_(0, pg.send_idx[o:o + cnt], recv[o:o + cnt].to(local_acc.dtype))
        o += cnt
    return local_acc


# ------------------------------------------------------------------------------------------ the layer
GRAD_EXCHANGE_DTYPE = torch.bfloat16   # dtype of the dK | dV halo rows on the wire (torch.float32: exact partial sums)


class _PhaseTimer:
    """Optional per-phase device timing of the partitioned step (AMPNET_B200_DIST_TIMING=1): CUDA events on the current
    stream, summarised by ``phase_summary()``.  Off by default: no events, no synchronisation."""

    def __init__(self):
        import os
        self.on = os.environ.get("AMPNET_B200_DIST_TIMING", "0") == "1"
        self.marks = []

    def mark(self, name):
        if self.on:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.marks.append((name, ev))

    def summary(self):
        if not self.marks:
            return {}
        torch.cuda.synchronize()
        acc, cnt = {}, {}
        for (n0, e0), (n1, e1) in zip(self.marks[:-1], self.marks[1:]):
            if n1 == "start":
                continue
            acc[n1] = acc.get(n1, 0.0) + e0.elapsed_time(e1)
            cnt[n1] = cnt.get(n1, 0) + 1
        self.marks = []
        return {k: acc[k] / cnt[k] for k in acc}


TIMER = _PhaseTimer()


def phase_summary():
    """Mean milliseconds per phase since the last call (empty unless AMPNET_B200_DIST_TIMING=1)."""
    return TIMER.summary()


class _DistAMPConvFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_local, w_in, b_in, w_out, b_out, pg, num_heads, group):
        dev = x_local.device
        g = pg.device_graph()
        pg.build_plan(group)
        n, width = x_local.shape
        d = w_in.shape[1]
        f = width // d
        hd = d // num_heads
        rows, krows = n * f, pg.num_kv_nodes * f
        with torch.cuda.device(dev):
            st = F_._stream(dev)
            ws = torch.zeros(64, dtype=torch.int32, device=dev)
            TIMER.mark("start")
            q = torch.empty((rows, d), dtype=torch.bfloat16, device=dev)
            # K, V of own nodes first, halo rows behind them (compact ids of pg.local_edge_index)
            k_all = torch.empty((krows, d), dtype=torch.bfloat16, device=dev)
            v_all = torch.empty((krows, d), dtype=torch.bfloat16, device=dev)
            _lib.call("ampconv_qkv_proj_tc", x_local, w_in, b_in, q, k_all, v_all, _lib.i64(rows), _lib.i32(d),
                      _lib.f32(F_.LOG2E / hd ** 0.5), ws, st)
            TIMER.mark("fwd qkv projection")
            if pg.world > 1:
                halo_gather(k_all[:rows].view(n, f * d), k_all[rows:].view(pg.n_halo, f * d), pg, group)
                halo_gather(v_all[:rows].view(n, f * d), v_all[rows:].view(pg.n_halo, f * d), pg, group)
            TIMER.mark("fwd halo exchange K|V")
            agg = torch.empty((rows, d), dtype=torch.float32, device=dev)
            lse2 = torch.empty((g.num_edges, num_heads, (f + 3) // 4 * 4), dtype=torch.float32, device=dev)
            out = torch.empty((n, width), dtype=torch.float32, device=dev)
            _lib.call("ampconv_attn_fwd_bf16_part", q, k_all, v_all, g.dst_rowptr, g.dst_src, g.inv_deg, g.order_dst, agg, lse2,
                      _lib.i64(n), _lib.i64(pg.num_kv_nodes), _lib.i64(g.num_edges), _lib.i32(f), _lib.i32(d),
                      _lib.i32(num_heads), ws, _lib.size_t(256), st)
            TIMER.mark("fwd attention")
            _lib.call("ampconv_out_proj_tc", agg, w_out, b_out, g.has_in, out, _lib.i64(n), _lib.i32(f), _lib.i32(d), ws, st)
            TIMER.mark("fwd out projection")
        ctx.save_for_backward(x_local, w_in, w_out)
        ctx.state = (pg, group, num_heads, q, k_all, v_all, agg, lse2, ws)
        return out

    @staticmethod
    def backward(ctx, d_out):
        x_local, w_in, w_out = ctx.saved_tensors
        pg, group, h, q, k_all, v_all, agg, lse2, bws = ctx.state
        g = pg.device_graph()
        dev = x_local.device
        n, width = x_local.shape
        d = w_in.shape[1]
        f = width // d
        rows = n * f
        e = g.num_edges
        with torch.cuda.device(dev):
            st = F_._stream(dev)
            d_out = d_out.contiguous()
            TIMER.mark("loss (caller)")
            d_agg = torch.empty((rows, d), dtype=torch.bfloat16, device=dev)
            d_w_out = torch.empty_like(w_out)
            d_b_out = torch.empty(d, dtype=torch.float32, device=dev)
            ws = F_._param_grad_ws(3 * d, d, dev)
            _lib.call("ampconv_out_proj_bwd_input_tc", d_out, w_out, g.inv_deg, d_agg, _lib.i64(n), _lib.i32(f), _lib.i32(d),
                      bws, st)
            _lib.call("ampconv_out_proj_bwd_params_tc", d_out, agg, g.has_in, d_w_out, d_b_out, _lib.i64(n), _lib.i32(f),
                      _lib.i32(d), ws, _lib.size_t(ws.numel()), bws, st)
            TIMER.mark("bwd out projection")
            d_q = torch.empty((rows, d), dtype=torch.float32, device=dev)
            delta = torch.empty_like(lse2)
            tail = (_lib.i64(n), _lib.i64(pg.num_kv_nodes), _lib.i64(e), _lib.i32(f), _lib.i32(d), _lib.i32(h), bws,
                    _lib.size_t(256), st)
            _lib.call("ampconv_attn_bwd_dq_bf16_part", q, k_all, v_all, d_agg, lse2, g.dst_rowptr, g.dst_src, g.order_dst, d_q, delta,
                      *tail)
            TIMER.mark("bwd attention dQ")
            # partial dK | dV of the local edges for every referenced source: own rows (fp32) and halo rows
            if pg.world > 1 and GRAD_EXCHANGE_DTYPE == torch.bfloat16:
                d_kv = torch.empty((rows, 2 * d), dtype=torch.float32, device=dev)
                halo = torch.empty((pg.n_halo, f * 2 * d), dtype=torch.bfloat16, device=dev)
                _lib.call("ampconv_attn_bwd_dkv_bf16_halo", q, k_all, v_all, d_agg, lse2, delta, g.src_rowptr, g.src_dst,
                          g.src_pos, g.order_src, d_kv, halo, _lib.i64(n), _lib.i64(n), _lib.i64(pg.num_kv_nodes), _lib.i64(e),
                          _lib.i32(f), _lib.i32(d), _lib.i32(h), bws, _lib.size_t(256), st)
                TIMER.mark("bwd attention dK|dV")
                recv = torch.empty((int(sum(pg.send_counts)), f * 2 * d), dtype=torch.bfloat16, device=dev)
                dist.all_to_all_single(recv, halo, output_split_sizes=pg.send_counts, input_split_sizes=pg.recv_counts,
                                       group=group)
                _lib.call("ampconv_halo_add_bf16", recv, pg.add_tgt, pg.add_rowptr, pg.add_pos, d_kv,
                          _lib.i64(pg.add_tgt.numel()), _lib.i64(f * 2 * d), st)
                del recv, halo
            else:
                d_kv_partial = torch.empty((pg.num_kv_nodes * f, 2 * d), dtype=torch.float32, device=dev)
                _lib.call("ampconv_attn_bwd_dkv_bf16_part", q, k_all, v_all, d_agg, lse2, delta, g.src_rowptr, g.src_dst,
                          g.src_pos, g.order_src, d_kv_partial, *tail)
                TIMER.mark("bwd attention dK|dV")
                d_kv = d_kv_partial[:rows]
                if pg.world > 1:
                    halo_scatter_add(d_kv_partial[rows:].view(pg.n_halo, f * 2 * d), d_kv.view(n, f * 2 * d), pg, group)
            TIMER.mark("bwd halo exchange dK|dV")
            d_qkv = torch.cat([d_q, d_kv], dim=1)
            d_x = torch.empty_like(x_local)
            d_w_in = torch.empty_like(w_in)
            d_b_in = torch.empty(3 * d, dtype=torch.float32, device=dev)
            _lib.call("ampconv_qkv_proj_bwd_input_tc", d_qkv, w_in, d_x, _lib.i64(rows), _lib.i32(d), bws, st)
            _lib.call("ampconv_qkv_proj_bwd_params_tc", x_local, d_qkv, d_w_in, d_b_in, _lib.i64(rows), _lib.i32(d), ws,
                      _lib.size_t(ws.numel()), bws, st)
            TIMER.mark("bwd qkv projection")
            flat = torch.cat([d_w_in.flatten(), d_b_in, d_w_out.flatten(), d_b_out])
            if pg.world > 1:
                dist.all_reduce(flat, group=group)
            TIMER.mark("parameter-gradient all-reduce")
            o = 0
            outs = []
            for t in (d_w_in, d_b_in, d_w_out, d_b_out):
                outs.append(flat[o:o + t.numel()].view_as(t))
                o += t.numel()
        return d_x, outs[0], outs[1], outs[2], outs[3], None, None, None


def dist_amp_conv(x_local, pg, w_in, b_in, w_out, b_out, num_heads, group=None):
    """AMPConv forward for the rows this rank owns (x_local = x[lo:hi]); differentiable; bf16 (tcgen05) family only."""
    d = w_in.shape[1]
    f = x_local.shape[1] // d
    if F_.bf16_supported(f, d, num_heads):
        return _DistAMPConvFunction.apply(x_local.contiguous(), w_in, b_in, w_out, b_out, pg, num_heads, group)
    if F_.bf16_grouped_supported(f, d, num_heads):
        # head_dim 8 (the ogbn-products shape): two 4-head passes over zero-padded heads, composed by autograd
        # (functional.hd8_compose; the algebra is pinned on the CPU by tests/test_hd8_grouping_cpu.py, the partitioned
        # composition itself has not run on hardware yet -- DESIGN.md section 6)
        def layer(x_, wi, bi, wo, bo, heads):
            return _DistAMPConvFunction.apply(x_.contiguous(), wi.contiguous(), bi.contiguous(), wo.contiguous(),
                                              bo.contiguous(), pg, heads, group)
        return F_.hd8_compose(layer, x_local, w_in, b_in, w_out, b_out)
    raise ValueError("the partitioned path uses the tcgen05 family: embed_dim 64, head_dim 8, 16 or 32, F <= 128")
