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

Two transports for the exchange step:

* ``"peer"`` (default on CUDA): ring-phased PUSH over peer memory.  The rank's edges are split into ``world`` phases by the
  owner of their source (phase 0 = own sources, phase t = owner ``(rank + t) % world``).  Every rank maps its peers' receive
  windows (CUDA IPC, ``ampconv_peer_*``) and pushes the K|V rows a peer needs straight into that peer's K/V tensors with
  stream-ordered device-to-device copies on a side stream -- copy engines over NVLink, so the persistent attention kernels
  keep every SM -- followed by a flag word; the receiver's compute stream waits for the flag right before the phase that
  consumes the rows.  Forward: phase 0 computes while the rows of phase 1 travel, and so on around the ring.  Backward:
  per phase ``dQ`` then ``dK|dV`` of the phase's sources (halo phases first, own sources last); a phase's bf16 ``dK|dV``
  block is pushed to its owner while the next phase computes; the owner adds the blocks in a fixed order.  Nothing on the
  data path is a collective; NCCL carries one tiny barrier per step and the all-reduce of the parameter gradients.
* ``"nccl"``: the whole halo in one ``all_to_all_single`` per tensor, serial with compute (round 1; also what the gloo CPU
  tests exercise).

Host logic only: the kernels are the ``*_part`` / ``*_phase`` entry points of include/ampconv.h.
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
        # every rank's recv_counts (matrix [world, world]) and local size: the ring-phase plan needs the peers' layouts
        mine = torch.tensor(self.recv_counts + [self.n_local], dtype=torch.int64, device=dev)
        allc = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(allc, mine, group=group)
        allc = torch.stack(allc).cpu()
        self.recv_matrix = allc[:, :self.world].tolist()
        self.n_local_all = allc[:, self.world].tolist()
        self.phase_plan = PhasePlan(self.rank, self.world, self.n_local_all, self.recv_matrix)
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


# ------------------------------------------------------------------------------------------ ring phases (pure torch, CPU or GPU)
class PhasePlan:
    """Where every block of the ring-phased exchange goes.  ``recv_matrix[p][o]`` = number of halo rows rank p receives from
    owner o (row p is rank p's ``recv_counts``; the matrix is all-gathered once per graph).  All offsets are in NODES.

    * ``ring[t]``        owner whose sources phase t consumes (``ring[0]`` = this rank);
    * ``src_range[t]``   compact-id range [lo, hi) of the phase's sources (own nodes for t = 0, else the owner's halo block:
                         halo nodes are sorted by global id, hence grouped by owner in ascending owner order);
    * forward, phase t >= 1: this rank pushes its rows ``send_idx[send_off[p] : send_off[p] + send_counts[p]]`` to
      ``p = fwd_dst[t] = (rank - t) % world`` at node offset ``fwd_dst_off[t]`` of p's K / V tensors (p's own nodes first,
      then p's halo blocks by owner) -- p consumes them in ITS phase t;
    * backward, phase t >= 1: the bf16 dK|dV block of owner ``o = ring[t]`` goes to node offset ``bwd_dst_off[t]`` of o's
      receive window (blocks by sender rank, ascending: the layout ``PartitionedGraph.build_plan`` derived the fixed-order
      add from); this rank receives from ``(rank - t) % world``."""

    def __init__(self, rank, world, n_local_all, recv_matrix):
        self.rank, self.world = rank, world
        rm = [[int(v) for v in row] for row in recv_matrix]
        self.recv_matrix = rm
        n_local = int(n_local_all[rank])
        self.ring = [(rank + t) % world for t in range(world)]
        hoff = [0] * (world + 1)
        for o in range(world):
            hoff[o + 1] = hoff[o] + rm[rank][o]
        self.halo_off = hoff
        self.src_range = [(0, n_local)] + [(n_local + hoff[o], n_local + hoff[o + 1]) for o in self.ring[1:]]
        send_counts = [rm[p][rank] for p in range(world)]
        soff = [0] * (world + 1)
        for p_ in range(world):
            soff[p_ + 1] = soff[p_] + send_counts[p_]
        self.send_counts, self.send_off = send_counts, soff
        self.fwd_dst, self.fwd_dst_off, self.fwd_rows = [None], [None], [0]
        self.bwd_dst_off, self.bwd_rows, self.bwd_src = [None], [0], [None]
        for t in range(1, world):
            p_ = (rank - t) % world
            self.fwd_dst.append(p_)
            self.fwd_dst_off.append(int(n_local_all[p_]) + sum(rm[p_][o] for o in range(rank)))
            self.fwd_rows.append(send_counts[p_])
            o = self.ring[t]
            self.bwd_dst_off.append(sum(rm[s_][o] for s_ in range(rank)))
            self.bwd_rows.append(rm[rank][o])
            self.bwd_src.append(p_)


def build_plans_local(pgs):
    """The one-time plan of ``PartitionedGraph.build_plan`` for ALL ranks of a (virtual) world held in one process -- no
    process group: tests run the ring-phase kernels of every virtual rank on one GPU, or check the plan on the CPU."""
    world = len(pgs)
    dev = pgs[0].halo_ids.device
    bounds = pgs[0].bounds.to(dev)
    rm = [pg.recv_counts for pg in pgs]
    n_local_all = [pg.n_local for pg in pgs]
    for pg in pgs:
        r = pg.rank
        pg.send_counts = [rm[p_][r] for p_ in range(world)]
        wants = []
        for p_ in range(world):
            h = pgs[p_].halo_ids
            owner = torch.searchsorted(bounds, h, right=True) - 1
            wants.append(h[owner == r])
        want = torch.cat(wants) if wants else torch.empty(0, dtype=torch.int64, device=dev)
        pg.send_idx = (want - pg.lo).contiguous()
        order = torch.sort(pg.send_idx, stable=True)
        pg.add_pos = order.indices.to(torch.int32).contiguous()
        tgt, counts = torch.unique_consecutive(order.values, return_counts=True)
        pg.add_tgt = tgt.to(torch.int32).contiguous()
        rowptr = torch.zeros(tgt.numel() + 1, dtype=torch.int64, device=dev)
        rowptr[1:] = torch.cumsum(counts, 0)
        pg.add_rowptr = rowptr.to(torch.int32).contiguous()
        pg.recv_matrix, pg.n_local_all = rm, n_local_all
        pg.phase_plan = PhasePlan(r, world, n_local_all, rm)
    return pgs


def phase_of_sources(compact_src, plan):
    """Phase index of every edge from its compact source id."""
    ph = torch.zeros_like(compact_src)
    for t in range(1, plan.world):
        lo, hi = plan.src_range[t]
        ph = torch.where((compact_src >= lo) & (compact_src < hi), torch.full_like(ph, t), ph)
    return ph


class PhaseGraphs:
    """Per-phase device CSR views of a rank's edges (one ``BipartiteGraph`` per phase over the same local destinations and
    compact source ids) plus the work lists of the ring-phase kernels."""

    def __init__(self, pg, plan):
        lei = pg.local_edge_index
        ph = phase_of_sources(lei[0], plan)
        self.graphs, self.order_dst, self.n_dst_active, self.order_src, self.edge_sel = [], [], [], [], []
        full = None
        for t in range(plan.world):
            sel = torch.nonzero(ph == t, as_tuple=False).squeeze(1)
            g = BipartiteGraph(lei[:, sel].contiguous(), pg.n_local, pg.num_kv_nodes)
            deg = g.dst_rowptr[1:] - g.dst_rowptr[:-1]
            self.graphs.append(g)
            self.edge_sel.append(sel)
            self.order_dst.append(g.order_dst)                       # descending degree: destinations without an edge last
            self.n_dst_active.append(int((deg > 0).sum()))
            lo, hi = plan.src_range[t]
            sdeg = g.src_rowptr[lo + 1:hi + 1] - g.src_rowptr[lo:hi]
            self.order_src.append((lo + torch.argsort(sdeg, descending=True)).to(torch.int32).contiguous())
            full = deg if full is None else full + deg
        # the mean uses the FULL in-degree of a destination, whatever phase an edge is in
        self.inv_deg = (1.0 / full.clamp(min=1).to(torch.float32)).contiguous()
        self.has_in = (full > 0).to(torch.float32).contiguous()
        self.max_phase_edges = max(g.num_edges for g in self.graphs)
        # the forward's lse2 is ONE tensor: phase t's block starts at edge offset lse_off[t] (its own destination-sorted slots)
        self.lse_off = [0]
        for g in self.graphs:
            self.lse_off.append(self.lse_off[-1] + g.num_edges)
        # coarse view for the backward (a destination appears once, not once per owner): ALL halo edges in one dQ launch, the
        # dK|dV launches per owner pick their sources from the same views.  halo_lse_map[slot] = forward lse2 index.
        self.halo = None
        if plan.world > 2:
            dev = lei.device
            fwd_index = torch.empty(lei.shape[1], dtype=torch.int64, device=dev)      # local edge -> forward lse2 index
            for t, (g, sel) in enumerate(zip(self.graphs, self.edge_sel)):
                if g.num_edges:
                    fwd_index[sel[g.dst_eid.to(torch.int64)]] = self.lse_off[t] + torch.arange(g.num_edges, device=dev)
            sel_h = torch.nonzero(ph != 0, as_tuple=False).squeeze(1)
            gh = BipartiteGraph(lei[:, sel_h].contiguous(), pg.n_local, pg.num_kv_nodes)
            self.halo = gh
            self.halo_lse_map = (fwd_index[sel_h[gh.dst_eid.to(torch.int64)]].to(torch.int32).contiguous()
                                 if gh.num_edges else torch.zeros(1, dtype=torch.int32, device=dev))
            self.halo_order_src = [None]
            for t in range(1, plan.world):
                lo, hi = plan.src_range[t]
                sdeg = gh.src_rowptr[lo + 1:hi + 1] - gh.src_rowptr[lo:hi]
                self.halo_order_src.append((lo + torch.argsort(sdeg, descending=True)).to(torch.int32).contiguous())


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


def forward_phases(q, k_all, v_all, pgs, num_kv_nodes, f, d, h, ws, stream, agg, before_phase=None, after_phase=None):
    """The ring phases of the forward attention on `stream`: phase 0 (own sources) overwrites agg, phases 1.. add to it;
    before_phase(t) is called right before phase t >= 1 is launched (the peer path waits there for the phase's K|V rows),
    after_phase(t) right after.  Returns lse2 [E_local, H, roundup4(F)]: phase t's block at edge offset pgs.lse_off[t]."""
    n = agg.shape[0] // f
    fs = (f + 3) // 4 * 4
    st = _lib.stream_ptr(stream)
    lse2 = torch.empty((max(pgs.lse_off[-1], 1), h, fs), dtype=torch.float32, device=agg.device)
    for t in range(len(pgs.graphs)):
        g = pgs.graphs[t]
        if t > 0 and before_phase is not None:
            before_phase(t)
        n_work = n if t == 0 else pgs.n_dst_active[t]
        _lib.call("ampconv_attn_fwd_bf16_phase", q, k_all, v_all, g.dst_rowptr, g.dst_src, pgs.inv_deg,
                  pgs.order_dst[t], _lib.i64(n_work), _lib.i32(0 if t == 0 else 1), agg, lse2[pgs.lse_off[t]:],
                  _lib.i64(n), _lib.i64(num_kv_nodes), _lib.i64(g.num_edges), _lib.i32(f), _lib.i32(d),
                  _lib.i32(h), ws, _lib.size_t(256), st)
        if after_phase is not None:
            after_phase(t)
    return lse2


COARSE_DELTA_BUDGET_BYTES = 16 << 30     # the coarse backward holds delta for all halo edges at once


def backward_phases(q, k_all, v_all, d_agg, lse2, pgs, plan, num_kv_nodes, f, d, h, ws, stream, d_qkv, send_slot,
                    after_halo_phase):
    """The attention backward of a rank on `stream`, halo sources first (their dK|dV blocks travel while the rest computes),
    own sources last.  d_qkv fp32 [rows, 3d] = dQ | dK | dV of the own rows; send_slot(t) -> (slot, bf16 buffer) receives the
    block of ring phase t's owner and after_halo_phase(t, slot) ships it.

    Coarse schedule (default when delta for all halo edges fits): ONE dQ launch over all halo edges -- a destination is
    visited once instead of once per owner -- then dK|dV per owner out of the same views, then dQ / dK|dV of the own-source
    edges.  Fine schedule: dQ and dK|dV per ring phase (delta only ever holds one phase)."""
    n = d_qkv.shape[0] // f
    fs = (f + 3) // 4 * 4
    st = _lib.stream_ptr(stream)
    world = len(pgs.graphs)
    dev = d_qkv.device
    tail = (_lib.i64(n), _lib.i64(num_kv_nodes))
    shp = (_lib.i32(f), _lib.i32(d), _lib.i32(h), ws, _lib.size_t(256), st)

    def dq(g, lse, lse_map, order, n_work, accumulate, delta):
        _lib.call("ampconv_attn_bwd_dq_bf16_phase", q, k_all, v_all, d_agg, lse, lse_map, g.dst_rowptr, g.dst_src,
                  order, _lib.i64(n_work), _lib.i32(accumulate), d_qkv, _lib.i64(3 * d), delta, *tail, _lib.i64(g.num_edges), *shp)

    def dkv_halo(g, lse, lse_map, delta, order, lo, hi, buf):
        _lib.call("ampconv_attn_bwd_dkv_bf16_phase", q, k_all, v_all, d_agg, lse, lse_map, delta, g.src_rowptr, g.src_dst,
                  g.src_pos, order, _lib.i64(hi - lo), None, _lib.i64(0), _lib.i64(0), _lib.i64(0), buf, _lib.i64(lo),
                  *tail, _lib.i64(g.num_edges), *shp)

    coarse = pgs.halo is not None and pgs.halo.num_edges * h * fs * 4 <= COARSE_DELTA_BUDGET_BYTES
    first = True
    if coarse:
        gh = pgs.halo
        delta = torch.empty((max(gh.num_edges, 1), h, fs), dtype=torch.float32, device=dev)
        dq(gh, lse2, pgs.halo_lse_map, gh.order_dst, n, 0, delta)
        first = False
        for t in range(1, world):
            lo, hi = plan.src_range[t]
            slot, buf = send_slot(t)
            if hi > lo:
                dkv_halo(gh, lse2, pgs.halo_lse_map, delta, pgs.halo_order_src[t], lo, hi, buf)
            after_halo_phase(t, slot)      # an empty block still raises its flag
        del delta
    delta = torch.empty((max(pgs.max_phase_edges if not coarse else pgs.graphs[0].num_edges, 1), h, fs), dtype=torch.float32,
                        device=dev)
    for t in ([] if coarse else list(range(1, world))) + [0]:
        g = pgs.graphs[t]
        lse_t = lse2[pgs.lse_off[t]:]
        dq(g, lse_t, None, pgs.order_dst[t], n if first else pgs.n_dst_active[t], 0 if first else 1, delta)
        first = False
        lo, hi = plan.src_range[t]
        if t == 0:
            _lib.call("ampconv_attn_bwd_dkv_bf16_phase", q, k_all, v_all, d_agg, lse_t, None, delta, g.src_rowptr,
                      g.src_dst, g.src_pos, pgs.order_src[t], _lib.i64(hi - lo), d_qkv, _lib.i64(3 * d), _lib.i64(d),
                      _lib.i64(2 * d), None, _lib.i64(num_kv_nodes), *tail, _lib.i64(g.num_edges), *shp)
            continue
        slot, buf = send_slot(t)
        if hi > lo:
            dkv_halo(g, lse_t, None, delta, pgs.order_src[t], lo, hi, buf)
        after_halo_phase(t, slot)


# ------------------------------------------------------------------------------------------ peer-memory engine (CUDA, world > 1)
class _RawCuda:
    """Adapter that lets torch view a window allocated by the C library (``ampconv_peer_alloc``)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class PeerWindow:
    """A cudaMalloc'ed window of this rank, mapped by every peer (CUDA IPC): ``.t`` = local uint8 view, ``.peer[p]`` =
    address of rank p's window in THIS process (``.peer[rank]`` = own address)."""

    def __init__(self, nbytes, dev, group):
        self.nbytes = max(int(nbytes), 256)
        self.dev = dev
        ptr = ctypes.c_void_p(0)
        with torch.cuda.device(dev):
            _lib.call("ampconv_peer_alloc", _lib.size_t(self.nbytes), ctypes.byref(ptr))
        self.ptr = int(ptr.value)
        self.t = torch.as_tensor(_RawCuda(self.ptr, self.nbytes), device=dev)
        if self.t.data_ptr() != self.ptr or self.t.device != dev:
            raise RuntimeError("torch copied the peer window instead of viewing it")
        handle = (ctypes.c_ubyte * 64)()
        with torch.cuda.device(dev):
            _lib.call("ampconv_peer_export", ctypes.c_void_p(self.ptr), handle)
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        handles = [None] * world
        dist.all_gather_object(handles, bytes(handle), group=group)
        self.peer = [0] * world
        self._opened = []
        for p_ in range(world):
            if p_ == rank:
                self.peer[p_] = self.ptr
                continue
            out = ctypes.c_void_p(0)
            buf = (ctypes.c_ubyte * 64).from_buffer_copy(handles[p_])
            with torch.cuda.device(dev):
                _lib.call("ampconv_peer_open", buf, ctypes.byref(out))
            self.peer[p_] = int(out.value)
            self._opened.append(int(out.value))

    def view(self, offset_bytes, shape, dtype):
        n = 1
        for v in shape:
            n *= int(v)
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        return self.t[offset_bytes:offset_bytes + nbytes].view(dtype).view(*shape)

    def close(self):
        try:
            with torch.cuda.device(self.dev):
                for a in self._opened:
                    _lib.call("ampconv_peer_close", ctypes.c_void_p(a))
                self._opened = []
                if self.ptr:
                    self.t = None
                    _lib.call("ampconv_peer_free", ctypes.c_void_p(self.ptr))
                    self.ptr = 0
        except Exception:
            pass


SEND_SLOT_BUDGET_BYTES = 12 << 30     # backward send staging: one slot per phase while it fits, else a ring of two
PEER_WAIT_SECONDS = 20.0              # a flag that does not arrive within this many seconds becomes status 601 (loud)


class PeerEngine:
    """Buffers, windows and streams of one partitioned AMPConv layer (one engine per layer: the K|V halo of a layer lives
    from its forward to its backward).  Collective to construct (IPC handle exchange), like ``build_plan``."""

    def __init__(self, pg, f, d, group=None):
        pg.build_plan(group)
        self.pg, self.plan, self.group = pg, pg.phase_plan, group
        self.f, self.d = f, d
        dev = pg.halo_ids.device
        self.dev = dev
        if getattr(pg, "phase_graphs", None) is None:
            pg.phase_graphs = PhaseGraphs(pg, self.plan)
        self.pgs = pg.phase_graphs
        world, n = pg.world, pg.n_local
        row_kv = f * d * 2                   # bytes of one node's K (or V) rows, bf16
        row_g = f * 2 * d * 2                # bytes of one node's dK|dV rows, bf16
        self.row_kv, self.row_g = row_kv, row_g
        kv_nodes = pg.num_kv_nodes
        a256 = lambda v: (v + 255) // 256 * 256
        self.k_off, self.v_off = 0, a256(kv_nodes * row_kv)
        # every peer lays its window out from ITS OWN node count: V of rank p starts behind p's K rows
        self.peer_v_off = [a256((int(pg.n_local_all[p_]) + sum(int(c) for c in pg.recv_matrix[p_])) * row_kv) for p_ in range(world)]
        assert self.peer_v_off[pg.rank] == self.v_off
        self.kv_win = PeerWindow(self.v_off + a256(kv_nodes * row_kv), dev, group)
        self.k_all = self.kv_win.view(self.k_off, (kv_nodes * f, d), torch.bfloat16)
        self.v_all = self.kv_win.view(self.v_off, (kv_nodes * f, d), torch.bfloat16)
        n_recv = int(sum(self.plan.send_counts))
        self.recv_win = PeerWindow(max(n_recv, 1) * row_g, dev, group)
        self.recv = self.recv_win.view(0, (max(n_recv, 1), f * 2 * d), torch.bfloat16)
        self.flag_win = PeerWindow(4096, dev, group)          # int32 [2][world]: forward / backward flag per sender
        self.flags = self.flag_win.view(0, (2, world), torch.int32)
        self.ramp = torch.empty(65536, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.call("ampconv_peer_ramp", self.ramp, _lib.i32(65536), F_._stream(dev))
        # forward staging: the K / V rows one receiver needs, packed (send_idx order); one slot per phase while it fits,
        # else a ring of two (a slot is repacked once its push has left)
        need_f = [2 * self.plan.fwd_rows[t] * row_kv for t in range(world)]
        self.n_fslots = max(1, world - 1 if sum(need_f) <= SEND_SLOT_BUDGET_BYTES else 2)
        frows = max([1] + self.plan.fwd_rows[1:])
        self.k_send = [torch.empty((frows, f * d), dtype=torch.bfloat16, device=dev) for _ in range(self.n_fslots)]
        self.v_send = [torch.empty((frows, f * d), dtype=torch.bfloat16, device=dev) for _ in range(self.n_fslots)]
        self.fslot_free = [None] * self.n_fslots
        # backward staging: the bf16 dK|dV block of a phase's owner
        need = [self.plan.bwd_rows[t] * row_g for t in range(world)]
        self.n_slots = world - 1 if sum(need) <= SEND_SLOT_BUDGET_BYTES else 2
        slot_rows = max([1] + self.plan.bwd_rows[1:])
        self.g_send = [torch.empty((slot_rows, f * 2 * d), dtype=torch.bfloat16, device=dev) for _ in range(max(self.n_slots, 1))]
        self.slot_free = [None] * max(self.n_slots, 1)          # event: the slot's last push has left
        self.comm = torch.cuda.Stream(device=dev)
        self.comm_done = None
        self.epoch = 0
        self._bar = torch.zeros(1, dtype=torch.int32, device=dev)
        torch.cuda.synchronize(dev)
        dist.barrier(group=group)

    # -- helpers -------------------------------------------------------------------------------------------------
    def next_epoch(self):
        self.epoch = self.epoch % 65535 + 1
        return self.epoch

    def _copy(self, dst_ptr, src_ptr, nbytes, stream):
        _lib.call("ampconv_peer_copy", ctypes.c_void_p(int(dst_ptr)), ctypes.c_void_p(int(src_ptr)), _lib.size_t(nbytes),
                  _lib.stream_ptr(stream))

    def _signal(self, peer, kind, stream):
        flag_ptr = self.flag_win.peer[peer] + 4 * (kind * self.pg.world + self.pg.rank)
        _lib.call("ampconv_peer_signal", ctypes.c_void_p(flag_ptr), self.ramp, _lib.i32(self.epoch), _lib.stream_ptr(stream))

    def wait_flag(self, sender, kind, ws, stream):
        flag_ptr = self.flag_win.ptr + 4 * (kind * self.pg.world + sender)
        _lib.call("ampconv_peer_wait", ctypes.c_void_p(flag_ptr), _lib.i32(self.epoch), ws, ctypes.c_double(PEER_WAIT_SECONDS),
                  _lib.stream_ptr(stream))

    def push_forward(self, t, compute):
        """Packs the K / V rows the receiver of ring phase t needs (compute stream) and pushes them into its K / V tensors
        on the copy stream, then raises its forward flag."""
        pg, plan = self.pg, self.plan
        p_ = plan.fwd_dst[t]
        cnt = plan.fwd_rows[t]
        slot = (t - 1) % self.n_fslots
        st = _lib.stream_ptr(compute)
        if cnt:
            if self.fslot_free[slot] is not None:
                compute.wait_event(self.fslot_free[slot])
            rows = pg.n_local * self.f
            idx = pg.send_idx[plan.send_off[p_]:plan.send_off[p_] + cnt]
            _lib.call("ampconv_gather_rows", self.k_all[:rows], idx, self.k_send[slot], _lib.i64(cnt), _lib.i64(self.row_kv), st)
            _lib.call("ampconv_gather_rows", self.v_all[:rows], idx, self.v_send[slot], _lib.i64(cnt), _lib.i64(self.row_kv), st)
        packed = torch.cuda.Event()
        packed.record(compute)
        self.comm.wait_event(packed)
        nb = cnt * self.row_kv
        dst_off = plan.fwd_dst_off[t] * self.row_kv
        self._copy(self.kv_win.peer[p_] + self.k_off + dst_off, self.k_send[slot].data_ptr(), nb, self.comm)
        self._copy(self.kv_win.peer[p_] + self.peer_v_off[p_] + dst_off, self.v_send[slot].data_ptr(), nb, self.comm)
        self._signal(p_, 0, self.comm)
        ev = torch.cuda.Event()
        ev.record(self.comm)
        self.fslot_free[slot] = ev

    def push_backward(self, t, slot, compute):
        """Phase t's dK|dV block (in send slot `slot`) -> its owner's receive window."""
        plan = self.plan
        o = plan.ring[t]
        done = torch.cuda.Event()
        done.record(compute)
        self.comm.wait_event(done)
        nb = plan.bwd_rows[t] * self.row_g
        self._copy(self.recv_win.peer[o] + plan.bwd_dst_off[t] * self.row_g, self.g_send[slot].data_ptr(), nb, self.comm)
        self._signal(o, 1, self.comm)
        ev = torch.cuda.Event()
        ev.record(self.comm)
        self.slot_free[slot] = ev

    def close(self):
        for w in (self.kv_win, self.recv_win, self.flag_win):
            w.close()


_engines = {}


def get_engine(pg, f, d, group=None, key=None):
    k = (id(pg), f, d, key)
    eng = _engines.get(k)
    if eng is None:
        eng = PeerEngine(pg, f, d, group)
        _engines[k] = eng
    return eng


def close_engines():
    for eng in _engines.values():
        eng.close()
    _engines.clear()


class _PeerAMPConvFunction(torch.autograd.Function):
    """Ring-phased partitioned layer over peer memory (module docstring, transport "peer")."""

    @staticmethod
    def forward(ctx, x_local, w_in, b_in, w_out, b_out, eng, num_heads):
        pg, plan, pgs = eng.pg, eng.plan, eng.pgs
        dev = x_local.device
        n, width = x_local.shape
        d = w_in.shape[1]
        f = width // d
        hd = d // num_heads
        rows = n * f
        fs = (f + 3) // 4 * 4
        world = pg.world
        with torch.cuda.device(dev):
            compute = torch.cuda.current_stream(dev)
            st = _lib.stream_ptr(compute)
            ws = torch.zeros(64, dtype=torch.int32, device=dev)
            TIMER.mark("start")
            # every rank has finished the previous step (its reads of the windows this step overwrites)
            dist.all_reduce(eng._bar, group=eng.group)
            if eng.comm_done is not None:
                compute.wait_event(eng.comm_done)
            eng.next_epoch()
            q = torch.empty((rows, d), dtype=torch.bfloat16, device=dev)
            _lib.call("ampconv_qkv_proj_tc", x_local, w_in, b_in, q, eng.k_all, eng.v_all, _lib.i64(rows), _lib.i32(d),
                      _lib.f32(F_.LOG2E / hd ** 0.5), ws, st)
            TIMER.mark("fwd qkv projection")
            for t in range(1, min(eng.n_fslots, world - 1) + 1):
                eng.push_forward(t, compute)
            TIMER.mark("fwd pack K|V rows")
            agg = torch.empty((rows, d), dtype=torch.float32, device=dev)
            out = torch.empty((n, width), dtype=torch.float32, device=dev)

            def after_phase(t):        # a staging slot is free again: pack and push the next receiver's rows
                nxt = t + 1 + eng.n_fslots
                if nxt < world:
                    eng.push_forward(nxt, compute)

            lse2 = forward_phases(q, eng.k_all, eng.v_all, pgs, pg.num_kv_nodes, f, d, num_heads, ws, compute, agg,
                                  before_phase=lambda t: eng.wait_flag(plan.ring[t], 0, ws, compute), after_phase=after_phase)
            TIMER.mark("fwd attention (ring phases, K|V pushes overlapped)")
            _lib.call("ampconv_out_proj_tc", agg, w_out, b_out, pgs.has_in, out, _lib.i64(n), _lib.i32(f), _lib.i32(d), ws, st)
            TIMER.mark("fwd out projection")
            ev = torch.cuda.Event()
            ev.record(eng.comm)
            eng.comm_done = ev
            F_._post_status(ws, "partitioned AMPConv forward")
        ctx.save_for_backward(x_local, w_in, w_out)
        ctx.state = (eng, num_heads, q, agg, lse2, ws)
        return out

    @staticmethod
    def backward(ctx, d_out):
        x_local, w_in, w_out = ctx.saved_tensors
        eng, h, q, agg, lse2, bws = ctx.state
        pg, plan, pgs = eng.pg, eng.plan, eng.pgs
        dev = x_local.device
        n, width = x_local.shape
        d = w_in.shape[1]
        f = width // d
        rows = n * f
        fs = (f + 3) // 4 * 4
        world = pg.world
        with torch.cuda.device(dev):
            compute = torch.cuda.current_stream(dev)
            st = _lib.stream_ptr(compute)
            d_out = d_out.contiguous()
            TIMER.mark("loss (caller)")
            d_agg = torch.empty((rows, d), dtype=torch.bfloat16, device=dev)
            d_w_out = torch.empty_like(w_out)
            d_b_out = torch.empty(d, dtype=torch.float32, device=dev)
            ws = F_._param_grad_ws(3 * d, d, dev)
            _lib.call("ampconv_out_proj_bwd_input_tc", d_out, w_out, pgs.inv_deg, d_agg, _lib.i64(n), _lib.i32(f), _lib.i32(d),
                      bws, st)
            _lib.call("ampconv_out_proj_bwd_params_tc", d_out, agg, pgs.has_in, d_w_out, d_b_out, _lib.i64(n), _lib.i32(f),
                      _lib.i32(d), ws, _lib.size_t(ws.numel()), bws, st)
            TIMER.mark("bwd out projection")
            d_qkv = torch.empty((rows, 3 * d), dtype=torch.float32, device=dev)

            def send_slot(t):
                slot = (t - 1) % eng.n_slots
                if eng.slot_free[slot] is not None:
                    compute.wait_event(eng.slot_free[slot])
                return slot, eng.g_send[slot]

            backward_phases(q, eng.k_all, eng.v_all, d_agg, lse2, pgs, plan, pg.num_kv_nodes, f, d, h, bws, compute, d_qkv,
                            send_slot, lambda t, slot: eng.push_backward(t, slot, compute))
            TIMER.mark("bwd attention dQ + dK|dV (ring phases, dK|dV pushes overlapped)")
            for t in range(1, world):
                eng.wait_flag(plan.bwd_src[t], 1, bws, compute)
            if pg.add_tgt.numel():
                _lib.call("ampconv_halo_add_bf16_strided", eng.recv, pg.add_tgt, pg.add_rowptr, pg.add_pos, d_qkv,
                          _lib.i64(pg.add_tgt.numel()), _lib.i64(f * 2 * d), _lib.i64(2 * d), _lib.i64(3 * d), _lib.i64(d), st)
            TIMER.mark("bwd wait for peers' dK|dV blocks + fixed-order add")
            d_x = torch.empty_like(x_local)
            d_w_in = torch.empty_like(w_in)
            d_b_in = torch.empty(3 * d, dtype=torch.float32, device=dev)
            _lib.call("ampconv_qkv_proj_bwd_input_tc", d_qkv, w_in, d_x, _lib.i64(rows), _lib.i32(d), bws, st)
            _lib.call("ampconv_qkv_proj_bwd_params_tc", x_local, d_qkv, d_w_in, d_b_in, _lib.i64(rows), _lib.i32(d), ws,
                      _lib.size_t(ws.numel()), bws, st)
            TIMER.mark("bwd qkv projection")
            flat = torch.cat([d_w_in.flatten(), d_b_in, d_w_out.flatten(), d_b_out])
            dist.all_reduce(flat, group=eng.group)
            TIMER.mark("parameter-gradient all-reduce")
            ev = torch.cuda.Event()
            ev.record(eng.comm)
            eng.comm_done = ev
            F_._post_status(bws, "partitioned AMPConv backward")
            o = 0
            outs = []
            for t_ in (d_w_in, d_b_in, d_w_out, d_b_out):
                outs.append(flat[o:o + t_.numel()].view_as(t_))
                o += t_.numel()
        return d_x, outs[0], outs[1], outs[2], outs[3], None, None


def default_transport(x_local, pg):
    import os
    t = os.environ.get("AMPNET_B200_EXCHANGE", "")
    if t:
        return t
    return "peer" if (x_local.is_cuda and pg.world > 1) else "nccl"


def dist_amp_conv(x_local, pg, w_in, b_in, w_out, b_out, num_heads, group=None, transport=None, key=None):
    """AMPConv forward for the rows this rank owns (x_local = x[lo:hi]); differentiable; bf16 (tcgen05) family only.
    transport: "peer" (ring-phased pushes over peer memory, overlapped with compute; default for world > 1) or "nccl" (one
    serial all-to-all per tensor).  key: distinguishes layers that share a PartitionedGraph (one peer engine per layer; the
    default key is the identity of in_proj_weight)."""
    d = w_in.shape[1]
    f = x_local.shape[1] // d
    F_.check_params(x_local, w_in, b_in, w_out, b_out)
    F_.check_status()
    transport = transport or default_transport(x_local, pg)
    if transport not in ("peer", "nccl"):
        raise ValueError(f"unknown transport {transport!r}")

    def layer(x_, wi, bi, wo, bo, heads, key_):
        if transport == "peer" and pg.world > 1:
            eng = get_engine(pg, f, d, group, key_)
            return _PeerAMPConvFunction.apply(x_.contiguous(), wi.contiguous(), bi.contiguous(), wo.contiguous(),
                                              bo.contiguous(), eng, heads)
        return _DistAMPConvFunction.apply(x_.contiguous(), wi.contiguous(), bi.contiguous(), wo.contiguous(),
                                          bo.contiguous(), pg, heads, group)

    if key is None:
        key = ("param", w_in.data_ptr())
    if F_.bf16_supported(f, d, num_heads):
        return layer(x_local, w_in, b_in, w_out, b_out, num_heads, key)
    raise ValueError("the partitioned path uses the tcgen05 family: embed_dim 64, head_dim 8, 16 or 32, F <= 128")
