"""Ring-phase plan of the partitioned path (ampnet_b200/distributed.py: PhasePlan, phase_of_sources) on CPU with gloo,
world sizes 2, 3 and 4.  The pushes into the peers' windows are emulated with point-to-point messages that carry the
SENDER-computed destination offset -- exactly the number the GPU path hands to ``ampconv_peer_copy`` -- so a wrong
offset, block size or ring order shows up as a wrong row here, without a GPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ampnet_b200 import distributed as D
from ampnet_b200.loader import make_graph


def _exchange(rank, world, sends, recv_from):
    """sends[t] = (peer, offset, block) for t = 1..world-1; returns {t: (offset, block)} received from recv_from[t]."""
    got = {}
    for t in range(1, world):
        peer, off, block = sends[t]
        meta = torch.tensor([off, block.shape[0]], dtype=torch.int64)
        rmeta = torch.empty(2, dtype=torch.int64)
        reqs = [dist.isend(meta, peer, tag=2 * t), dist.irecv(rmeta, recv_from[t], tag=2 * t)]
        for q in reqs:
            q.wait()
        rblock = torch.empty((int(rmeta[1]),) + tuple(block.shape[1:]), dtype=block.dtype)
        reqs = [dist.isend(block.contiguous(), peer, tag=2 * t + 1), dist.irecv(rblock, recv_from[t], tag=2 * t + 1)]
        for q in reqs:
            q.wait()
        got[t] = (int(rmeta[0]), rblock)
    return got


def _worker(rank, world, port, n, e, c, graph, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ei = torch.from_numpy(make_graph(graph, n, e, seed=13))
        gen = torch.Generator().manual_seed(1)
        feat = torch.randn(n, c, generator=gen)
        grad_msg = torch.randn(e, c, generator=gen)
        pg = D.PartitionedGraph(ei, n, world, rank).build_plan()
        plan = pg.phase_plan
        errs = {}
        # --- phases partition the rank's edges by source owner, ranges are disjoint and cover the compact id space
        ph = D.phase_of_sources(pg.local_edge_index[0], plan)
        owner_of_src = torch.searchsorted(pg.bounds, ei[0, pg.edge_ids], right=True) - 1
        ring = torch.tensor(plan.ring)
        assert torch.equal(ring[ph], owner_of_src)
        covered = sorted(plan.src_range)
        assert covered[0][0] == 0 and covered[-1][1] == pg.num_kv_nodes
        assert all(a[1] == b[0] for a, b in zip(covered[:-1], covered[1:]))
        # --- forward: own rows, then pushes around the ring at sender-computed offsets
        k_all = torch.full((pg.num_kv_nodes, c), float("nan"))
        k_all[:pg.n_local] = feat[pg.lo:pg.hi]
        staging = k_all[:pg.n_local][pg.send_idx]
        sends = {t: (plan.fwd_dst[t], plan.fwd_dst_off[t],
                     staging[plan.send_off[plan.fwd_dst[t]]:plan.send_off[plan.fwd_dst[t]] + plan.fwd_rows[t]])
                 for t in range(1, world)}
        got = _exchange(rank, world, sends, {t: plan.ring[t] for t in range(1, world)})
        for t, (off, block) in got.items():
            lo, hi = plan.src_range[t]
            assert off == lo and block.shape[0] == hi - lo, (t, off, lo, hi, block.shape)
            k_all[off:off + block.shape[0]] = block
        errs["fwd"] = float((k_all[pg.local_edge_index[0]] - feat[ei[0, pg.edge_ids]]).abs().max()) if pg.edge_ids.numel() else 0.0
        # --- backward: per-phase blocks of partial sums go to their owners' receive windows, fixed-order add there
        partial = torch.zeros(pg.num_kv_nodes, c).index_add_(0, pg.local_edge_index[0], grad_msg[pg.edge_ids])
        mine = partial[:pg.n_local].clone()
        sends = {t: (plan.ring[t], plan.bwd_dst_off[t], partial[plan.src_range[t][0]:plan.src_range[t][1]])
                 for t in range(1, world)}
        got = _exchange(rank, world, sends, {t: plan.bwd_src[t] for t in range(1, world)})
        recv = torch.full((int(sum(plan.send_counts)), c), float("nan"))
        for t, (off, block) in got.items():
            s = plan.bwd_src[t]
            assert off == plan.send_off[s] and block.shape[0] == plan.send_counts[s]
            recv[off:off + block.shape[0]] = block
        for i in range(pg.add_tgt.numel()):
            for j in range(int(pg.add_rowptr[i]), int(pg.add_rowptr[i + 1])):
                mine[int(pg.add_tgt[i])] += recv[int(pg.add_pos[j])]
        ref = torch.zeros(n, c).index_add_(0, ei[0], grad_msg)[pg.lo:pg.hi]
        errs["bwd"] = float((mine - ref).abs().max()) if mine.numel() else 0.0
        ret[rank] = errs
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,e,graph", [(2, 120, 1500, "skewed"), (3, 90, 700, "uniform"), (4, 64, 300, "skewed"),
                                             (6, 40, 400, "skewed")])      # a hub: some ranks own no node at all
def test_ring_phase_plan_moves_every_row_to_the_right_place(world, n, e, graph):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, n, e, 4, graph, ret), nprocs=world, join=True)
    assert len(ret) == world
    for rank, errs in ret.items():
        assert errs["fwd"] == 0.0 and errs["bwd"] < 1e-4, (rank, errs)


def test_phase_plan_offsets_by_hand():
    # 3 ranks with 4, 5, 6 local nodes; recv_matrix[p][o]: rows rank p needs from owner o
    rm = [[0, 2, 3], [1, 0, 4], [2, 5, 0]]
    plan = D.PhasePlan(1, 3, [4, 5, 6], rm)
    assert plan.ring == [1, 2, 0]
    assert plan.src_range == [(0, 5), (5 + 1, 5 + 5), (5, 5 + 1)]      # halo blocks by owner: owner 0 (1 row), owner 2 (4 rows)
    assert plan.send_counts == [2, 0, 5] and plan.send_off == [0, 2, 2, 7]
    # forward: phase 1 sends to rank 0 (which consumes owner 1 in ITS phase 1) behind its 4 own nodes; phase 2 to rank 2
    assert plan.fwd_dst[1:] == [0, 2] and plan.fwd_dst_off[1:] == [4 + 0, 6 + 2] and plan.fwd_rows[1:] == [2, 5]
    # backward: phase 1 returns owner 2's 4 rows behind rank 0's block of 3, phase 2 returns owner 0's row at its sender slot
    assert plan.bwd_rows[1:] == [4, 1] and plan.bwd_dst_off[1:] == [3, 0] and plan.bwd_src[1:] == [0, 2]
