"""Host logic of the destination-partitioned (N > 1) path on CPU with gloo, world_size 2 and 3: partition balance,
compact-id remapping, the halo exchange of K/V rows and the return of the source-side gradient rows to their owners.
The CUDA kernels cannot run here; plain torch index ops stand in for them (this tests the plumbing, not the math)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ampnet_b200 import distributed as D
from oracle import cases


def test_partition_ranges_balance_in_edges():
    ei = torch.from_numpy(cases.make_graph("skewed", 1000, 20000, seed=5))
    deg = torch.bincount(ei[1], minlength=1000)
    for world in (1, 2, 4, 8):
        b = D.partition_ranges(deg, world)
        assert b[0] == 0 and b[-1] == 1000 and torch.all(b[1:] >= b[:-1])
        loads = torch.stack([deg[b[r]:b[r + 1]].sum() for r in range(world)]).float()
        # contiguous ranges cannot split a hub: allow one max-degree node of slack
        assert float(loads.max()) <= 20000 / world + float(deg.max())


def test_local_edges_cover_the_global_edge_set_and_compact_ids_invert():
    n, e, world = 300, 4000, 4
    ei = torch.from_numpy(cases.make_graph("uniform", n, e, seed=9))
    seen = torch.zeros(e, dtype=torch.int64)
    for r in range(world):
        pg = D.PartitionedGraph(ei, n, world, r)
        seen[pg.edge_ids] += 1
        src_c, dst_loc = pg.local_edge_index
        assert torch.equal(pg.global_id(src_c), ei[0, pg.edge_ids])
        assert torch.equal(dst_loc + pg.lo, ei[1, pg.edge_ids])
        assert int(dst_loc.max()) < pg.n_local and int(src_c.max()) < pg.num_kv_nodes
        # halo = exactly the remote sources of the local edges, grouped by owner, none owned by this rank
        remote = ei[0, pg.edge_ids]
        remote = remote[(remote < pg.lo) | (remote >= pg.hi)]
        assert torch.equal(pg.halo_ids, torch.unique(remote))
        assert pg.recv_counts[r] == 0 and sum(pg.recv_counts) == pg.n_halo
    assert torch.all(seen == 1)


def _worker(rank, world, port, n, e, c, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ei = torch.from_numpy(cases.make_graph("skewed", n, e, seed=11))
        g = torch.Generator().manual_seed(0)
        feat = torch.randn(n, c, generator=g)          # stands for the projected K (or V) of every node
        grad_msg = torch.randn(e, c, generator=g)      # stands for the per-edge source-side gradient
        pg = D.PartitionedGraph(ei, n, world, rank).build_plan()
        assert sum(pg.send_counts) == pg.send_idx.numel() and pg.send_counts[rank] == 0
        # forward exchange: own rows + halo rows, sources looked up by compact id
        rows = torch.empty(pg.num_kv_nodes, c)
        rows[:pg.n_local] = feat[pg.lo:pg.hi]
        D.halo_gather(rows[:pg.n_local], rows[pg.n_local:], pg)
        assert torch.equal(rows[pg.local_edge_index[0]], feat[ei[0, pg.edge_ids]])
        # backward exchange: partial scatter-add over compact sources, halo rows added at their owners
        partial = torch.zeros(pg.num_kv_nodes, c).index_add_(0, pg.local_edge_index[0], grad_msg[pg.edge_ids])
        mine = partial[:pg.n_local].clone()
        D.halo_scatter_add(partial[pg.n_local:], mine, pg)
        ref = torch.zeros(n, c).index_add_(0, ei[0], grad_msg)[pg.lo:pg.hi]
        ret[rank] = float((mine - ref).abs().max())
    finally:
        dist.destroy_process_group()


def test_exchange_steps_world2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, 200, 3000, 8, ret), nprocs=world, join=True)
    assert len(ret) == world and all(v < 1e-4 for v in ret.values()), dict(ret)


def test_exchange_steps_world3_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    world = 3
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, 150, 1200, 4, ret), nprocs=world, join=True)
    assert len(ret) == world and all(v < 1e-4 for v in ret.values()), dict(ret)
