"""Host logic of the per-phase views (ampnet_b200/distributed.py: PhaseGraphs) on the CPU.  The CSR builder itself is a CUDA
kernel (ampconv_graph_build_bipartite, checked bit-exact against numpy on the GPU); here a numpy stand-in with the same
contract replaces it, so that what is checked is the bookkeeping on top: the phases partition a rank's edges by source owner,
the forward's lse2 offsets tile the edge set, the coarse halo view lists every halo edge once and ``halo_lse_map`` points each
of its slots at the forward block of the SAME edge, and the per-owner source lists cover exactly the owner's compact-id range."""
import numpy as np
import pytest
import torch

from ampnet_b200 import distributed as D
from ampnet_b200.loader import make_graph


class _NumpyBipartite:
    """Contract of ampconv_graph_build_bipartite: stable sort by destination, then by source of the destination-sorted slots."""

    def __init__(self, edge_index, num_dst, num_src):
        ei = edge_index.cpu().numpy()
        src, dst = ei[0], ei[1]
        e = src.shape[0]
        self.num_edges, self.num_nodes, self.num_src = e, int(num_dst), int(num_src)
        order = np.argsort(dst, kind="stable")
        i32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32))
        self.dst_eid, self.dst_src = i32(order), i32(src[order])
        deg = np.bincount(dst, minlength=num_dst)
        self.dst_rowptr = i32(np.concatenate([[0], np.cumsum(deg)]))
        sorder = np.argsort(src[order], kind="stable")
        self.src_pos, self.src_dst = i32(sorder), i32(dst[order][sorder])
        self.src_rowptr = i32(np.concatenate([[0], np.cumsum(np.bincount(src, minlength=num_src))]))
        self.inv_deg = torch.from_numpy((1.0 / np.maximum(deg, 1)).astype(np.float32))
        self.has_in = torch.from_numpy((deg > 0).astype(np.float32))
        self.order_dst = torch.argsort(self.dst_rowptr[1:] - self.dst_rowptr[:-1], descending=True).to(torch.int32)
        self.order_src = torch.argsort(self.src_rowptr[1:] - self.src_rowptr[:-1], descending=True).to(torch.int32)


@pytest.mark.parametrize("world,n,e,graph", [(3, 60, 500, "uniform"), (4, 80, 900, "skewed"), (8, 50, 600, "skewed")])
def test_phase_views_and_coarse_halo_map(monkeypatch, world, n, e, graph):
    monkeypatch.setattr(D, "BipartiteGraph", _NumpyBipartite)
    ei = torch.from_numpy(make_graph(graph, n, e, seed=21))
    pgs = D.build_plans_local([D.PartitionedGraph(ei, n, world, r) for r in range(world)])
    total_edges = 0
    for pg in pgs:
        plan = pg.phase_plan
        views = D.PhaseGraphs(pg, plan)
        lei = pg.local_edge_index
        e_local = lei.shape[1]
        total_edges += e_local
        # phases partition the rank's edges; the forward's lse2 blocks tile [0, e_local)
        assert sum(g.num_edges for g in views.graphs) == e_local and views.lse_off[-1] == e_local
        seen = torch.zeros(e_local, dtype=torch.int64)
        for t, (g, sel) in enumerate(zip(views.graphs, views.edge_sel)):
            seen[sel] += 1
            lo, hi = plan.src_range[t]
            if g.num_edges:
                assert int(g.dst_src.min()) >= lo and int(g.dst_src.max()) < hi
            # the phase's source work list is exactly its compact-id range
            assert sorted(views.order_src[t].tolist()) == list(range(lo, hi))
            assert views.n_dst_active[t] == int(((g.dst_rowptr[1:] - g.dst_rowptr[:-1]) > 0).sum())
        assert torch.all(seen == 1)
        # full in-degree in the mean, whatever phase an edge is in
        deg = torch.bincount(lei[1], minlength=pg.n_local)
        assert torch.allclose(views.inv_deg, 1.0 / deg.clamp(min=1).float())
        assert torch.equal(views.has_in, (deg > 0).float())
        if world <= 2:
            assert views.halo is None
            continue
        gh = views.halo
        halo_edges = torch.nonzero(lei[0] >= pg.n_local).squeeze(1)
        assert gh.num_edges == halo_edges.numel()
        if gh.num_edges == 0:
            continue
        # forward index of every local edge, brute force: phase t, slot p holds edge edge_sel[t][dst_eid[p]]
        fwd_of_edge = torch.full((e_local,), -1, dtype=torch.int64)
        for t, (g, sel) in enumerate(zip(views.graphs, views.edge_sel)):
            for p in range(g.num_edges):
                fwd_of_edge[sel[int(g.dst_eid[p])]] = views.lse_off[t] + p
        assert int(fwd_of_edge.min()) >= 0
        for p in range(gh.num_edges):
            edge = int(halo_edges[int(gh.dst_eid[p])])
            assert int(views.halo_lse_map[p]) == int(fwd_of_edge[edge])
            assert int(gh.dst_src[p]) == int(lei[0, edge])
        for t in range(1, world):
            lo, hi = plan.src_range[t]
            assert sorted(views.halo_order_src[t].tolist()) == list(range(lo, hi))
    assert total_edges == e
