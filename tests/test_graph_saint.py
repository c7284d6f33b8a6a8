"""GraphSAINT random-walk sampler (ampnet_b200/loader/graph_saint.py) on CPU tensors against brute-force restatements of
the semantics the reference vendors (visualization/visualize_graphsaint_subgraphs.py:107-173, 195-199)."""
import numpy as np
import torch

from ampnet_b200.loader import GraphSAINTRandomWalkSampler, SubgraphData, cora_shaped_data


def _toy(n=60, e=300, seed=0):
    g = torch.Generator().manual_seed(seed)
    ei = torch.randint(0, n, (2, e), generator=g)
    ei[0][ei[0] == 7] = 8                       # node 7 has no out-edges
    return SubgraphData(x=torch.randn(n, 5, generator=g), y=torch.randint(0, 3, (n,), generator=g), edge_index=ei,
                        edge_attr=torch.arange(e).float(), num_nodes=n)


def test_walks_follow_out_edges_and_isolated_nodes_stay():
    data = _toy()
    s = GraphSAINTRandomWalkSampler(data, batch_size=16, walk_length=12, generator=torch.Generator().manual_seed(1))
    walk = s.sample_nodes().view(16, 13)
    edges = set(map(tuple, data.edge_index.t().tolist()))
    outdeg = torch.bincount(data.edge_index[0], minlength=data.num_nodes)
    assert int(outdeg[7]) == 0
    for w in walk.tolist():
        for a, b in zip(w[:-1], w[1:]):
            assert (a, b) in edges or (a == b and int(outdeg[a]) == 0)
    # every start node is a valid node id and walks have walk_length + 1 entries
    assert walk.shape == (16, 13) and int(walk.min()) >= 0 and int(walk.max()) < data.num_nodes


def test_induced_subgraph_and_collate_match_brute_force():
    data = _toy(seed=3)
    s = GraphSAINTRandomWalkSampler(data, batch_size=6, walk_length=5, num_steps=4, generator=torch.Generator().manual_seed(2))
    for batch in s:
        node_idx = torch.unique(torch.cat([torch.nonzero((data.x == r).all(dim=1)).view(-1) for r in batch.x]))
        assert batch.num_nodes == batch.x.size(0) == batch.y.size(0)
        assert torch.equal(data.x[node_idx], batch.x) and torch.equal(data.y[node_idx], batch.y)
        in_set = torch.zeros(data.num_nodes, dtype=torch.bool)
        in_set[node_idx] = True
        src, dst = data.edge_index
        keep = in_set[src] & in_set[dst]
        # same multiset of (global src, global dst, edge attribute)
        ref = sorted(zip(src[keep].tolist(), dst[keep].tolist(), data.edge_attr[keep].tolist()))
        got = sorted(zip(node_idx[batch.edge_index[0]].tolist(), node_idx[batch.edge_index[1]].tolist(), batch.edge_attr.tolist()))
        assert ref == got


def test_normalisation_coefficients_follow_the_graphsaint_formulas():
    data = _toy(seed=5)
    gen = torch.Generator().manual_seed(9)
    s = GraphSAINTRandomWalkSampler(data, batch_size=5, walk_length=4, num_steps=7, sample_coverage=20, generator=gen)
    assert s.node_norm.shape == (data.num_nodes,) and s.edge_norm.shape == (data.edge_index.size(1),)
    assert bool((s.node_norm > 0).all()) and bool((s.edge_norm >= 0).all()) and float(s.edge_norm.max()) <= 1e4
    # replay the estimation with the same generator state
    gen2 = torch.Generator().manual_seed(9)
    r = GraphSAINTRandomWalkSampler(data, batch_size=5, walk_length=4, num_steps=7, sample_coverage=0, generator=gen2)
    node_count = torch.zeros(data.num_nodes)
    edge_count = torch.zeros(data.edge_index.size(1))
    num_samples = total = 0
    while total < data.num_nodes * 20:
        for _ in range(7):
            node_idx, _, edge_idx = r.sample()
            node_count[node_idx] += 1
            edge_count[edge_idx] += 1
            total += node_idx.numel()
        num_samples += 7
    t = node_count[data.edge_index[0]]
    edge_norm = (t / edge_count).clamp(0, 1e4)
    edge_norm[torch.isnan(edge_norm)] = 0.1
    node_count[node_count == 0] = 0.1
    assert torch.allclose(s.node_norm, num_samples / node_count / data.num_nodes)
    assert torch.allclose(s.edge_norm, edge_norm)
    batch = next(iter(s))
    assert batch.node_norm.shape == (batch.num_nodes,) and batch.edge_norm.shape == (batch.num_edges,)


def test_cora_shaped_data_has_the_planetoid_shape():
    d = cora_shaped_data()
    assert d.x.shape == (2708, 1433) and d.edge_index.shape == (2, 10556) and int(d.y.max()) == 6
    assert int(d.train_mask.sum()) == 140 and int(d.val_mask.sum()) == 500 and int(d.test_mask.sum()) == 1000
    assert bool(((d.x != 0).sum(dim=1) >= 1).all()) and set(np.unique(d.x.numpy())) == {0.0, 1.0}
    src, dst = d.edge_index
    assert torch.equal(src[:5278], dst[5278:]) and torch.equal(dst[:5278], src[5278:])      # symmetrised
