"""ampnet_b200.loader.synthetic_xor against a restatement of the reference generator's graph construction
(synthetic_benchmark/synthetic_xor.py:69-101: sklearn ball tree with k + 1 neighbours, adjacency, row-major edge list)."""
import numpy as np
import pytest
import torch

from ampnet_b200.loader import create_duplicated_xor_data, knn_self_edges


def _reference_graph(x, k):
    from sklearn.neighbors import NearestNeighbors
    nbrs = NearestNeighbors(n_neighbors=k + 1, algorithm="ball_tree", metric="minkowski").fit(x)
    _, indices = nbrs.kneighbors(x)
    n = x.shape[0]
    adj = np.zeros((n, n), dtype=np.uint8)
    for row in range(n):
        for col in range(indices.shape[1]):
            adj[row, indices[row, col]] = 1
    src, dst = [], []
    for row in range(n):
        for col in range(n):
            if adj[row][col] > 0:
                src.append(row)
                dst.append(col)
    return adj, np.array([src, dst])


def test_shapes_labels_and_config_3_edge_count():
    g = torch.Generator().manual_seed(1)
    x, y, adj, ei = create_duplicated_xor_data(400, 0.3, 20, 1, generator=g)      # the call of synthetic_training_modular.py
    assert x.shape == (400, 2) and x.dtype == torch.float32 and y.shape == (400,)
    assert ei.shape == (2, 8400) and ei.dtype == torch.int64                      # 400 * (20 neighbours + self): SURVEY C3
    assert int(adj.sum()) == 8400 and bool((adj.diagonal() == 1).all())
    assert y.tolist() == [0.0] * 100 + [1.0] * 200 + [0.0] * 100
    corner = torch.tensor([[0, 0], [0, 1], [1, 0], [1, 1]], dtype=torch.float32).repeat_interleave(100, dim=0)
    assert float((x - corner).std()) == pytest.approx(0.3, rel=0.1)
    assert bool(((x.round().clamp(0, 1)[:, 0] != x.round().clamp(0, 1)[:, 1]).float() == y).float().mean() > 0.8)
    with pytest.raises(ValueError):
        create_duplicated_xor_data(402)


def test_graph_equals_the_reference_construction_on_the_same_features():
    g = torch.Generator().manual_seed(7)
    x, _, adj, ei = create_duplicated_xor_data(120, 0.3, 10, 3, generator=g)
    ref_adj, ref_ei = _reference_graph(x.double().numpy(), 10)
    # the generator measured distances on the float64 features; compare on exactly those
    g = torch.Generator().manual_seed(7)
    corners = torch.tensor([[0.0, 0.0], [0.0, 1.0], [1.0, 0.0], [1.0, 1.0]], dtype=torch.float64)
    x64 = corners.repeat_interleave(30, dim=0).repeat(1, 3) + 0.3 * torch.randn((120, 6), dtype=torch.float64, generator=g)
    ref_adj, ref_ei = _reference_graph(x64.numpy(), 10)
    assert np.array_equal(adj.numpy(), ref_adj)
    assert np.array_equal(ei.numpy(), ref_ei)
    adj2, ei2 = knn_self_edges(x64, 10)
    assert torch.equal(adj2, adj) and torch.equal(ei2, ei)
