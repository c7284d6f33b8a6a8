"""The synthetic XOR graph of the reference's benchmark (BASELINE config 3), built with tensor ops where the model lives.

Reference: ``synthetic_benchmark/synthetic_xor.py:24-101`` (``create_duplicated_xor_data``), called as
``create_duplicated_xor_data(400, 0.3, 20, 1)`` by ``synthetic_training_modular.py:124-137`` / ``xor_training_utils.py``.
Four equally sized groups of nodes carry the corners (0,0), (0,1), (1,0), (1,1) of the XOR table, repeated
``feature_repeats`` times along the feature axis, plus Gaussian noise; the label is the XOR of the corner; every node is
linked to its ``num_nearest_neighbors`` nearest neighbours in feature space AND to itself (the reference asks sklearn for
``k + 1`` neighbours and keeps column 0, the point itself, ``:69-76``); the edge list is the adjacency matrix read in
row-major order, row = source, column = destination (``:92-101``).

The reference does this with sklearn's ball tree and two Python double loops; here it is one ``cdist`` + ``topk`` and a
``nonzero``.  With the same noisy features the adjacency is identical whenever the k-th and (k+1)-th distances of a node
differ (``tests/test_synthetic_xor.py`` checks against sklearn); the noise itself comes from a ``torch.Generator`` -- bit
parity with ``np.random.normal`` is neither possible nor needed.
"""
import torch

__all__ = ["knn_self_edges", "create_duplicated_xor_data"]


def knn_self_edges(x, num_nearest_neighbors):
    """x [N, C] -> (adjacency [N, N] uint8, edge_index [2, E] int64): row i is linked to itself and to its k nearest other
    rows (Euclidean); edges in row-major order of the adjacency matrix, row = source, column = destination."""
    n = x.shape[0]
    k = min(num_nearest_neighbors + 1, n)
    dist = torch.cdist(x.double(), x.double())
    dist.fill_diagonal_(-1.0)                                   # the point itself is always its own first neighbour
    nbr = torch.topk(dist, k, dim=1, largest=False).indices
    adj = torch.zeros((n, n), dtype=torch.uint8, device=x.device)
    adj.scatter_(1, nbr, 1)
    edge_index = torch.nonzero(adj, as_tuple=False).t().contiguous()
    return adj, edge_index


def create_duplicated_xor_data(num_samples, noise_std=0.1, num_nearest_neighbors=10, feature_repeats=5, generator=None,
                               device=None):
    """-> (x [N, 2*feature_repeats] float32, y [N] float32, adjacency [N, N] uint8, edge_index [2, E] int64), the tuple the
    reference's function returns (``synthetic_xor.py:101``)."""
    if num_samples % 4 != 0:
        raise ValueError("num_samples must be an integer divisible by 4.")       # the reference asserts (:46)
    device = torch.device(device) if device is not None else (generator.device if generator is not None else torch.device("cpu"))
    corners = torch.tensor([[0.0, 0.0], [0.0, 1.0], [1.0, 0.0], [1.0, 1.0]], dtype=torch.float64, device=device)
    labels = torch.tensor([0.0, 1.0, 1.0, 0.0], dtype=torch.float64, device=device)
    rep = num_samples // 4
    x = corners.repeat_interleave(rep, dim=0).repeat(1, feature_repeats)
    y = labels.repeat_interleave(rep)
    x = x + noise_std * torch.randn(x.shape, dtype=torch.float64, device=device, generator=generator)
    adj, edge_index = knn_self_edges(x, num_nearest_neighbors)
    return x.float(), y.float(), adj, edge_index
