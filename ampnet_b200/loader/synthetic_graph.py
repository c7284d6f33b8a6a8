"""Seeded synthetic graphs and inputs of SURVEY.md section 8(d): the shapes of BASELINE.json's configs with no dataset
on disk (ogbn-arxiv / ogbn-products shapes, the XOR k-NN pattern, the edge cases the tests pin).  numpy only; shared by
bench.py, the tests and the golden generator (oracle/cases.py re-exports these)."""
import numpy as np


def make_graph(kind, n, e, seed=7):
    """Returns int64 edge_index [2, E]; row 0 = source, row 1 = destination."""
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        src = rng.integers(0, n, size=e)
        dst = rng.integers(0, n, size=e)
    elif kind == "uniform_with_isolated":
        # the last fifth of the nodes never receives an edge
        src = rng.integers(0, n, size=e)
        dst = rng.integers(0, max(1, (4 * n) // 5), size=e)
    elif kind == "skewed":
        # Zipf-like in-degree (alpha ~ 2.1) to mimic ogbn in-degree tails; src uniform
        w = 1.0 / np.arange(1, n + 1) ** 1.1
        w = rng.permutation(w / w.sum())
        dst = rng.choice(n, size=e, p=w)
        src = rng.integers(0, n, size=e)
    elif kind == "knn_self":
        # XOR-benchmark style: every node receives from itself and k random "neighbours"
        k = e // n - 1
        dst = np.repeat(np.arange(n), k + 1)
        src = np.concatenate([np.concatenate([[i], rng.choice(n, size=k, replace=False)]) for i in range(n)])
    elif kind == "edge_cases":
        # self loops, duplicate edges, a hub, isolated nodes (5, 7, 8 receive nothing)
        pairs = [(0, 0), (1, 1), (0, 1), (0, 1), (0, 1), (2, 3), (3, 2), (4, 2), (5, 2), (6, 2),
                 (7, 2), (8, 2), (2, 6), (6, 4)]
        assert len(pairs) == e
        src = np.array([p[0] for p in pairs])
        dst = np.array([p[1] for p in pairs])
    else:
        raise ValueError(kind)
    return np.stack([src, dst]).astype(np.int64)


def make_inputs(n, e, f, d, h, graph="uniform", seed=1234, relu_x=False, dtype=np.float32):
    """x ~ N(0,1) (optionally relu'ed, like the input of a second layer), parameters with
    the reference's init scale but *non-zero* biases, upstream gradient ~ N(0,1)."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, f * d)).astype(dtype)
    if relu_x:
        x = np.maximum(x, 0)
    bound_in = np.sqrt(6.0 / (3 * d + d))          # xavier_uniform on [3d, d]
    bound_out = 1.0 / np.sqrt(d)                    # default nn.Linear
    params = {
        "in_proj_weight": rng.uniform(-bound_in, bound_in, size=(3 * d, d)).astype(dtype),
        "in_proj_bias": (0.1 * rng.standard_normal(3 * d)).astype(dtype),
        "out_proj_weight": rng.uniform(-bound_out, bound_out, size=(d, d)).astype(dtype),
        "out_proj_bias": (0.1 * rng.standard_normal(d)).astype(dtype),
    }
    d_out = rng.standard_normal((n, f * d)).astype(dtype)
    edge_index = make_graph(graph, n, e, seed=seed + 7)
    return x, edge_index, params, d_out
