"""Seeded synthetic inputs shared by the golden generator, the tests and bench.py.
TEST/BENCH INFRASTRUCTURE (numpy only; no reference code involved).

Shapes follow SURVEY.md section 8 (C1..C5); see BASELINE.json ``configs``.
"""
import numpy as np

# name -> (N, E, F, d, H, graph)
GOLDEN_CASES = {
    "tiny_generic": dict(n=50, e=400, f=7, d=12, h=3, graph="uniform_with_isolated"),
    "xor_c3": dict(n=40, e=240, f=2, d=3, h=1, graph="knn_self"),
    "c2_shape": dict(n=40, e=200, f=20, d=128, h=4, graph="uniform"),
    "c4_tokens": dict(n=12, e=36, f=128, d=64, h=4, graph="skewed"),
    "c5_tokens": dict(n=10, e=32, f=100, d=64, h=8, graph="skewed"),
    "edge_cases": dict(n=9, e=14, f=3, d=8, h=2, graph="edge_cases"),
    "head_dim_one": dict(n=10, e=30, f=5, d=4, h=4, graph="uniform"),
    "c1_tokens": dict(n=6, e=10, f=200, d=12, h=3, graph="uniform"),
}

FULL_SHAPES = {
    # SURVEY.md section 8 table
    "C2": dict(n=750, e=3000, f=20, d=128, h=4),
    "C3": dict(n=400, e=8400, f=2, d=3, h=1),
    "C4": dict(n=169343, e=1166243, f=128, d=64, h=4),
    "C5": dict(n=2449029, e=61859140, f=100, d=64, h=8),
}


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
