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


from ampnet_b200.loader.synthetic_graph import make_graph, make_inputs  # noqa: E402,F401  (generators live in the package)
