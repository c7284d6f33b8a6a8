import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_names():
    """Per-layer (AMPConv) goldens; the model-level `ampgcn_*` / `ampnetclf_*` files have their own schema and tests."""
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR)
                  if f.endswith(".npz") and not f.startswith(("ampgcn_", "ampnetclf_")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: z[k] for k in z.files}
    for k in ("n", "e", "f", "d", "h"):
        g[k] = int(g[k])
    g["edge_index"] = g["edge_index"].astype(np.int64)
    g["params"] = {k[len("param_"):]: g[k] for k in list(g) if k.startswith("param_")}
    return g


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = max(float(np.abs(b).max()), 1e-30)
    return float(np.abs(a - b).max()) / denom


@pytest.fixture(params=golden_names())
def golden(request):
    return request.param, load_golden(request.param)
