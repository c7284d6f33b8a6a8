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


def rms_err(a, b):
    """Per-tensor RMS-relative error ||a - b||_2 / ||b||_2 (small-magnitude tensors cannot hide behind a large maximum)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.sqrt(np.mean((a - b) ** 2)) / max(float(np.sqrt(np.mean(b ** 2))), 1e-30))


def row_err(a, b):
    """Worst row of a 2-D tensor: RMS error of the row relative to max(RMS of the row, 10 % of the tensor's RMS) -- low-degree
    output rows and small gradient rows are held to the same relative bar as the large ones (mixed abs / rel)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.ndim < 2:
        return rms_err(a, b)
    a = a.reshape(a.shape[0], -1)
    b = b.reshape(b.shape[0], -1)
    floor = 0.1 * max(float(np.sqrt(np.mean(b ** 2))), 1e-30)
    denom = np.maximum(np.sqrt(np.mean(b ** 2, axis=1)), floor)
    return float((np.sqrt(np.mean((a - b) ** 2, axis=1)) / denom).max())


def assert_close(a, b, tol, what=""):
    """The parity bar in three readings: global max-relative, per-tensor RMS-relative, worst row (mixed abs / rel)."""
    errs = {"max_rel": rel_err(a, b), "rms_rel": rms_err(a, b), "row_rel": row_err(a, b)}
    assert all(v < tol for v in errs.values()), (what, errs, tol)
    return errs


@pytest.fixture(params=golden_names())
def golden(request):
    return request.param, load_golden(request.param)
