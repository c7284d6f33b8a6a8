"""bf16 tensor-core mode (tcgen05 kernels) against the reference goldens and the numpy oracle.
Tolerance 2e-2 relative (BASELINE.json north_star, "bf16 mode")."""
import numpy as np
import pytest
import torch

from conftest import assert_close, load_golden, rel_err
from test_gpu_parity import GRAD_KEYS, _make_conv, _run

pytestmark = pytest.mark.gpu
TOL_BF16 = 2e-2


def _status(conv):
    from ampnet_b200 import functional as F_
    return F_.bf16_status(conv._holder["saved"])


@pytest.mark.parametrize("name", ["c4_tokens"])
def test_bf16_mode_matches_reference_golden(name):
    dev = torch.device("cuda:0")
    g = load_golden(name)
    conv = _make_conv(g["d"], g["h"], g["params"], dev, mode="bf16")
    res = _run(conv, g["x"], g["edge_index"], g["d_out"], dev)
    assert _status(conv) == 0
    assert_close(res["out"], g["out"], TOL_BF16)
    assert_close(res["d_x"], g["d_x"], TOL_BF16)
    for k in GRAD_KEYS:
        assert_close(res[k], g[k], TOL_BF16, k)
    deg = np.bincount(g["edge_index"][1], minlength=g["n"])
    assert np.all(res["out"][deg == 0] == 0.0)
    we = g["weight_edges"]
    assert np.abs(conv.attn_output_weights.cpu().numpy()[we] - g["attn_output_weights"]).max() < 2e-2


@pytest.mark.parametrize("shape", [
    dict(n=1500, e=9000, f=128, d=64, h=4, graph="skewed"),      # C4 token shape, many nodes per CTA, hubs + isolated
    dict(n=700, e=4000, f=100, d=64, h=4, graph="uniform"),      # ragged token count (pad + mask)
    dict(n=500, e=3000, f=20, d=64, h=2, graph="skewed"),        # head_dim 32, small F
    dict(n=300, e=300, f=1, d=64, h=4, graph="uniform"),         # single token per node
    dict(n=3, e=700, f=64, d=64, h=4, graph="uniform"),          # fewer nodes than SMs, long edge lists
])
def test_bf16_mode_matches_numpy_oracle(shape):
    from oracle import cases, numpy_oracle
    dev = torch.device("cuda:0")
    x, ei, p, d_out = cases.make_inputs(shape["n"], shape["e"], shape["f"], shape["d"], shape["h"],
                                        graph=shape["graph"], seed=77)
    conv = _make_conv(shape["d"], shape["h"], p, dev, mode="bf16")
    res = _run(conv, x, ei, d_out, dev)
    assert _status(conv) == 0
    ref = numpy_oracle.backward(x, ei, p["in_proj_weight"], p["in_proj_bias"], p["out_proj_weight"],
                                p["out_proj_bias"], shape["h"], d_out)
    assert_close(res["out"], ref["out"], TOL_BF16)
    assert_close(res["d_x"], ref["d_x"], TOL_BF16)
    for k in GRAD_KEYS:
        assert_close(res[k], ref[k], TOL_BF16, k)


def test_bf16_mode_rejects_unsupported_shapes_and_auto_falls_back_to_fp32_kernels():
    from ampnet_b200 import AMPConv
    dev = torch.device("cuda:0")
    x = torch.randn(10, 5 * 12, device=dev)
    ei = torch.randint(0, 10, (2, 30), device=dev)
    with pytest.raises(ValueError):
        AMPConv(12, 3, mode="bf16").to(dev)(x, ei)
    out = AMPConv(12, 3, mode="auto").to(dev)(x, ei)
    assert out.shape == x.shape
