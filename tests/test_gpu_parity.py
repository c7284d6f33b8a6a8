"""Parity of the CUDA path (through the nn.Module mirror -> C ABI) with the reference's outputs
(tests/golden, produced by the reference's own amp_conv.py) and with the numpy oracle.
Strict fp32 mode: 1e-4 relative (BASELINE.json north_star); attention coefficients abs 1e-5."""
import numpy as np
import pytest
import torch

from conftest import assert_close, golden_names, load_golden, rel_err

pytestmark = pytest.mark.gpu

GRAD_KEYS = ["d_in_proj_weight", "d_in_proj_bias", "d_out_proj_weight", "d_out_proj_bias"]
TOL_STRICT = 1e-4


def _make_conv(d, h, params, dev, mode="fp32"):
    from ampnet_b200 import AMPConv
    conv = AMPConv(d, h, mode=mode).to(dev)
    mha = conv.multi_head_attention
    with torch.no_grad():
        mha.in_proj_weight.copy_(torch.from_numpy(np.asarray(params["in_proj_weight"], np.float32)))
        mha.in_proj_bias.copy_(torch.from_numpy(np.asarray(params["in_proj_bias"], np.float32)))
        mha.out_proj.weight.copy_(torch.from_numpy(np.asarray(params["out_proj_weight"], np.float32)))
        mha.out_proj.bias.copy_(torch.from_numpy(np.asarray(params["out_proj_bias"], np.float32)))
    return conv


def _run(conv, x, edge_index, d_out, dev):
    xt = torch.from_numpy(np.asarray(x, np.float32)).to(dev).requires_grad_(True)
    ei = torch.from_numpy(edge_index).to(dev)
    out = conv(xt, ei)
    (out * torch.from_numpy(np.asarray(d_out, np.float32)).to(dev)).sum().backward()
    torch.cuda.synchronize()
    mha = conv.multi_head_attention
    return {
        "out": out.detach().cpu().numpy(), "d_x": xt.grad.cpu().numpy(),
        "d_in_proj_weight": mha.in_proj_weight.grad.cpu().numpy(),
        "d_in_proj_bias": mha.in_proj_bias.grad.cpu().numpy(),
        "d_out_proj_weight": mha.out_proj.weight.grad.cpu().numpy(),
        "d_out_proj_bias": mha.out_proj.bias.grad.cpu().numpy(),
    }


@pytest.mark.parametrize("name", golden_names())
def test_strict_mode_matches_reference_goldens(name):
    dev = torch.device("cuda:0")
    g = load_golden(name)
    conv = _make_conv(g["d"], g["h"], g["params"], dev)
    res = _run(conv, g["x"], g["edge_index"], g["d_out"], dev)
    assert_close(res["out"], g["out"], TOL_STRICT)
    assert_close(res["d_x"], g["d_x"], TOL_STRICT)
    for k in GRAD_KEYS:
        assert_close(res[k], g[k], TOL_STRICT, k)
    # exact zeros for nodes nobody sends to
    deg = np.bincount(g["edge_index"][1], minlength=g["n"])
    assert np.all(res["out"][deg == 0] == 0.0)
    # side outputs, original edge order
    we = g["weight_edges"]
    w = conv.attn_output_weights
    assert tuple(w.shape) == (g["e"], g["f"], g["f"])
    assert np.abs(w.cpu().numpy()[we] - g["attn_output_weights"]).max() < 1e-5
    assert np.allclose(w.sum(-1).cpu().numpy(), 1.0, atol=1e-4)
    ao = conv.attn_output
    assert tuple(ao.shape) == (g["e"], g["f"], g["d"])
    assert_close(ao.cpu().numpy()[we], g["attn_output"], TOL_STRICT)


def test_graph_build_matches_numpy():
    from ampnet_b200.graph import Graph
    from oracle import cases
    dev = torch.device("cuda:0")
    n, e = 1000, 20000
    ei = cases.make_graph("skewed", n, e, seed=3)
    g = Graph(torch.from_numpy(ei).to(dev), n)
    src, dst = ei
    order = np.argsort(dst, kind="stable")
    assert np.array_equal(g.dst_eid.cpu().numpy(), order)
    assert np.array_equal(g.dst_src.cpu().numpy(), src[order])
    deg = np.bincount(dst, minlength=n)
    assert np.array_equal(g.dst_rowptr.cpu().numpy(), np.concatenate([[0], np.cumsum(deg)]))
    order2 = np.argsort(src[order], kind="stable")
    assert np.array_equal(g.src_pos.cpu().numpy(), order2)
    assert np.array_equal(g.src_dst.cpu().numpy(), dst[order][order2])
    assert np.array_equal(g.src_rowptr.cpu().numpy(), np.concatenate([[0], np.cumsum(np.bincount(src, minlength=n))]))
    assert np.allclose(g.inv_deg.cpu().numpy(), 1.0 / np.maximum(deg, 1))
    assert np.array_equal(g.has_in.cpu().numpy(), (deg > 0).astype(np.float32))


def test_out_of_range_edge_index_is_reported():
    from ampnet_b200 import _lib
    from ampnet_b200.graph import Graph
    dev = torch.device("cuda:0")
    ei = torch.tensor([[0, 1, 7], [1, 2, 0]], device=dev)
    with pytest.raises(_lib.AmpConvError) as err:
        Graph(ei, 5)
    assert err.value.status == -3


def test_empty_edge_set_and_isolated_graph():
    dev = torch.device("cuda:0")
    from ampnet_b200 import AMPConv
    conv = AMPConv(8, 2).to(dev)
    x = torch.randn(6, 3 * 8, device=dev, requires_grad=True)
    out = conv(x, torch.zeros(2, 0, dtype=torch.long, device=dev))
    out.sum().backward()
    torch.cuda.synchronize()
    assert torch.equal(out, torch.zeros_like(out))
    assert torch.equal(x.grad, torch.zeros_like(x))
    assert tuple(conv.attn_output_weights.shape) == (0, 3, 3)


@pytest.mark.parametrize("shape", [dict(n=300, e=2500, f=20, d=128, h=4), dict(n=400, e=8400, f=2, d=3, h=1),
                                   dict(n=200, e=1500, f=33, d=24, h=2), dict(n=64, e=500, f=9, d=40, h=1)])
def test_strict_mode_matches_numpy_oracle_on_seeded_inputs(shape):
    """Larger seeded cases than the fixtures (C2 / C3 shapes, odd F, head_dim 12 and 40 -> generic kernel)."""
    from oracle import cases, numpy_oracle
    dev = torch.device("cuda:0")
    x, ei, p, d_out = cases.make_inputs(shape["n"], shape["e"], shape["f"], shape["d"], shape["h"],
                                        graph="skewed", seed=2024)
    conv = _make_conv(shape["d"], shape["h"], p, dev)
    res = _run(conv, x, ei, d_out, dev)
    ref = numpy_oracle.backward(x, ei, p["in_proj_weight"], p["in_proj_bias"], p["out_proj_weight"],
                                p["out_proj_bias"], shape["h"], d_out)
    assert_close(res["out"], ref["out"], TOL_STRICT)
    assert_close(res["d_x"], ref["d_x"], TOL_STRICT)
    for k in GRAD_KEYS:
        assert_close(res[k], ref[k], TOL_STRICT, k)


def test_edge_permutation_invariance_and_linearity_in_upstream_gradient():
    """Size-independent properties: permuting edge_index columns leaves out unchanged (up to fp
    reassociation) and permutes attn_output_weights; the backward is linear in d_out."""
    from oracle import cases
    dev = torch.device("cuda:0")
    n, e, f, d, h = 500, 6000, 16, 32, 4
    x, ei, p, d_out = cases.make_inputs(n, e, f, d, h, graph="skewed", seed=5)
    conv = _make_conv(d, h, p, dev)
    a = _run(conv, x, ei, d_out, dev)
    w_a = conv.attn_output_weights.cpu().numpy()
    perm = np.random.default_rng(0).permutation(e)
    conv.zero_grad()
    b = _run(conv, x, ei[:, perm], d_out, dev)
    w_b = conv.attn_output_weights.cpu().numpy()
    assert_close(b["out"], a["out"], 1e-5)
    assert_close(b["d_x"], a["d_x"], 1e-5)
    assert np.abs(w_b - w_a[perm]).max() < 1e-6
    conv.zero_grad()
    c = _run(conv, x, ei, 2.0 * d_out, dev)
    assert_close(c["d_x"], 2.0 * a["d_x"], 1e-5)
