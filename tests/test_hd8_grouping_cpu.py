"""Algebra of the head-group decomposition that serves head_dim 8 (embed 64, 8 heads: config C5) with the head_dim-16
tensor-core kernels (ampnet_b200/functional.py: hd8_group_params, hd8_compact, _forward_bf16_hd8, _backward_bf16_hd8).

No kernels here: the two group launches are emulated in float64 torch (four zero-padded heads of width 16, the kernels'
conventions: d_agg already divided by the in-degree, dQ scaled by 1/sqrt(16)), the product's own padding / compaction
helpers do the rest, and the result must equal the numpy oracle of the 8-head layer (pinned to the reference)."""
import numpy as np
import torch

from ampnet_b200.functional import hd8_compact, hd8_group_params
from oracle import cases, numpy_oracle


def _group_attention_sum(q, k, v, ei):
    """sum over in-edges of softmax(q_t k_s^T / sqrt(8)) v_s for four padded heads of width 16; q, k, v [n, f, 64]."""
    n, f, _ = q.shape
    out = torch.zeros_like(q)
    for s, t in ei.T.tolist():
        for hh in range(4):
            c = slice(16 * hh, 16 * hh + 16)
            p = torch.softmax(q[t, :, c] @ k[s, :, c].T / 8 ** 0.5, dim=-1)
            out[t, :, c] = out[t, :, c] + p @ v[s, :, c]
    return out


def test_head_group_decomposition_reproduces_the_eight_head_layer():
    n, e, f, d, h = 6, 14, 5, 64, 8
    x, ei, p, d_out = cases.make_inputs(n, e, f, d, h, graph="skewed", seed=3)
    ref = numpy_oracle.backward(x, ei, p["in_proj_weight"], p["in_proj_bias"], p["out_proj_weight"], p["out_proj_bias"], h, d_out)
    t = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float64))
    tokens = t(x).reshape(n * f, d)
    w_in, b_in, w_out, b_out = (t(p[k]) for k in ("in_proj_weight", "in_proj_bias", "out_proj_weight", "out_proj_bias"))
    deg = torch.bincount(torch.from_numpy(ei[1]), minlength=n).double()
    inv_deg = torch.where(deg > 0, 1.0 / deg.clamp(min=1), torch.zeros_like(deg)).repeat_interleave(f)[:, None]
    has_in = (deg > 0).double().repeat_interleave(f)[:, None]

    # ---- forward: two padded 4-head problems, compacted
    groups, aggs = [], []
    for g in range(2):
        w_in_g, b_in_g, _ = hd8_group_params(w_in, b_in, w_out, g)
        assert w_in_g.shape == (192, 64) and float(w_in_g.abs().sum()) > 0
        qkv = (tokens @ w_in_g.T + b_in_g).reshape(n, f, 3, 64)
        q, k, v = (qkv[:, :, i].clone().requires_grad_(True) for i in range(3))
        s = _group_attention_sum(q, k, v, ei)
        groups.append((q, k, v, s))
        aggs.append((s.reshape(n * f, 64) * inv_deg).detach())
    agg = hd8_compact(aggs, 1)
    out = agg @ w_out.T + b_out * has_in
    assert np.abs(out.numpy().reshape(n, f * d) - ref["out"]).max() < 1e-10

    # ---- backward with the kernels' conventions
    d_o = t(d_out).reshape(n * f, d)
    d_w_out = (d_o * has_in).T @ agg
    d_qkv_groups = []
    for g, (q, k, v, s) in enumerate(groups):
        _, _, w_out_g = hd8_group_params(w_in, b_in, w_out, g)
        d_agg_g = (d_o @ w_out_g) * inv_deg                          # ampconv_out_proj_bwd_input_tc with the padded columns
        dq, dk, dv = torch.autograd.grad(s, (q, k, v), grad_outputs=d_agg_g.reshape(n, f, 64))
        d_qkv_groups.append(torch.cat([dq / 2 ** 0.5, dk, dv], dim=2).reshape(n * f, 192))   # dQ kernel: 1/sqrt(16)
    d_qkv = hd8_compact(d_qkv_groups, 3)
    d_qkv[:, :d] *= 2.0 ** 0.5
    d_x = (d_qkv @ w_in).reshape(n, f * d)
    d_w_in = d_qkv.T @ tokens
    d_b_in = d_qkv.sum(0)
    assert np.abs(d_x.numpy() - ref["d_x"]).max() < 1e-10
    assert np.abs(d_w_in.numpy() - ref["d_in_proj_weight"]).max() < 1e-9
    assert np.abs(d_b_in.numpy() - ref["d_in_proj_bias"]).max() < 1e-9
    assert np.abs(d_w_out.numpy() - ref["d_out_proj_weight"]).max() < 1e-9


def test_padded_columns_stay_zero():
    w_in = torch.randn(192, 64, dtype=torch.float64)
    b_in = torch.randn(192, dtype=torch.float64)
    w_out = torch.randn(64, 64, dtype=torch.float64)
    for g in range(2):
        w_in_g, b_in_g, w_out_g = hd8_group_params(w_in, b_in, w_out, g)
        pad = torch.arange(64).view(4, 16)[:, 8:].reshape(-1)
        for s in range(3):
            assert float(w_in_g[64 * s + pad].abs().max()) == 0.0 and float(b_in_g[64 * s + pad].abs().max()) == 0.0
        assert float(w_out_g[:, pad].abs().max()) == 0.0
        real = torch.arange(64).view(4, 16)[:, :8].reshape(-1)
        assert torch.equal(w_out_g[:, real], w_out[:, 32 * g:32 * g + 32])
        assert torch.equal(w_in_g[128 + real], w_in[128 + 32 * g:128 + 32 * g + 32])


def _torch_layer(x, ei, w_in, b_in, w_out, b_out, h):
    """Differentiable float64 restatement of the layer (amp_conv.py:24-51 with stock MHA arithmetic), any head count."""
    n = x.shape[0]
    d = w_in.shape[1]
    f = x.shape[1] // d
    hd = d // h
    e = ei.shape[1]
    qkv = x.view(n, f, d) @ w_in.T + b_in
    q, k, v = qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:]
    qe = q[ei[1]].view(e, f, h, hd).transpose(1, 2)
    ke = k[ei[0]].view(e, f, h, hd).transpose(1, 2)
    ve = v[ei[0]].view(e, f, h, hd).transpose(1, 2)
    p = torch.softmax(qe @ ke.transpose(-1, -2) / hd ** 0.5, dim=-1)
    o = (p @ ve).transpose(1, 2).reshape(e, f, d) @ w_out.T + b_out
    deg = torch.bincount(ei[1], minlength=n).clamp(min=1).double()
    return (torch.zeros(n, f, d, dtype=x.dtype).index_add_(0, ei[1], o) / deg[:, None, None]).reshape(n, f * d)


def test_autograd_level_composition_equals_the_eight_head_layer():
    """functional.hd8_compose (used by the partitioned path): out, dX and all four parameter gradients."""
    from ampnet_b200.functional import hd8_compose
    n, e, f, d, h = 7, 20, 4, 64, 8
    x, ei, p, d_out = cases.make_inputs(n, e, f, d, h, graph="skewed", seed=4)
    ref = numpy_oracle.backward(x, ei, p["in_proj_weight"], p["in_proj_bias"], p["out_proj_weight"], p["out_proj_bias"], h, d_out)
    t = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float64)).requires_grad_(True)
    xt, w_in, b_in, w_out, b_out = t(x), t(p["in_proj_weight"]), t(p["in_proj_bias"]), t(p["out_proj_weight"]), t(p["out_proj_bias"])
    eit = torch.from_numpy(ei)
    out = hd8_compose(lambda x_, wi, bi, wo, bo, heads: _torch_layer(x_, eit, wi, bi, wo, bo, heads), xt, w_in, b_in, w_out, b_out)
    (out * torch.from_numpy(d_out).double()).sum().backward()
    assert np.abs(out.detach().numpy() - ref["out"]).max() < 1e-10
    assert np.abs(xt.grad.numpy() - ref["d_x"]).max() < 1e-10
    for got, key in ((w_in, "d_in_proj_weight"), (b_in, "d_in_proj_bias"), (w_out, "d_out_proj_weight"), (b_out, "d_out_proj_bias")):
        assert np.abs(got.grad.numpy() - ref[key]).max() < 1e-9, key
