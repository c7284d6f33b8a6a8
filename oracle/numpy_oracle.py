"""Independent fp64 restatement of the AMPConv layer in numpy, forward and backward.
TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

It follows the *node-level* formulation the CUDA path uses, which is algebraically the
reference's per-edge one:

* in-projection once per node instead of once per edge --
  ``q = x_i Wq^T + bq``, ``[k|v] = x_j [Wk;Wv]^T + [bk;bv]``
  (reference era arithmetic ``src/ampnet/conv/custom_multihead_attn_forward.py:4031-4084``;
  the rows of ``in_proj_weight`` are ordered q, k, v);
* head split ``d -> H x hd`` (``:4376-4387``), ``q * hd**-0.5`` (``:4173``),
  ``S = q k^T`` (``:4175``), softmax over the *source* tokens (the line the vendored copy
  comments out at ``:4179-4180`` and stock torch executes), ``O = P v`` (``:4185``);
* ``out_proj`` (``:4436-4437``) commutes with the mean because the mean is linear; its bias
  survives only where the in-degree is non-zero;
* head-averaged weights ``P.mean(heads)`` (``:4441-4442``), edge order = ``edge_index`` order;
* mean aggregation at ``edge_index[1]`` with zero rows for isolated nodes
  (``src/ampnet/conv/amp_conv.py:11``; ``synthetic_benchmark/testing_message_passing_pyg.py:37-40``).

The backward is written out by hand (no autograd) so that it is an independent check of
the reference's autograd result stored in ``tests/golden``.
"""
import numpy as np


def _split(w_in, b_in, d):
    return (w_in[:d], w_in[d:2 * d], w_in[2 * d:]), (b_in[:d], b_in[d:2 * d], b_in[2 * d:])


def _edge_chunks(num_edges, chunk):
    for lo in range(0, num_edges, chunk):
        yield lo, min(num_edges, lo + chunk)


def _softmax_last(s):
    s = s - s.max(axis=-1, keepdims=True)
    p = np.exp(s)
    return p / p.sum(axis=-1, keepdims=True)


def forward(x, edge_index, w_in, b_in, w_out, b_out, num_heads, need_weights=False,
            need_edge_output=False, chunk=2048):
    """Returns dict(out [N,F*d], deg [N], and optionally weights [E,F,F], edge_output [E,F,d])."""
    x = np.asarray(x, dtype=np.float64)
    w_in, b_in = np.asarray(w_in, np.float64), np.asarray(b_in, np.float64)
    w_out, b_out = np.asarray(w_out, np.float64), np.asarray(b_out, np.float64)
    edge_index = np.asarray(edge_index, dtype=np.int64)
    n, width = x.shape
    d = w_in.shape[1]
    if width % d:
        raise ValueError("x.shape[1] must be a multiple of embed_dim")
    if d % num_heads:
        raise ValueError("embed_dim must be divisible by num_heads")
    f, h, hd = width // d, num_heads, d // num_heads
    src, dst = edge_index[0], edge_index[1]
    e = src.shape[0]
    xt = x.reshape(n, f, d)
    (wq, wk, wv), (bq, bk, bv) = _split(w_in, b_in, d)
    q = (xt @ wq.T + bq).reshape(n, f, h, hd)
    k = (xt @ wk.T + bk).reshape(n, f, h, hd)
    v = (xt @ wv.T + bv).reshape(n, f, h, hd)
    scale = hd ** -0.5
    agg = np.zeros((n, f, h, hd))
    weights = np.zeros((e, f, f)) if need_weights else None
    edge_out = np.zeros((e, f, d)) if need_edge_output else None
    for lo, hi in _edge_chunks(e, chunk):
        s = np.einsum("eihc,ejhc->ehij", q[dst[lo:hi]] * scale, k[src[lo:hi]])
        p = _softmax_last(s)
        o = np.einsum("ehij,ejhc->eihc", p, v[src[lo:hi]])
        np.add.at(agg, dst[lo:hi], o)
        if need_weights:
            weights[lo:hi] = p.mean(axis=1)
        if need_edge_output:
            edge_out[lo:hi] = o.reshape(hi - lo, f, d) @ w_out.T + b_out
    deg = np.bincount(dst, minlength=n).astype(np.float64)
    agg = agg.reshape(n, f, d) / np.maximum(deg, 1.0)[:, None, None]
    out = agg @ w_out.T + b_out * (deg > 0)[:, None, None]
    res = {"out": out.reshape(n, width), "deg": deg, "agg": agg, "q": q, "k": k, "v": v}
    if need_weights:
        res["weights"] = weights
    if need_edge_output:
        res["edge_output"] = edge_out
    return res


def backward(x, edge_index, w_in, b_in, w_out, b_out, num_heads, d_out, chunk=2048):
    """Gradients of ``(out * d_out).sum()`` w.r.t. x and the four parameters."""
    x = np.asarray(x, dtype=np.float64)
    d_out = np.asarray(d_out, dtype=np.float64)
    w_in, b_in = np.asarray(w_in, np.float64), np.asarray(b_in, np.float64)
    w_out, b_out = np.asarray(w_out, np.float64), np.asarray(b_out, np.float64)
    edge_index = np.asarray(edge_index, dtype=np.int64)
    fw = forward(x, edge_index, w_in, b_in, w_out, b_out, num_heads, chunk=chunk)
    n, width = x.shape
    d = w_in.shape[1]
    f, h, hd = width // d, num_heads, d // num_heads
    src, dst = edge_index[0], edge_index[1]
    e = src.shape[0]
    q, k, v, deg, agg = fw["q"], fw["k"], fw["v"], fw["deg"], fw["agg"]
    scale = hd ** -0.5
    g = d_out.reshape(n, f, d)
    # out = agg W_o^T + b_o [deg>0]
    d_w_out = np.einsum("nfa,nfb->ab", g, agg)
    d_b_out = (g * (deg > 0)[:, None, None]).sum(axis=(0, 1))
    d_agg = (g @ w_out) / np.maximum(deg, 1.0)[:, None, None]   # = dO of every in-edge
    d_agg = d_agg.reshape(n, f, h, hd)
    dq = np.zeros_like(q)
    dk = np.zeros_like(k)
    dv = np.zeros_like(v)
    for lo, hi in _edge_chunks(e, chunk):
        qe, ke, ve = q[dst[lo:hi]], k[src[lo:hi]], v[src[lo:hi]]
        do = d_agg[dst[lo:hi]]
        p = _softmax_last(np.einsum("eihc,ejhc->ehij", qe * scale, ke))
        dp = np.einsum("eihc,ejhc->ehij", do, ve)
        ds = p * (dp - (p * dp).sum(axis=-1, keepdims=True))
        np.add.at(dv, src[lo:hi], np.einsum("ehij,eihc->ejhc", p, do))
        np.add.at(dq, dst[lo:hi], np.einsum("ehij,ejhc->eihc", ds, ke) * scale)
        np.add.at(dk, src[lo:hi], np.einsum("ehij,eihc->ejhc", ds, qe) * scale)
    dqkv = np.concatenate([dq.reshape(n, f, d), dk.reshape(n, f, d), dv.reshape(n, f, d)], axis=-1)
    xt = x.reshape(n, f, d)
    d_x = (dqkv @ w_in).reshape(n, width)
    d_w_in = np.einsum("nfa,nfb->ab", dqkv, xt)
    d_b_in = dqkv.sum(axis=(0, 1))
    return {"out": fw["out"], "d_x": d_x, "d_in_proj_weight": d_w_in, "d_in_proj_bias": d_b_in,
            "d_out_proj_weight": d_w_out, "d_out_proj_bias": d_b_out}
