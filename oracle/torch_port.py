"""Eager-PyTorch port of the reference AMPConv op chain (CPU).  TEST INFRASTRUCTURE ONLY.

This restates *what the reference executes*, op for op, so that it can be timed
on the GPU box's host cores (``bench.py`` ``cpu_baseline`` / ``--impl reference``,
kind "port": ``/root/reference`` does not travel to the GPU box) and used as an
autograd checker:

* gather ``x_j = x[src]``, ``x_i = x[dst]`` -- PyG ``propagate``
  (reference ``src/ampnet/conv/amp_conv.py:24-26``);
* reshape ``[E, F*d] -> [E, F, d]`` (``amp_conv.py:35-36``);
* stock ``torch.nn.MultiheadAttention(batch_first=True)`` with query = destination
  tokens and key = value = source tokens (``amp_conv.py:18-22,39``), whose side
  outputs are kept as ``attn_output`` / ``attn_output_weights``;
* flatten (``amp_conv.py:49``) and mean-aggregate at the destination
  (``amp_conv.py:11``; ``testing_message_passing_pyg.py:37-40``).

Parameter names match the reference's state_dict
(``multi_head_attention.{in_proj_weight,in_proj_bias,out_proj.weight,out_proj.bias}``).
"""
import torch
import torch.nn as nn


class AMPConvPort(nn.Module):
    def __init__(self, embed_dim, num_heads):
        super().__init__()
        self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.attn_output = None
        self.attn_output_weights = None
        self.multi_head_attention = nn.MultiheadAttention(
            embed_dim=embed_dim, num_heads=num_heads, batch_first=True, bias=True)

    def forward(self, x, edge_index):
        src, dst = edge_index[0], edge_index[1]
        n, width = x.shape
        if width % self.embed_dim != 0:
            raise ValueError("x.shape[1] must be a multiple of embed_dim")
        tokens = width // self.embed_dim
        x_i = x.index_select(0, dst).reshape(-1, tokens, self.embed_dim)
        x_j = x.index_select(0, src).reshape(-1, tokens, self.embed_dim)
        if x_i.shape[0] == 0:
            self.attn_output = x.new_zeros((0, tokens, self.embed_dim))
            self.attn_output_weights = x.new_zeros((0, tokens, tokens))
            return x.new_zeros((n, width))
        self.attn_output, self.attn_output_weights = self.multi_head_attention(
            query=x_i, key=x_j, value=x_j)
        msg = self.attn_output.reshape(-1, width)
        out = msg.new_zeros((n, width)).index_add_(0, dst, msg)
        deg = torch.bincount(dst, minlength=n).clamp(min=1).to(msg.dtype)
        return out / deg.unsqueeze(1)


def fwd_bwd_chunked(conv, x, edge_index, d_out, chunk_edges):
    """fwd+bwd of ``conv`` with the edge set processed in chunks (mean is linear, so
    chunks of the *sum* are accumulated and divided by the full in-degree).  Used to time
    C4/C5 token shapes on a bounded edge sample without holding [E,H,F,F] at once."""
    n = x.shape[0]
    deg = torch.bincount(edge_index[1], minlength=n).clamp(min=1).to(x.dtype)
    out = torch.zeros_like(x)
    x = x.detach().requires_grad_(True)
    width = x.shape[1]
    tokens = width // conv.embed_dim
    for lo in range(0, edge_index.shape[1], chunk_edges):
        ei = edge_index[:, lo:lo + chunk_edges]
        x_i = x.index_select(0, ei[1]).reshape(-1, tokens, conv.embed_dim)
        x_j = x.index_select(0, ei[0]).reshape(-1, tokens, conv.embed_dim)
        o, _ = conv.multi_head_attention(query=x_i, key=x_j, value=x_j)
        msg = o.reshape(-1, width)
        part = msg.new_zeros((n, width)).index_add_(0, ei[1], msg) / deg.unsqueeze(1)
        (part * d_out).sum().backward()
        out += part.detach()
    return out, x.grad
