"""Stand-in for ``torch_geometric.nn.MessagePassing`` (PyG is not installable here).

Restates only the semantics the AMPConv path uses
(reference call sites ``src/ampnet/conv/amp_conv.py:9,11,25``; semantics pinned by
``synthetic_benchmark/testing_message_passing_pyg.py:6-19,37-40``):

* ``x_j = x[edge_index[0]]`` (source), ``x_i = x[edge_index[1]]`` (destination);
* ``message(x_i=..., x_j=...)`` produces one row per edge;
* ``aggr='mean'``: row ``n`` of the output is the mean of the messages whose
  destination is ``n`` and exactly zero when there is none; duplicate edges and
  self loops are ordinary edges.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).
"""
import sys
import types

import torch


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr="add", flow="source_to_target", node_dim=0):
        super().__init__()
        if flow != "source_to_target" or node_dim != 0:
            raise NotImplementedError("stub covers the AMPConv configuration only")
        if aggr not in ("add", "sum", "mean"):
            raise NotImplementedError(aggr)
        self.aggr = aggr

    def propagate(self, edge_index, x=None, size=None):
        src, dst = edge_index[0], edge_index[1]
        n = x.size(0)
        msg = self.message(x_i=x.index_select(0, dst), x_j=x.index_select(0, src))
        out = msg.new_zeros((n, msg.size(1))).index_add_(0, dst, msg)
        if self.aggr == "mean":
            deg = torch.bincount(dst, minlength=n).clamp(min=1).to(msg.dtype)
            out = out / deg.unsqueeze(1)
        return out

    def message(self, x_j):  # pragma: no cover - overridden by users
        return x_j


def install():
    """Register the stand-in as ``torch_geometric`` / ``torch_geometric.nn``."""
    if "torch_geometric" in sys.modules and not getattr(sys.modules["torch_geometric"], "_ampnet_stub", False):
        return sys.modules["torch_geometric"]
    pkg = types.ModuleType("torch_geometric")
    pkg._ampnet_stub = True
    nn_mod = types.ModuleType("torch_geometric.nn")
    nn_mod.MessagePassing = MessagePassing
    pkg.nn = nn_mod
    sys.modules["torch_geometric"] = pkg
    sys.modules["torch_geometric.nn"] = nn_mod
    return pkg
