"""Callers of the hot path: the reference's 2-layer model around ``AMPConv``."""
from .amp_gcn import AMPGCN, dropout_adj

__all__ = ["AMPGCN", "dropout_adj"]
