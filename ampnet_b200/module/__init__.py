"""Callers of the hot path: the reference's models around ``AMPConv``."""
from .amp_gcn import AMPGCN, dropout_adj
from .amp_net_classifier import AMPNetClassifier

__all__ = ["AMPGCN", "AMPNetClassifier", "dropout_adj"]
