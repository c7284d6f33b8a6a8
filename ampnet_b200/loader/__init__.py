"""Data side of the hot path's callers: GraphSAINT random-walk subgraph sampling."""
from .graph_saint import GraphSAINTRandomWalkSampler, SubgraphData, cora_shaped_data

__all__ = ["GraphSAINTRandomWalkSampler", "SubgraphData", "cora_shaped_data"]
