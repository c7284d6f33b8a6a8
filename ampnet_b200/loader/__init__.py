"""Data side of the hot path's callers: GraphSAINT random-walk subgraph sampling, and the
double-buffered host -> device feed of full-graph inputs."""
from .graph_saint import GraphSAINTRandomWalkSampler, SubgraphData, cora_shaped_data
from .host_feed import HostFeed

__all__ = ["GraphSAINTRandomWalkSampler", "SubgraphData", "cora_shaped_data", "HostFeed"]
