"""Data side of the hot path's callers: GraphSAINT random-walk subgraph sampling, the synthetic XOR
graph of BASELINE config 3, and the double-buffered host -> device feed of full-graph inputs."""
from .graph_saint import GraphSAINTRandomWalkSampler, SubgraphData, cora_shaped_data
from .host_feed import HostFeed
from .synthetic_graph import make_graph, make_inputs
from .synthetic_xor import create_duplicated_xor_data, knn_self_edges

__all__ = ["GraphSAINTRandomWalkSampler", "SubgraphData", "cora_shaped_data", "HostFeed",
           "create_duplicated_xor_data", "knn_self_edges", "make_graph", "make_inputs"]
