"""Class-pair attention heat map of the reference's Cora analysis, without the Python loops.

The reference walks every edge between two node classes and every (destination token, source token) pair in three nested
Python loops, looks each sampled feature id up in the classes' top-30 lists with ``np.where`` and averages the matching
coefficients (``experiments/visualize_cora_attn_coeffs.py:68-109``; edge selection ``:15-34``, top features ``:37-63``).
Here the same reduction is two table look-ups and one ``index_add_`` per edge chunk, on whatever device the coefficients
are on -- ``conv.attn_output_weights`` of ``ampnet_b200.AMPConv`` is ``[E, F, F]`` (row = destination token, column = source
token, original edge order), exactly what the reference's function consumes (``:115-123``).
"""
import torch

__all__ = ["edge_indices_between_classes", "top_features_for_class", "class_pair_attention_heatmap"]


def edge_indices_between_classes(edge_index, y, class_src, class_dst):
    """Indices of the edges that go from a node of ``class_src`` to a node of ``class_dst``, ascending
    (``get_edge_indices_between_nodes``, ``visualize_cora_attn_coeffs.py:15-34``)."""
    y = torch.as_tensor(y, device=edge_index.device)
    keep = (y[edge_index[0]] == class_src) & (y[edge_index[1]] == class_dst)
    return torch.nonzero(keep, as_tuple=False).flatten()


def top_features_for_class(x, y, class_idx, k=30):
    """The ``k`` features most often present in nodes of class ``class_idx`` (``get_top_30_feature_idxs_for_class``,
    ``visualize_cora_attn_coeffs.py:37-63``).  The reference takes them with ``np.argpartition`` (same set, unspecified
    order); here they come back ordered by descending count, ties by ascending feature id."""
    y = torch.as_tensor(y, device=x.device)
    counts = x[y == class_idx].sum(dim=0)
    order = torch.sort(-counts.double(), stable=True).indices
    return order[:k]


def class_pair_attention_heatmap(weights, sampled_features, edge_index, y, class_src, class_dst, src_top, dst_top,
                                 chunk_edges=4096):
    """Average attention coefficient between the top features of two node classes
    (``calculate_attn_heatmap``, ``visualize_cora_attn_coeffs.py:68-109``).

    weights           [E, F, F]  coefficients per edge, row = destination token, column = source token
    sampled_features  [N, F]     feature id of every token of every node (``model.sampled_node_feat_indices``)
    src_top, dst_top  [K], [K']  feature ids of the rows (source class) / columns (destination class) of the heat map
    returns           [K, K']    float64; cell (r, c) = mean of ``weights[e, j, i]`` over the class-pair edges ``e`` and the token
                                 pairs with ``sampled[dst(e), j] == dst_top[c]`` and ``sampled[src(e), i] == src_top[r]``;
                                 0 where no coefficient matched (as the reference's masked division)."""
    dev = weights.device
    sampled = torch.as_tensor(sampled_features, device=dev).long()
    src_top = torch.as_tensor(src_top, device=dev).long()
    dst_top = torch.as_tensor(dst_top, device=dev).long()
    edge_index = edge_index.to(dev)
    k_src, k_dst = src_top.numel(), dst_top.numel()
    num_features = int(max(sampled.max(), src_top.max(), dst_top.max())) + 1
    lut_src = torch.full((num_features,), -1, dtype=torch.long, device=dev)
    lut_dst = torch.full((num_features,), -1, dtype=torch.long, device=dev)
    lut_src[src_top] = torch.arange(k_src, device=dev)
    lut_dst[dst_top] = torch.arange(k_dst, device=dev)
    edges = edge_indices_between_classes(edge_index, y, class_src, class_dst)
    total = torch.zeros(k_src * k_dst, dtype=torch.float64, device=dev)
    count = torch.zeros(k_src * k_dst, dtype=torch.float64, device=dev)
    for lo in range(0, edges.numel(), chunk_edges):
        eid = edges[lo:lo + chunk_edges]
        row = lut_src[sampled[edge_index[0, eid]]]            # [c, F] heat-map row of every source token (-1: not listed)
        col = lut_dst[sampled[edge_index[1, eid]]]            # [c, F] heat-map column of every destination token
        cell = row[:, None, :] * k_dst + col[:, :, None]       # [c, F_dst, F_src], aligned with weights[e, j, i]
        hit = (row[:, None, :] >= 0) & (col[:, :, None] >= 0)
        w = weights[eid].double()
        total.index_add_(0, cell[hit], w[hit])
        count.index_add_(0, cell[hit], torch.ones_like(w[hit]))
    heat = torch.where(count > 0, total / count.clamp(min=1), torch.zeros_like(total))
    return heat.view(k_src, k_dst)
