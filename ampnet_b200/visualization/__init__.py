"""Reductions behind the reference's attention-coefficient figures, as tensor code that runs
where the coefficients live.  Plotting itself (seaborn heat maps / cluster maps) stays with the caller."""
from .attn_coeffs import class_pair_attention_heatmap, edge_indices_between_classes, top_features_for_class

__all__ = ["class_pair_attention_heatmap", "edge_indices_between_classes", "top_features_for_class"]
