"""Pre-processing on the caller side of the hot path (mirror of ``src/ampnet/utils``; plotting helpers are out of scope)."""
from .preprocess import embed_features, embed_features_old, pca_scores

__all__ = ["embed_features", "embed_features_old", "pca_scores"]
