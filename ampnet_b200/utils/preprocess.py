"""The PCA tokeniser of the reference's Cora example (BASELINE config 1), as tensor code on the device that holds ``x``.

Reference: ``src/ampnet/utils/preprocess.py:7-25`` (``embed_features_old``; ``examples/cora_benchmark.py:9,38`` imports it
under the name ``embed_features``, which the reference package does not define -- both names are provided here).  Every
feature becomes one token ``[PCA embedding of the feature (feature_embed_dim) || the node's value of it, repeated
value_embed_dim times]``; the ``[N, F * (feature_embed_dim + value_embed_dim)]`` matrix is then z-scored column by column.

The reference round-trips through numpy / sklearn on the CPU; here the PCA is one SVD of the centred ``[F, N]`` matrix
(sign convention of the installed sklearn: the largest-magnitude entry of every principal axis is positive) and the
z-score is the population-std StandardScaler formula (zero-variance columns keep scale 1).  Checked against sklearn on
matrices small enough for its exact solver (``tests/test_preprocess.py``); on Cora-sized inputs sklearn switches to a
randomised solver, so the reference itself is only reproducible there up to that solver's tolerance.
"""
import torch

__all__ = ["pca_scores", "embed_features", "embed_features_old"]


def pca_scores(samples, n_components):
    """sklearn ``PCA(n_components).fit_transform(samples)`` for ``samples [S, C]``: scores ``[S, n_components]`` (float64)."""
    a = samples.double()
    centred = a - a.mean(dim=0, keepdim=True)
    _, _, vt = torch.linalg.svd(centred, full_matrices=False)
    vt = vt[:n_components]
    pivot = vt.abs().argmax(dim=1, keepdim=True)
    vt = vt * torch.sign(vt.gather(1, pivot))
    return centred @ vt.t()


def embed_features(x, feature_embed_dim, value_embed_dim):
    """x [N, F] -> tokens [N, F * (feature_embed_dim + value_embed_dim)] float32 (``preprocess.py:7-25``)."""
    n, f = x.shape
    feat = pca_scores(x.t(), feature_embed_dim)                                   # [F, feature_embed_dim]: one row per feature
    tokens = torch.cat([feat.unsqueeze(0).expand(n, f, feature_embed_dim),
                        x.double().unsqueeze(2).expand(n, f, value_embed_dim)], dim=2).reshape(n, f * (feature_embed_dim + value_embed_dim))
    mean = tokens.mean(dim=0, keepdim=True)
    std = tokens.std(dim=0, unbiased=False, keepdim=True)
    std = torch.where(std < 10 * torch.finfo(torch.float64).eps, torch.ones_like(std), std)   # StandardScaler: zero variance -> 1
    return ((tokens - mean) / std).float()


embed_features_old = embed_features
