"""ampnet_b200.utils.preprocess against the sklearn calls of the reference (src/ampnet/utils/preprocess.py:7-25)."""
import numpy as np
import torch

from ampnet_b200.utils import embed_features, pca_scores


def _reference_embed(x, k, v):
    from sklearn.decomposition import PCA
    from sklearn.preprocessing import StandardScaler
    xt = torch.from_numpy(x)
    gene = torch.from_numpy(PCA(n_components=k).fit_transform(x.transpose()))
    reshaped = torch.reshape(xt, (x.shape[0] * x.shape[1], 1))
    cat = torch.cat([gene.repeat(x.shape[0], 1), reshaped.repeat(1, v)], dim=1)
    per_node = torch.reshape(cat, (x.shape[0], x.shape[1] * (k + v)))
    return StandardScaler().fit_transform(per_node.numpy())


def test_pca_scores_equal_sklearn():
    from sklearn.decomposition import PCA
    rng = np.random.default_rng(0)
    a = rng.normal(size=(60, 25)) @ rng.normal(size=(25, 25))
    ref = PCA(n_components=6).fit_transform(a)
    got = pca_scores(torch.from_numpy(a), 6).numpy()
    assert np.abs(got - ref).max() < 1e-9 * np.abs(ref).max()


def test_embed_features_equals_the_reference_tokeniser():
    rng = np.random.default_rng(1)
    x = (rng.random((30, 40)) < 0.2).astype(np.float64)        # Cora-like binary features, 30 nodes x 40 features
    x[:, 7] = 0.0                                              # a feature no node has: zero-variance value columns
    ref = _reference_embed(x, 8, 4)
    got = embed_features(torch.from_numpy(x), 8, 4)
    assert got.shape == (30, 40 * 12) and got.dtype == torch.float32
    assert np.abs(got.numpy() - ref).max() < 1e-5
