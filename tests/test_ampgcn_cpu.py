"""CPU checks of the AMPGCN mirror (no kernels): state_dict key compatibility with the reference model and the
device tokeniser against the reference formula ``[embedding_table[f] || z-score(x)[n, f]]`` (``amp_gcn.py:120-153``),
both pinned by the goldens the reference's own AMPGCN produced (oracle/gen_golden_gcn.py)."""
import os

import numpy as np
import pytest
import torch

from ampnet_b200 import AMPGCN

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _model(g):
    n, e, nf, s, d, h, classes, downsample, repeats, avg, softmax = [int(v) for v in g["config"]]
    return AMPGCN(device="cpu", embedding_dim=d, num_heads=h, num_node_features=nf, num_sampled_vectors=s,
                  output_dim=classes, softmax_out=bool(softmax), feat_emb_dim=d - 1, val_emb_dim=1,
                  downsample_feature_vectors=bool(downsample), average_pooling_flag=bool(avg), dropout_rate=0.0,
                  dropout_adj_rate=0.0, feature_repeats=repeats)


@pytest.mark.parametrize("name", ["ampgcn_sampled", "ampgcn_xor"])
def test_reference_state_dict_loads_strictly(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    model = _model(g)
    state = {k[len("param/"):]: torch.from_numpy(g[k]).float() for k in g.files if k.startswith("param/")}
    assert sorted(state) == sorted(model.state_dict().keys())
    model.load_state_dict(state, strict=True)


@pytest.mark.parametrize("name", ["ampgcn_sampled", "ampgcn_xor"])
def test_tokeniser_matches_reference_formula(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    model = _model(g)
    table = g["param/feature_embedding_table.weight"]
    model.feature_embedding_table.weight.data = torch.from_numpy(table).float()
    x = g["x"]
    std = x.std(axis=0)
    z = (x - x.mean(axis=0)) / np.where(std < 1e-12, 1.0, std)          # StandardScaler: population std, 0 -> 1
    n = x.shape[0]
    if model.downsampling_vectors:
        idx = g["sampled_indices"]
        ref = np.concatenate([table[idx], z[np.arange(n)[:, None], idx][..., None]], axis=2).reshape(n, -1)
        tok, got_idx = model.normalize_features_and_add_feature_table_embedding(torch.from_numpy(x).float(), idx)
        assert np.array_equal(got_idx.numpy(), idx)
    else:
        ref = np.concatenate([np.broadcast_to(np.tile(table, (model.feature_repeats, 1)), (n,) + table.shape),
                              z[..., None]], axis=2).reshape(n, -1)
        tok, got_idx = model.normalize_features_and_add_feature_table_embedding(torch.from_numpy(x).float())
        assert got_idx is None
    assert np.abs(tok.detach().numpy() - ref).max() < 1e-5


def test_dropout_adj_semantics():
    from ampnet_b200.module import dropout_adj
    ei = torch.arange(20).reshape(2, 10)
    assert dropout_adj(ei, p=0.5, training=False)[0] is ei and dropout_adj(ei, p=0.0, training=True)[0] is ei
    torch.manual_seed(0)
    kept = dropout_adj(torch.arange(20000).reshape(2, 10000), p=0.3, training=True)[0]
    assert 0.65 < kept.size(1) / 10000 < 0.75
    with pytest.raises(ValueError):
        dropout_adj(ei, p=1.5)


@pytest.mark.parametrize("name", ["ampnetclf_small", "ampnetclf_d64"])
def test_reference_classifier_state_dict_loads_strictly(name):
    """AMPNetClassifier mirror: same parameter names and shapes as the reference class
    (amp_net_classifier_Rahul.py:7-43), pinned by the state_dict the reference's own model produced."""
    from ampnet_b200 import AMPNetClassifier
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    n, e, f, d, h, classes = [int(v) for v in g["config"]]
    model = AMPNetClassifier(num_heads=h, embed_dim=d, n_original_features=f, out_dim=classes)
    state = {k[len("param/"):]: torch.from_numpy(g[k]).float() for k in g.files if k.startswith("param/")}
    assert sorted(state) == sorted(model.state_dict().keys())
    model.load_state_dict(state, strict=True)
    with pytest.raises((TypeError, RuntimeError, ImportError)):
        # CPU tensors: the layers have no CPU path
        from types import SimpleNamespace
        model(SimpleNamespace(x=torch.zeros(n, f * d), edge_index=torch.zeros(2, 1, dtype=torch.long)))


def test_classifier_glue_with_stand_in_layers():
    """The glue around the two layers (dropout off in eval, ELU after each layer, read-out, log-softmax, embeddings kept),
    with the CUDA layers replaced by CPU stand-ins: amp_net_classifier_Rahul.py:45-57."""
    from types import SimpleNamespace
    from ampnet_b200 import AMPNetClassifier
    torch.manual_seed(0)
    model = AMPNetClassifier(num_heads=2, embed_dim=4, n_original_features=3, out_dim=5).eval()
    assert [k for k, _ in model.named_children()] == ["layer_norm", "conv1", "post_conv_linear1", "conv2", "post_conv_linear2",
                                                      "linear_out"]
    model.conv1.forward = lambda x, ei: x * 2.0 - 1.0
    model.conv2.forward = lambda x, ei: x.flip(1) + 0.5
    x = torch.randn(7, 12)
    out = model(SimpleNamespace(x=x, edge_index=torch.zeros(2, 1, dtype=torch.long)))
    e1 = x * 2.0 - 1.0
    e2 = torch.nn.functional.elu(e1).flip(1) + 0.5
    ref = torch.log_softmax(model.linear_out(torch.nn.functional.elu(e2)), dim=1)
    assert torch.allclose(out, ref) and torch.equal(model.conv1_embedding, e1) and torch.equal(model.conv2_embedding, e2)
    model.train()
    torch.manual_seed(1)
    a = model(SimpleNamespace(x=x, edge_index=torch.zeros(2, 1, dtype=torch.long)))
    assert not torch.allclose(a, ref)                      # dropout(0.6) is live in training mode
