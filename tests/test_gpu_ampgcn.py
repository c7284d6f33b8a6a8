"""The 2-layer model ``ampnet_b200.AMPGCN`` against goldens produced by the reference's own ``AMPGCN``
(``/root/reference/src/ampnet/module/amp_gcn.py``, executed verbatim in float64 by ``oracle/gen_golden_gcn.py``):
the reference's state_dict loads with strict=True, and with the reference's sampled feature indices the
log-probabilities, both layer embeddings and every parameter gradient agree to 1e-4 (strict fp32 family)."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("name", ["ampgcn_sampled", "ampgcn_xor"])
def test_ampgcn_matches_reference_model(name):
    from ampnet_b200 import AMPGCN
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    n, e, nf, s, d, h, classes, downsample, repeats, avg, softmax = [int(v) for v in g["config"]]
    dev = torch.device("cuda:0")
    model = AMPGCN(device=dev, embedding_dim=d, num_heads=h, num_node_features=nf, num_sampled_vectors=s,
                   output_dim=classes, softmax_out=bool(softmax), feat_emb_dim=d - 1, val_emb_dim=1,
                   downsample_feature_vectors=bool(downsample), average_pooling_flag=bool(avg), dropout_rate=0.0,
                   dropout_adj_rate=0.0, feature_repeats=repeats, mode="fp32").to(dev)
    state = {k[len("param/"):]: torch.from_numpy(g[k]).float() for k in g.files if k.startswith("param/")}
    model.load_state_dict(state, strict=True)          # reference checkpoints are drop-in
    data = SimpleNamespace(x=torch.from_numpy(g["x"]).float(), edge_index=torch.from_numpy(g["edge_index"]))
    idx = g["sampled_indices"] if downsample else None
    out = model(data, sampled_indices=idx)
    (out * torch.from_numpy(g["d_out"]).float().to(dev)).sum().backward()
    assert _rel(out.detach().cpu().numpy(), g["out"]) < 1e-4
    assert _rel(model.conv1_embedding.detach().cpu().numpy(), g["conv1_embedding"]) < 1e-4
    assert _rel(model.conv2_embedding.detach().cpu().numpy(), g["conv2_embedding"]) < 1e-4
    if downsample:
        assert np.array_equal(model.sampled_node_feat_indices, g["sampled_indices"])
    for k, p in model.named_parameters():
        ref = g["grad/" + k]
        got = p.grad.cpu().numpy() if p.grad is not None else np.zeros_like(ref)
        assert _rel(got, ref) < 2e-4 or np.abs(ref).max() < 1e-12, k


def test_ampgcn_own_sampler_draws_present_features_only():
    from ampnet_b200 import AMPGCN
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(2)
    x = (torch.rand(50, 30, generator=g) < 0.2).float()
    x[torch.arange(50), torch.randint(0, 30, (50,), generator=g)] = 1.0
    ei = torch.randint(0, 50, (2, 200), generator=g)
    model = AMPGCN(device=dev, embedding_dim=16, num_heads=2, num_node_features=30, num_sampled_vectors=5, output_dim=3,
                   feat_emb_dim=15, val_emb_dim=1, dropout_rate=0.0, dropout_adj_rate=0.0, mode="fp32").to(dev)
    out = model(SimpleNamespace(x=x, edge_index=ei))
    assert out.shape == (50, 3) and torch.isfinite(out).all()
    idx = model.sampled_node_feat_indices
    assert idx.shape == (50, 5)
    assert bool((x.numpy()[np.arange(50)[:, None], idx] != 0).all())


def test_graphsaint_training_on_cora_shaped_graph_learns():
    """Config 2 of BASELINE.json in miniature: a few GraphSAINT iterations of the 2-layer model on a Cora-shaped graph;
    the node-normalised training loss must go down and the sampled subgraphs must be consistent."""
    import importlib.util
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "experiments", "cora_graphsaint_b200.py")
    spec = importlib.util.spec_from_file_location("cora_graphsaint_b200", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    from ampnet_b200.loader import cora_shaped_data
    data = cora_shaped_data(num_nodes=600, num_features=140, num_undirected_edges=1500, nnz_per_node=8, num_train=200,
                            num_val=100, num_test=200, seed=3)
    model, hist = mod.train(iters=40, batch_size=4, walk_length=40, num_steps=20, sample_coverage=10, lr=0.01, seed=3,
                            embedding_dim=32, num_heads=4, num_sampled_vectors=8, mode="fp32", data=data, quiet=True)
    losses = np.array([h[0] for h in hist])
    assert np.isfinite(losses).all()
    assert losses[-10:].mean() < 0.8 * losses[:10].mean()
    assert all(h[2] > 0 and h[3] >= 0 for h in hist)
