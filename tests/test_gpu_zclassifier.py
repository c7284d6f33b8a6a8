"""``ampnet_b200.AMPNetClassifier`` against goldens produced by the reference's own class
(``/root/reference/src/ampnet/module/amp_net_classifier_Rahul.py``, executed verbatim in float64, eval mode, by
``oracle/gen_golden_classifier.py``): the reference's state_dict loads with strict=True; log-probabilities, both layer
embeddings, dX and every parameter gradient agree to 1e-4 in the strict fp32 family, and the log-probabilities to 4e-2
(two stacked layers at the 2e-2 per-layer bar) in the bf16 tensor-core family."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def _build(g, mode, dev):
    from ampnet_b200 import AMPNetClassifier
    n, e, f, d, h, classes = [int(v) for v in g["config"]]
    model = AMPNetClassifier(num_heads=h, embed_dim=d, n_original_features=f, out_dim=classes, mode=mode).to(dev).eval()
    state = {k[len("param/"):]: torch.from_numpy(g[k]).float() for k in g.files if k.startswith("param/")}
    model.load_state_dict(state, strict=True)          # reference checkpoints are drop-in
    x = torch.from_numpy(g["x"]).float().to(dev).requires_grad_(True)
    data = SimpleNamespace(x=x, edge_index=torch.from_numpy(g["edge_index"]).to(dev))
    return model, x, data


@pytest.mark.parametrize("name", ["ampnetclf_small", "ampnetclf_d64"])
def test_classifier_matches_reference_model_fp32(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    dev = torch.device("cuda:0")
    model, x, data = _build(g, "fp32", dev)
    out = model(data)
    (out * torch.from_numpy(g["d_out"]).float().to(dev)).sum().backward()
    assert _rel(out.detach().cpu().numpy(), g["out"]) < 1e-4
    assert _rel(model.conv1_embedding.detach().cpu().numpy(), g["conv1_embedding"]) < 1e-4
    assert _rel(model.conv2_embedding.detach().cpu().numpy(), g["conv2_embedding"]) < 1e-4
    assert _rel(x.grad.cpu().numpy(), g["d_x"]) < 2e-4
    for k, p in model.named_parameters():
        ref = g["grad/" + k]
        got = p.grad.cpu().numpy() if p.grad is not None else np.zeros_like(ref)
        assert _rel(got, ref) < 2e-4 or np.abs(ref).max() < 1e-12, k


def test_classifier_runs_on_the_tensor_core_family():
    g = np.load(os.path.join(GOLDEN, "ampnetclf_d64.npz"))
    dev = torch.device("cuda:0")
    model, x, data = _build(g, "bf16", dev)
    out = model(data)
    (out * torch.from_numpy(g["d_out"]).float().to(dev)).sum().backward()
    assert _rel(out.detach().cpu().numpy(), g["out"]) < 4e-2
    assert _rel(model.conv2_embedding.detach().cpu().numpy(), g["conv2_embedding"]) < 4e-2
    assert torch.isfinite(x.grad).all()
