"""Generate ``tests/golden/ampnetclf_*.npz`` by executing the REFERENCE's own ``AMPNetClassifier``
(``/root/reference/src/ampnet/module/amp_net_classifier_Rahul.py:7-57``, loaded by path, unmodified) on CPU in float64.

    python -m oracle.gen_golden_classifier          # build container only (/root/reference must exist)

The model runs in eval mode (its three dropouts have a hard-wired p = 0.6, ``:47,51,55``; eval makes them the identity so
that the fixture is deterministic).  Stored: inputs, the state_dict, log-probabilities, both layer embeddings and the
gradients of ``(out * d_out).sum()`` w.r.t. the input tokens and every parameter.  TEST INFRASTRUCTURE ONLY."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

from . import cases, reference_loader

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    # generic shape (strict fp32 kernels): 5 tokens of width 12, 3 heads
    "ampnetclf_small": dict(n=70, e=420, f=5, d=12, h=3, classes=4, graph="skewed"),
    # tensor-core token width (d = 64, head_dim 16): checked in both kernel families
    "ampnetclf_d64": dict(n=40, e=200, f=6, d=64, h=4, classes=3, graph="uniform"),
}


def main():
    if not reference_loader.available():
        sys.exit("reference tree not found; goldens can only be regenerated in the build container")
    mod = reference_loader.load_amp_net_classifier_module()
    torch.set_default_dtype(torch.float64)
    for name, c in CASES.items():
        rng = np.random.default_rng(31 + len(name))
        x = rng.normal(size=(c["n"], c["f"] * c["d"]))
        edge_index = cases.make_graph(c["graph"], c["n"], c["e"], seed=29)
        torch.manual_seed(11)
        model = mod.AMPNetClassifier(num_heads=c["h"], embed_dim=c["d"], n_original_features=c["f"],
                                     out_dim=c["classes"]).double().eval()
        with torch.no_grad():
            for conv in (model.conv1, model.conv2):
                conv.multi_head_attention.in_proj_bias.normal_(0, 0.1)
                conv.multi_head_attention.out_proj.bias.normal_(0, 0.1)
        xt = torch.from_numpy(x).requires_grad_(True)
        data = SimpleNamespace(x=xt, edge_index=torch.from_numpy(edge_index).long())
        out = model(data)
        d_out = torch.from_numpy(rng.normal(size=tuple(out.shape)))
        (out * d_out).sum().backward()
        payload = {
            "x": x, "edge_index": edge_index.astype(np.int64), "d_out": d_out.numpy(), "out": out.detach().numpy(),
            "d_x": xt.grad.numpy(),
            "conv1_embedding": model.conv1_embedding.detach().numpy(), "conv2_embedding": model.conv2_embedding.detach().numpy(),
            "config": np.array([c["n"], c["e"], c["f"], c["d"], c["h"], c["classes"]], dtype=np.int64),
        }
        for k, v in model.state_dict().items():
            payload["param/" + k] = v.numpy()
        for k, p in model.named_parameters():
            payload["grad/" + k] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy()
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **payload)
        print(f"{name}: out {tuple(out.shape)}, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
