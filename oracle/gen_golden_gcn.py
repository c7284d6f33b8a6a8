"""Generate ``tests/golden/ampgcn_*.npz`` by executing the REFERENCE's own 2-layer model ``AMPGCN``
(``/root/reference/src/ampnet/module/amp_gcn.py:19-276``, loaded by path, unmodified) on CPU in float64.

    python -m oracle.gen_golden_gcn          # build container only (/root/reference must exist)

Two cases: the down-sampling tokeniser of the Cora experiments (``amp_gcn.py:128-153``: present features sampled with
replacement, ``np.random.seed`` fixed so the sampled indices are part of the fixture) and the non-sampling XOR tokeniser
(``amp_gcn.py:168-180``).  Stored: inputs, the model's state_dict, the sampled indices, log-probabilities, both layer
embeddings, and the gradients of ``(out * d_out).sum()`` w.r.t. every parameter.  TEST INFRASTRUCTURE ONLY."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

from . import cases, reference_loader

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    # Cora-like: binary features, a few present per node, S sampled tokens per node
    "ampgcn_sampled": dict(n=60, e=360, nf=40, s=6, d=16, h=2, classes=5, downsample=True, repeats=1, avg=True, softmax=True),
    # XOR-like: every feature is a token (no sampling), token 0 pooling, sigmoid head
    "ampgcn_xor": dict(n=48, e=300, nf=4, s=4, d=6, h=2, classes=1, downsample=False, repeats=1, avg=False, softmax=False),
}


def main():
    if not reference_loader.available():
        sys.exit("reference tree not found; goldens can only be regenerated in the build container")
    mod = reference_loader.load_amp_gcn_module()
    torch.set_default_dtype(torch.float64)
    for name, c in CASES.items():
        rng = np.random.default_rng(17 + len(name))
        if c["downsample"]:
            x = (rng.random((c["n"], c["nf"])) < 0.15).astype(np.float64)
            x[np.arange(c["n"]), rng.integers(0, c["nf"], c["n"])] = 1.0      # at least one present feature per node
        else:
            x = rng.normal(size=(c["n"], c["nf"]))
        edge_index = cases.make_graph("skewed", c["n"], c["e"], seed=23)
        torch.manual_seed(5)
        model = mod.AMPGCN(device="cpu", embedding_dim=c["d"], num_heads=c["h"], num_node_features=c["nf"],
                           num_sampled_vectors=c["s"], output_dim=c["classes"], softmax_out=c["softmax"],
                           feat_emb_dim=c["d"] - 1, val_emb_dim=1, downsample_feature_vectors=c["downsample"],
                           average_pooling_flag=c["avg"], dropout_rate=0.0, dropout_adj_rate=0.0,
                           feature_repeats=c["repeats"]).double()
        # non-trivial biases so that every term of the forward is exercised
        with torch.no_grad():
            for conv in (model.conv1, model.conv2):
                conv.multi_head_attention.in_proj_bias.normal_(0, 0.1)
                conv.multi_head_attention.out_proj.bias.normal_(0, 0.1)
        np.random.seed(99)                                                     # the tokeniser samples with np.random.choice
        data = SimpleNamespace(x=torch.from_numpy(x), edge_index=torch.from_numpy(edge_index).long())
        out = model(data)
        d_out = torch.from_numpy(rng.normal(size=tuple(out.shape)))
        (out * d_out).sum().backward()
        payload = {
            "x": x, "edge_index": edge_index.astype(np.int64), "d_out": d_out.numpy(), "out": out.detach().numpy(),
            "conv1_embedding": model.conv1_embedding.detach().numpy(), "conv2_embedding": model.conv2_embedding.detach().numpy(),
            "sampled_indices": (model.sampled_node_feat_indices if model.sampled_node_feat_indices is not None
                                else np.zeros((0,), dtype=np.int64)),
            "config": np.array([c["n"], c["e"], c["nf"], c["s"], c["d"], c["h"], c["classes"], int(c["downsample"]),
                                c["repeats"], int(c["avg"]), int(c["softmax"])], dtype=np.int64),
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
