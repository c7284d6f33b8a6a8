#!/usr/bin/env python
"""BASELINE.json config 3 on one B200: full-graph training of the 2-layer AMPGCN on the synthetic XOR graph.

Mirror of the reference's ``synthetic_benchmark/synthetic_training_modular.py:24-99,124-137`` (duplicated-XOR data with
``num_samples=400, noise_std=0.3, num_nearest_neighbors=20, feature_repeats=1``; model keywords of
``xor_training_utils.py:56-72``; Adam lr 0.01 / weight decay 5e-4, ``NLLLoss``, gradient clipping at 1.0, a fresh test graph of
the same size) with ``ampnet_b200.AMPGCN`` and the tensor-code generator of ``ampnet_b200.loader``; gradient / activation
plots and checkpoint directories are left out.  After training, ``model.conv1.attn_output_weights`` is the ``[8400, 2, 2]``
tensor the reference's ``visualize_attention_coefficients.py:222-232`` reads.

Every part is exercised by the GPU tests (AMPGCN without down-sampling: ``tests/test_gpu_ampgcn.py``, golden ``ampgcn_xor``;
embed 3 / 1 head runs in the strict fp32 kernel family); this script itself has not been run on the B200 yet.

    python experiments/xor_b200.py --epochs 200
"""
import argparse
import os
import sys
import time
from types import SimpleNamespace

import torch
from torch import nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from ampnet_b200 import AMPGCN                                    # noqa: E402
from ampnet_b200.loader import create_duplicated_xor_data         # noqa: E402


def xor_graph(num_samples, noise_std, num_nearest_neighbors, feature_repeats, generator, device):
    x, y, _, edge_index = create_duplicated_xor_data(num_samples, noise_std, num_nearest_neighbors, feature_repeats,
                                                     generator=generator, device=device)
    return SimpleNamespace(x=x, y=y, edge_index=edge_index)


def train(epochs=200, num_samples=400, noise_std=0.3, num_nearest_neighbors=20, feature_repeats=1, lr=0.01, dropout=0.0,
          seed=1, mode="auto", device="cuda:0", log_every=20, quiet=False):
    torch.manual_seed(seed)
    dev = torch.device(device)
    gen = torch.Generator(device=dev).manual_seed(seed)
    train_data = xor_graph(num_samples, noise_std, num_nearest_neighbors, feature_repeats, gen, dev)
    test_data = xor_graph(num_samples, noise_std, num_nearest_neighbors, feature_repeats, gen, dev)
    model = AMPGCN(device=dev, embedding_dim=3, num_heads=1, num_node_features=feature_repeats * 2, num_sampled_vectors=2,
                   output_dim=2, softmax_out=True, feat_emb_dim=2, val_emb_dim=1, downsample_feature_vectors=False,
                   average_pooling_flag=True, dropout_rate=dropout, dropout_adj_rate=dropout, feature_repeats=feature_repeats,
                   mode=mode).to(dev)
    optimizer = torch.optim.Adam(model.parameters(), lr=lr, weight_decay=5e-4)
    criterion = nn.NLLLoss()
    history = []
    t0 = time.time()
    for epoch in range(epochs):
        model.train()
        optimizer.zero_grad()
        out = model(train_data)
        train_loss = criterion(out, train_data.y.long())
        train_acc = float((out.argmax(dim=1) == train_data.y.long()).float().mean())
        train_loss.backward()
        nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        optimizer.step()
        model.eval()
        with torch.no_grad():
            out = model(test_data)
            test_loss = criterion(out, test_data.y.long())
            test_acc = float((out.argmax(dim=1) == test_data.y.long()).float().mean())
        history.append((float(train_loss.detach()), train_acc, float(test_loss), test_acc))
        if not quiet and epoch % log_every == 0:
            print(f"Epoch {epoch:05d} | Train Loss {history[-1][0]:.4f}; Acc {train_acc:.4f} | Test Loss {history[-1][2]:.4f} | "
                  f"Acc {test_acc:.4f}")
    torch.cuda.synchronize(dev)
    if not quiet:
        e = int(train_data.edge_index.shape[1])
        print(f"{epochs} epochs in {time.time() - t0:.1f} s; {e} edges per graph; max train acc {max(h[1] for h in history):.3f}, "
              f"max test acc {max(h[3] for h in history):.3f}; conv1.attn_output_weights {tuple(model.conv1.attn_output_weights.shape)}")
    return model, history


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=200)
    ap.add_argument("--mode", default="auto")
    ap.add_argument("--feature-repeats", type=int, default=1)
    a = ap.parse_args()
    train(epochs=a.epochs, mode=a.mode, feature_repeats=a.feature_repeats)
