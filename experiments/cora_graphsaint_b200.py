#!/usr/bin/env python
"""BASELINE.json config 2 on one B200: GraphSAINT random-walk subgraph training of the 2-layer AMPGCN.

Mirror of the reference's ``experiments/cora_benchmark_graphsaint.py:59-131`` (model keywords, sampler settings, Adam +
cosine warm restarts, the ``nll_loss * node_norm`` masked sum) with ``ampnet_b200.AMPGCN`` and the torch-only sampler of
``ampnet_b200.loader``; plotting / checkpoint directories are left out.  Planetoid Cora cannot be downloaded in this
environment, so the data is the Cora-shaped synthetic graph of ``cora_shaped_data`` (same N, feature width, edge count,
classes and split sizes).  Everything -- sampling, tokeniser, both AMPConv layers, optimiser -- runs on the GPU.

    python experiments/cora_graphsaint_b200.py --iters 200
"""
import argparse
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from ampnet_b200 import AMPGCN                                             # noqa: E402
from ampnet_b200.loader import GraphSAINTRandomWalkSampler, cora_shaped_data   # noqa: E402


def train(iters=200, batch_size=8, walk_length=150, num_steps=200, sample_coverage=100, lr=0.1, seed=1, embedding_dim=128,
          num_heads=4, num_sampled_vectors=20, mode="auto", log_every=10, data=None, device="cuda:0", quiet=False):
    torch.manual_seed(seed)
    dev = torch.device(device)
    data = (data if data is not None else cora_shaped_data(seed=seed)).to(dev)
    num_classes = int(data.y.max()) + 1
    model = AMPGCN(device=dev, embedding_dim=embedding_dim, num_heads=num_heads, num_node_features=data.x.size(1),
                   num_sampled_vectors=num_sampled_vectors, output_dim=num_classes, softmax_out=True,
                   feat_emb_dim=embedding_dim - 1, val_emb_dim=1, downsample_feature_vectors=True, average_pooling_flag=True,
                   dropout_rate=0.0, dropout_adj_rate=0.0, feature_repeats=None, mode=mode).to(dev)
    gen = torch.Generator(device=dev).manual_seed(seed)
    loader = GraphSAINTRandomWalkSampler(data, batch_size=batch_size, walk_length=walk_length, num_steps=num_steps,
                                         sample_coverage=sample_coverage, generator=gen)
    optimizer = torch.optim.Adam(model.parameters(), lr=lr, weight_decay=1e-4)
    scheduler = torch.optim.lr_scheduler.CosineAnnealingWarmRestarts(optimizer, T_0=400, T_mult=2)
    history = []
    it = 0
    t0 = time.time()
    edges = 0
    while it < iters:
        for batch in loader:
            model.train()
            optimizer.zero_grad()
            out = model(batch)
            loss = F.nll_loss(out, batch.y, reduction="none")
            loss = (loss * batch.node_norm)[batch.train_mask].sum()
            loss.backward()
            optimizer.step()
            scheduler.step()
            edges += 2 * batch.num_edges                                   # two AMPConv layers
            with torch.no_grad():
                mask = batch.test_mask
                acc = float((out[mask].argmax(dim=1) == batch.y[mask]).float().mean()) if bool(mask.any()) else float("nan")
            history.append((float(loss.detach()), acc, batch.num_nodes, batch.num_edges))
            if not quiet and it % log_every == 0:
                print(f"iter {it:5d} lr {scheduler.get_last_lr()[0]:.4f} | nodes {batch.num_nodes:5d} edges {batch.num_edges:6d} | "
                      f"train NLL {float(loss.detach()):.4f} | test acc {acc:.3f}")
            it += 1
            if it >= iters:
                break
    torch.cuda.synchronize(dev)
    dt = time.time() - t0
    if not quiet:
        print(f"{iters} iterations in {dt:.1f} s ({edges / dt:,.0f} layer-edges/s incl. sampling, tokeniser, optimiser)")
    return model, history


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--mode", default="auto")
    ap.add_argument("--lr", type=float, default=0.1)
    a = ap.parse_args()
    train(iters=a.iters, mode=a.mode, lr=a.lr)
