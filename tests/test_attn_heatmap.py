"""ampnet_b200.visualization: the class-pair attention heat map against a literal restatement of the reference's three
nested loops (experiments/visualize_cora_attn_coeffs.py:15-34, 68-109) on a small case with repeated sampled features."""
import numpy as np
import torch

from ampnet_b200.visualization import class_pair_attention_heatmap, edge_indices_between_classes, top_features_for_class


def _reference_loops(w, sampled, edge_index, y, src_top, dst_top, class_src, class_dst):
    edge_idxs = np.array([e for e in range(edge_index.shape[1])
                          if y[edge_index[0, e]] == class_src and y[edge_index[1, e]] == class_dst], dtype=np.int64)
    edges = edge_index[:, edge_idxs]
    ew = w[edge_idxs]
    src_feats, dst_feats = sampled[edges[0]], sampled[edges[1]]
    heat = np.zeros((len(src_top), len(dst_top)))
    cnt = np.zeros_like(heat)
    for e in range(ew.shape[0]):
        for dst_idx, df in enumerate(dst_feats[e]):
            for src_idx, sf in enumerate(src_feats[e]):
                if sf in src_top and df in dst_top:
                    r = np.where(src_top == sf)[0]
                    c = np.where(dst_top == df)[0]
                    heat[r, c] += ew[e, dst_idx, src_idx]
                    cnt[r, c] += 1
    return np.divide(heat, cnt, out=np.zeros_like(heat), where=cnt != 0), edge_idxs


def test_heatmap_equals_the_reference_loops():
    rng = np.random.default_rng(5)
    n, e, f, nf, classes = 40, 300, 6, 25, 3
    y = rng.integers(0, classes, n)
    edge_index = rng.integers(0, n, (2, e))
    sampled = rng.integers(0, nf, (n, f))                      # drawn with replacement: repeated features per node
    w = rng.random((e, f, f))
    src_top = rng.permutation(nf)[:8]
    dst_top = rng.permutation(nf)[:7]
    ref, ref_edges = _reference_loops(w, sampled, edge_index, y, src_top, dst_top, 1, 2)
    got = class_pair_attention_heatmap(torch.from_numpy(w), sampled, torch.from_numpy(edge_index), y, 1, 2, src_top, dst_top,
                                       chunk_edges=17)
    assert got.shape == (8, 7)
    assert np.abs(got.numpy() - ref).max() < 1e-12
    assert np.array_equal(edge_indices_between_classes(torch.from_numpy(edge_index), y, 1, 2).numpy(), ref_edges)
    assert (ref != 0).sum() > 10                                # the case is not vacuous


def test_top_features_are_the_most_present_ones():
    rng = np.random.default_rng(6)
    x = (rng.random((50, 30)) < 0.3).astype(np.float32)
    y = rng.integers(0, 2, 50)
    top = top_features_for_class(torch.from_numpy(x), y, 1, k=5).numpy()
    counts = x[y == 1].sum(0)
    # same set as the reference's np.argpartition(counts, -5)[-5:] whenever the 5th and 6th counts differ
    assert counts[top].min() >= np.sort(counts)[-5]
    assert len(set(top.tolist())) == 5 and list(counts[top]) == sorted(counts[top], reverse=True)
