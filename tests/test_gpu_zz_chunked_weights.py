"""Chunked reader of the ``attn_output_weights`` side output (at the ogbn-arxiv shape all of [E, F, F] is 76 GB,
so a caller like synthetic_benchmark/visualize_attention_coefficients.py:221-232 must be able to walk it in slices):
``conv.attention_weights(edge_ids)`` and ``conv.attention_weights_chunks()`` against the full tensor and the reference golden."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from test_gpu_parity import _make_conv, _run

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode,tol", [("fp32", 1e-5), ("bf16", 2e-2)])
def test_sliced_weights_equal_the_full_tensor_and_the_golden(mode, tol):
    dev = torch.device("cuda:0")
    g = load_golden("c4_tokens")
    conv = _make_conv(g["d"], g["h"], g["params"], dev, mode=mode)
    _run(conv, g["x"], g["edge_index"], g["d_out"], dev)
    full = conv.attn_output_weights
    ids = torch.tensor([5, 0, g["e"] - 1, 7, 7], device=dev)
    part = conv.attention_weights(ids)
    assert tuple(part.shape) == (5, g["f"], g["f"])
    assert torch.equal(part, full[ids])
    we = g["weight_edges"]
    got = conv.attention_weights(torch.as_tensor(we)).cpu().numpy()
    assert np.abs(got - g["attn_output_weights"]).max() < tol
    seen = torch.zeros(g["e"], dtype=torch.bool, device=dev)
    for chunk_ids, w in conv.attention_weights_chunks(chunk_edges=10):
        assert torch.equal(w, full[chunk_ids])
        seen[chunk_ids] = True
    assert bool(seen.all())
    with pytest.raises(IndexError):
        conv.attention_weights(torch.tensor([g["e"]], device=dev))
