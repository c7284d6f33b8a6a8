"""CPU checks of the boundary: the C-ABI library loads and exports every symbol include/ampconv.h
declares; argument validation that needs no GPU; the module mirrors the reference's interface."""
import ctypes

import numpy as np
import pytest
import torch

from ampnet_b200 import AMPConv, AMPConvV2, _lib
from conftest import load_golden


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _lib.declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), n
    assert lib.ampconv_abi_version() == 1
    assert lib.ampconv_strerror(0) == b"ok"
    assert b"edge_index" in lib.ampconv_strerror(-3)


def test_argument_validation_without_gpu():
    lib = _lib.load()
    # d % H != 0 and null pointers are rejected before any CUDA call
    rc = lib.ampconv_attn_fwd_f32(None, None, None, None, None, None, ctypes.c_int64(4), ctypes.c_int64(0),
                                  ctypes.c_int(2), ctypes.c_int(6), ctypes.c_int(4), None)
    assert rc == -1
    nbytes = ctypes.c_size_t(0)
    assert lib.ampconv_param_grad_workspace_bytes(ctypes.c_int(192), ctypes.c_int(64), ctypes.byref(nbytes)) == 0
    assert nbytes.value > 0
    with pytest.raises(_lib.AmpConvError):
        _lib.call("ampconv_qkv_proj_f32", None, None, None, None, _lib.i64(8), _lib.i32(4), None)


def test_module_mirrors_reference_interface():
    g = load_golden("tiny_generic")
    conv = AMPConv(g["d"], g["h"])
    assert sorted(conv.state_dict().keys()) == list(g["state_dict_keys"])
    assert conv.embed_dim == g["d"] and conv.num_heads == g["h"]
    assert conv.attn_output is None and conv.attn_output_weights is None
    mha = conv.multi_head_attention
    assert tuple(mha.in_proj_weight.shape) == (3 * g["d"], g["d"])
    assert tuple(mha.out_proj.weight.shape) == (g["d"], g["d"])
    assert float(mha.in_proj_bias.abs().max()) == 0.0 and float(mha.out_proj.bias.abs().max()) == 0.0
    assert issubclass(AMPConvV2, AMPConv)
    # a checkpoint written by the reference-style module loads into the mirror and back
    from oracle.torch_port import AMPConvPort
    port = AMPConvPort(g["d"], g["h"])
    conv.load_state_dict(port.state_dict())
    port.load_state_dict(conv.state_dict())
    assert torch.equal(conv.multi_head_attention.in_proj_weight, port.multi_head_attention.in_proj_weight)


def test_errors_are_loud_and_there_is_no_cpu_path():
    conv = AMPConv(4, 2)
    ei = torch.zeros(2, 3, dtype=torch.long)
    with pytest.raises(ValueError):
        conv(torch.zeros(5, 7), ei)          # width not a multiple of embed_dim
    with pytest.raises(TypeError):
        conv(torch.zeros(5, 8), ei)          # CPU tensor: no fallback
    with pytest.raises(AssertionError):   # same assertion as the reference's MHA (custom_multihead_attn.py:58-59)
        AMPConv(6, 4)


def test_mode_resolution_per_baseline_config():
    """Which kernel family serves which BASELINE config (host logic + the library's pure-C support predicate; no GPU)."""
    from ampnet_b200 import functional as F_
    assert F_.resolve_mode("auto", 128, 64, 4) == "bf16"        # C4: ogbn-arxiv token shape, head_dim 16
    assert F_.resolve_mode("bf16", 20, 64, 2) == "bf16"         # head_dim 32
    assert F_.resolve_mode("auto", 100, 64, 8) == "bf16"        # C5: ogbn-products token shape, head_dim 8 native (padding TMA boxes)
    assert F_.resolve_mode("bf16", 100, 64, 8) == "bf16"
    assert F_.resolve_mode("fp32", 100, 64, 8) == "fp32"        # the strict family is always available
    assert F_.resolve_mode("auto", 20, 128, 4) == "fp32"        # C2: embed 128 is outside the tensor-core family
    assert F_.resolve_mode("auto", 1433, 12, 3) == "fp32"       # C1
    assert F_.resolve_mode("auto", 2, 3, 1) == "fp32"           # C3
    assert F_.resolve_mode("auto", 129, 64, 4) == "fp32"        # more than 128 tokens per node
    for f, d, h in ((20, 128, 4), (129, 64, 8), (100, 64, 16)):
        with pytest.raises(ValueError):
            F_.resolve_mode("bf16", f, d, h)
    with pytest.raises(ValueError):
        F_.resolve_mode("tf32", 128, 64, 4)
