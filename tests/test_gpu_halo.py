"""Single-GPU coverage of the multi-GPU building blocks (the 2-GPU equality test in test_gpu_distributed.py is skipped on
a one-GPU box): the deterministic halo-add kernel, the dK|dV kernel variant that writes halo rows as bf16 for the wire,
and the whole partitioned layer at world size 1 (NCCL process group of one rank) against the plain layer."""
import ctypes
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_halo_add_bf16_matches_index_add_in_sender_order():
    from ampnet_b200 import _lib
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(3)
    n_local, row_elems = 50, 256
    # receive buffer: three sender blocks; targets repeat across blocks, unique inside a block
    blocks = [torch.randperm(n_local, generator=g, device=dev)[:c] for c in (50, 40, 17)]
    send_idx = torch.cat(blocks)
    n_recv = send_idx.numel()
    recv = torch.randn(n_recv, row_elems, generator=g, device=dev).to(torch.bfloat16)
    acc = torch.randn(n_local, row_elems, generator=g, device=dev)
    ref = acc.clone().index_add_(0, send_idx, recv.float())
    order = torch.sort(send_idx, stable=True)
    tgt, counts = torch.unique_consecutive(order.values, return_counts=True)
    rowptr = torch.zeros(tgt.numel() + 1, dtype=torch.int64, device=dev)
    rowptr[1:] = torch.cumsum(counts, 0)
    _lib.call("ampconv_halo_add_bf16", recv, tgt.to(torch.int32), rowptr.to(torch.int32), order.indices.to(torch.int32), acc,
              _lib.i64(tgt.numel()), _lib.i64(row_elems), _lib.stream_ptr(torch.cuda.current_stream(dev)))
    torch.cuda.synchronize()
    assert torch.allclose(acc, ref, rtol=1e-6, atol=1e-6)
    untouched = torch.ones(n_local, dtype=torch.bool, device=dev)
    untouched[tgt] = False
    assert torch.equal(acc[untouched], ref[untouched])


@pytest.mark.parametrize("f,h", [(128, 4), (100, 4), (40, 2)])
def test_dkv_halo_variant_equals_fp32_variant(f, h):
    """own sources: identical fp32 rows; halo sources: the same rows rounded to bf16."""
    from ampnet_b200 import _lib
    from ampnet_b200.distributed import BipartiteGraph
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(11)
    n_dst, n_halo, e, d = 120, 90, 1500, 64
    n_kv = n_dst + n_halo
    src = torch.randint(0, n_kv, (e,), generator=g, device=dev)
    src[:n_halo] = torch.arange(n_dst, n_kv, device=dev)          # every halo source has at least one edge
    dst = torch.randint(0, n_dst, (e,), generator=g, device=dev)
    bg = BipartiteGraph(torch.stack([src, dst]), n_dst, n_kv)
    q = (0.3 * torch.randn(n_dst * f, d, generator=g, device=dev)).to(torch.bfloat16)
    k = torch.randn(n_kv * f, d, generator=g, device=dev).to(torch.bfloat16)
    v = torch.randn(n_kv * f, d, generator=g, device=dev).to(torch.bfloat16)
    d_agg = torch.randn(n_dst * f, d, generator=g, device=dev).to(torch.bfloat16)
    fs = (f + 3) // 4 * 4
    agg = torch.empty(n_dst * f, d, device=dev)
    lse2 = torch.zeros(e, h, fs, device=dev)
    delta = torch.zeros_like(lse2)
    d_q = torch.empty(n_dst * f, d, device=dev)
    ws = torch.zeros(64, dtype=torch.int32, device=dev)
    st = _lib.stream_ptr(torch.cuda.current_stream(dev))
    tail = (_lib.i64(n_dst), _lib.i64(n_kv), _lib.i64(e), _lib.i32(f), _lib.i32(d), _lib.i32(h), ws, _lib.size_t(256), st)
    _lib.call("ampconv_attn_fwd_bf16_part", q, k, v, bg.dst_rowptr, bg.dst_src, bg.inv_deg, bg.order_dst, agg, lse2, *tail)
    _lib.call("ampconv_attn_bwd_dq_bf16_part", q, k, v, d_agg, lse2, bg.dst_rowptr, bg.dst_src, bg.order_dst, d_q, delta, *tail)
    full = torch.empty(n_kv * f, 2 * d, device=dev)
    _lib.call("ampconv_attn_bwd_dkv_bf16_part", q, k, v, d_agg, lse2, delta, bg.src_rowptr, bg.src_dst, bg.src_pos,
              bg.order_src, full, *tail)
    own = torch.empty(n_dst * f, 2 * d, device=dev)
    halo = torch.empty(n_halo * f, 2 * d, dtype=torch.bfloat16, device=dev)
    _lib.call("ampconv_attn_bwd_dkv_bf16_halo", q, k, v, d_agg, lse2, delta, bg.src_rowptr, bg.src_dst, bg.src_pos,
              bg.order_src, own, halo, _lib.i64(n_dst), _lib.i64(n_dst), _lib.i64(n_kv), _lib.i64(e), _lib.i32(f), _lib.i32(d),
              _lib.i32(h), ws, _lib.size_t(256), st)
    status = ctypes.c_int(0)
    _lib.call("ampconv_bf16_status", ws, ctypes.byref(status), st)
    assert status.value == 0
    assert torch.isfinite(full).all()
    assert torch.equal(own, full[:n_dst * f])
    assert torch.equal(halo, full[n_dst * f:].to(torch.bfloat16))


def test_partitioned_layer_at_world_size_one_matches_plain_layer():
    import torch.distributed as dist
    from ampnet_b200 import AMPConv, distributed as D
    from oracle import cases
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1, device_id=dev)
    try:
        n, e, f, d, h = 400, 3000, 96, 64, 4
        x, ei, p, d_out = cases.make_inputs(n, e, f, d, h, graph="skewed", seed=5)
        conv = AMPConv(d, h, mode="bf16").to(dev)
        mha = conv.multi_head_attention
        with torch.no_grad():
            mha.in_proj_weight.copy_(torch.from_numpy(p["in_proj_weight"]))
            mha.in_proj_bias.copy_(torch.from_numpy(p["in_proj_bias"]))
            mha.out_proj.weight.copy_(torch.from_numpy(p["out_proj_weight"]))
            mha.out_proj.bias.copy_(torch.from_numpy(p["out_proj_bias"]))
        eit = torch.from_numpy(ei).to(dev)
        go = torch.from_numpy(d_out).to(dev)
        xt = torch.from_numpy(x).to(dev).requires_grad_(True)
        out = conv(xt, eit)
        out.backward(go)
        ref = [out.detach().clone(), xt.grad.clone()] + [q.grad.clone() for q in conv.parameters()]
        conv.zero_grad()
        pg = D.PartitionedGraph(eit, n, 1, 0)
        assert pg.n_halo == 0 and pg.n_local == n
        xl = torch.from_numpy(x).to(dev).requires_grad_(True)
        out_l = D.dist_amp_conv(xl, pg, mha.in_proj_weight, mha.in_proj_bias, mha.out_proj.weight, mha.out_proj.bias, h)
        out_l.backward(go)
        got = [out_l.detach(), xl.grad] + [q.grad for q in conv.parameters()]
        for a, b in zip(got, ref):
            assert float((a - b).abs().max()) <= 1e-4 * float(b.abs().max()) + 1e-7
    finally:
        dist.destroy_process_group()
