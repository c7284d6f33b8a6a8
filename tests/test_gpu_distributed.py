"""Destination-partitioned path on >= 2 GPUs (NCCL): every rank's rows of out / dX and the all-reduced parameter
gradients must equal the single-GPU bf16-mode result computed from the same kernels."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, spec, ret):
    import torch.distributed as dist
    from ampnet_b200 import AMPConv, distributed as D
    from oracle import cases
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        n, e, f, d, h = spec
        x, ei, p, d_out = cases.make_inputs(n, e, f, d, h, graph="skewed", seed=31)
        conv = AMPConv(d, h, mode="bf16").to(dev)
        mha = conv.multi_head_attention
        with torch.no_grad():
            mha.in_proj_weight.copy_(torch.from_numpy(p["in_proj_weight"]))
            mha.in_proj_bias.copy_(torch.from_numpy(p["in_proj_bias"]))
            mha.out_proj.weight.copy_(torch.from_numpy(p["out_proj_weight"]))
            mha.out_proj.bias.copy_(torch.from_numpy(p["out_proj_bias"]))
        # single-GPU result (every rank computes it; same kernels, whole graph)
        xt = torch.from_numpy(x).to(dev).requires_grad_(True)
        eit = torch.from_numpy(ei).to(dev)
        out = conv(xt, eit)
        (out * torch.from_numpy(d_out).to(dev)).sum().backward()
        ref_grads = [q.grad.clone() for q in conv.parameters()]
        conv.zero_grad()
        # partitioned
        pg = D.PartitionedGraph(eit, n, world, rank)
        xl = torch.from_numpy(x[pg.lo:pg.hi]).to(dev).requires_grad_(True)
        out_l = D.dist_amp_conv(xl, pg, mha.in_proj_weight, mha.in_proj_bias, mha.out_proj.weight, mha.out_proj.bias, h)
        (out_l * torch.from_numpy(d_out[pg.lo:pg.hi]).to(dev)).sum().backward()
        torch.cuda.synchronize()

        def rel(a, b):
            return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))

        errs = {"out": rel(out_l, out[pg.lo:pg.hi].detach()), "d_x": rel(xl.grad, xt.grad[pg.lo:pg.hi])}
        for name, q, g0 in zip(["w_in", "b_in", "w_out", "b_out"], conv.parameters(), ref_grads):
            errs[name] = rel(q.grad, g0)
        ret[rank] = errs
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_partitioned_matches_single_gpu():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    world = 2
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, (900, 7000, 128, 64, 4), ret), nprocs=world, join=True)
    assert len(ret) == world
    for rank, errs in ret.items():
        for k, v in errs.items():
            assert v < 5e-3, (rank, k, v)
