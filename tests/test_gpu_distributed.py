"""Destination-partitioned path on >= 2 GPUs: every rank's rows of out / dX and the all-reduced parameter gradients
against the numpy ORACLE (not against the same kernels on one GPU), for both transports -- "peer" (ring-phased pushes
over CUDA-IPC windows, overlapped with the phases' compute) and "nccl" (serial all-to-all) -- and for the C4 (4 heads of
16) and C5 (8 heads of 8, 100 tokens) token shapes.  Skipped on a 1-GPU box; the same check is printed by every
``bench.py --gpus N`` line (``parity_check``), and tests/test_gpu_ring_phases.py runs the ring-phase kernels of a virtual
world on one GPU."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu
TOL_BF16 = 2e-2


def _worker(rank, world, port, spec, transport, ret):
    import numpy as np
    import torch.distributed as dist
    from ampnet_b200 import AMPConv, distributed as D
    from ampnet_b200 import functional as F_
    from oracle import cases, numpy_oracle
    import datetime
    import traceback
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev, timeout=datetime.timedelta(seconds=90))
    try:
        n, e, f, d, h = spec
        x, ei, p, d_out = cases.make_inputs(n, e, f, d, h, graph="skewed", seed=31)
        ref = numpy_oracle.backward(x, ei, p["in_proj_weight"], p["in_proj_bias"], p["out_proj_weight"], p["out_proj_bias"], h, d_out)
        conv = AMPConv(d, h, mode="bf16").to(dev)
        mha = conv.multi_head_attention
        with torch.no_grad():
            mha.in_proj_weight.copy_(torch.from_numpy(p["in_proj_weight"]))
            mha.in_proj_bias.copy_(torch.from_numpy(p["in_proj_bias"]))
            mha.out_proj.weight.copy_(torch.from_numpy(p["out_proj_weight"]))
            mha.out_proj.bias.copy_(torch.from_numpy(p["out_proj_bias"]))
        eit = torch.from_numpy(ei).to(dev)
        pg = D.PartitionedGraph(eit, n, world, rank)
        errs = {}
        for it in range(2):      # twice: the second step reuses the windows, flags and send slots of the first
            conv.zero_grad()
            xl = torch.from_numpy(x[pg.lo:pg.hi]).to(dev).requires_grad_(True)
            out_l = D.dist_amp_conv(xl, pg, mha.in_proj_weight, mha.in_proj_bias, mha.out_proj.weight, mha.out_proj.bias, h,
                                    transport=transport, key="layer")
            out_l.backward(torch.from_numpy(d_out[pg.lo:pg.hi]).to(dev))
            torch.cuda.synchronize()
            F_.check_status(sync=True)

            def rel(a, b, scale):
                return float(np.abs(a.astype(np.float64) - b).max() / max(float(np.abs(scale).max()), 1e-30))

            errs[f"out{it}"] = rel(out_l.detach().cpu().numpy(), ref["out"][pg.lo:pg.hi], ref["out"])
            errs[f"d_x{it}"] = rel(xl.grad.cpu().numpy(), ref["d_x"][pg.lo:pg.hi], ref["d_x"])
            for name, q in (("d_in_proj_weight", mha.in_proj_weight), ("d_in_proj_bias", mha.in_proj_bias),
                            ("d_out_proj_weight", mha.out_proj.weight), ("d_out_proj_bias", mha.out_proj.bias)):
                errs[f"{name}{it}"] = rel(q.grad.cpu().numpy(), ref[name], ref[name])
        ret[rank] = errs
        D.close_engines()
        dist.destroy_process_group()
    except BaseException:
        # a rank that fails must not wait for its peers (destroy_process_group would): die at once so that the parent sees it
        traceback.print_exc()
        os._exit(1)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
@pytest.mark.parametrize("transport", ["peer", "nccl"])
@pytest.mark.parametrize("spec", [(900, 7000, 128, 64, 4), (500, 3600, 100, 64, 8)])
def test_partitioned_matches_numpy_oracle(spec, transport):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    world = min(torch.cuda.device_count(), 4)
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, spec, transport, ret), nprocs=world, join=True)
    assert len(ret) == world
    for rank, errs in ret.items():
        for k, v in errs.items():
            assert v < TOL_BF16, (rank, k, v)
