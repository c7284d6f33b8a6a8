#!/usr/bin/env python
"""bench.py -- AMPConv fwd+bwd edges/sec on B200 (BASELINE.json metric), driver contract.

    python bench.py --gpus 1 --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port)

A "step" is one forward + backward of one AMPConv layer over the whole synthetic graph with a fixed
upstream gradient d_out (out.backward(d_out)); the e2e leg additionally forms the scalar <out, d_out> and reads it
and the four parameter gradients back to the host.  At N=1 the workload is config C4 of SURVEY.md section 8 (synthetic
ogbn-arxiv shape: 169 343 nodes, 1 166 243 edges, 128 feature tokens, embed 64, 4 heads), the
largest configuration of BASELINE.json that fits one GPU.  Inputs are larger than L2 (x alone is
5.5 GB), so no explicit L2 flush is needed between iterations.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: N, E, F, d, H  
    "C4": dict(n=169343, e=1166243, f=128, d=64, h=4, desc="synthetic ogbn-arxiv shape, full graph"),
    "C2": dict(n=750, e=3000, f=20, d=128, h=4, desc="GraphSAINT-Cora subgraph shape"),
    "C3": dict(n=400, e=8400, f=2, d=3, h=1, desc="XOR graph shape"),
    "C4s": dict(n=16934, e=116624, f=128, d=64, h=4, desc="C4 token shape at 1/10 of the nodes and edges"),
    "C5": dict(n=2449029, e=61859140, f=100, d=64, h=8, desc="synthetic ogbn-products shape (multi-GPU only)"),
    "C5s": dict(n=122451, e=3092957, f=100, d=64, h=8, desc="ogbn-products token shape at 1/20 of the nodes and edges"),
    "C5h": dict(n=1224514, e=30929570, f=100, d=64, h=8, desc="ogbn-products token shape at 1/2 of the nodes and edges (8 GPUs)"),
}
METRIC = "AMPConv fwd+bwd edges/sec"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        p = json.load(open(path))
        return dict(hbm_gbs=float(p["hbm_gbs"]), bf16_tflops=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                    source="MEASURED_PEAKS.json")
    return dict(hbm_gbs=6650.0, bf16_tflops=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for name, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_problem(spec, graph_kind, seed=1234):
    """Synthetic inputs of SURVEY.md section 8(d): x ~ N(0,1), d_out ~ N(0,1), uniform or skewed graph."""
    from ampnet_b200.loader import make_graph
    return make_graph(graph_kind, spec["n"], spec["e"], seed=7 + seed)


def init_conv(conv, seed=0):
    g = torch.Generator().manual_seed(seed)
    mha = conv.multi_head_attention
    with torch.no_grad():
        d = conv.embed_dim
        bound = (6.0 / (4 * d)) ** 0.5
        mha.in_proj_weight.copy_((torch.rand(3 * d, d, generator=g) * 2 - 1) * bound)
        mha.in_proj_bias.copy_(0.1 * torch.randn(3 * d, generator=g))
        mha.out_proj.weight.copy_((torch.rand(d, d, generator=g) * 2 - 1) / d ** 0.5)
        mha.out_proj.bias.copy_(0.1 * torch.randn(d, generator=g))


# ------------------------------------------------------------------------------------------
# CPU arm: the reference's eager op chain (oracle/torch_port.py) on a bounded edge sample
# ------------------------------------------------------------------------------------------
def cpu_sample_run(spec, sample_edges, sample_nodes, reps, warmup):
    """The CPU arm (the one place bench.py executes oracle/): the reference's own AMPConv (oracle/_ref/amp_conv.py, the
    verbatim copy build() stages, behind the PyG stand-in) when present -- kind "reference" -- else the op-for-op port
    (oracle/torch_port.py, kind "port").  The edge sample is processed in chunks (each chunk is one forward + backward of
    the layer over the chunk's edges: every op on the path is per-edge or a linear scatter, so time is linear in edges)."""
    from ampnet_b200.loader import make_graph
    from oracle import reference_loader
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    f, d, h = spec["f"], spec["d"], spec["h"]
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(sample_nodes, f * d, generator=g)
    d_out = torch.randn(sample_nodes, f * d, generator=g)
    ei = torch.from_numpy(make_graph("uniform", sample_nodes, sample_edges, seed=7))
    chunk = max(64, min(sample_edges, int(2.5e8 // max(1, h * f * f))))   # keeps [chunk,H,F,F] fp32 near 1 GB
    if reference_loader.amp_conv_path() is not None:
        kind = "reference"
        conv = reference_loader.load_amp_conv_module().AMPConv(embed_dim=d, num_heads=h)

        def run():
            xin = x.detach().requires_grad_(True)
            for lo in range(0, sample_edges, chunk):
                out = conv(xin, ei[:, lo:lo + chunk])
                out.backward(d_out)
    else:
        kind = "port"
        from oracle.torch_port import AMPConvPort, fwd_bwd_chunked
        conv = AMPConvPort(d, h)

        def run():
            fwd_bwd_chunked(conv, x, ei, d_out, chunk)
    init_conv(conv)
    times = []
    for it in range(warmup + reps):
        conv.zero_grad()
        t0 = time.perf_counter()
        run()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return dict(times=times, cores=cores, chunk=chunk, kind=kind)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    spec = WORKLOADS[args.workload]
    se = min(spec["e"], args.cpu_sample_edges)
    sn = min(spec["n"], max(2, se // 2))
    r = cpu_sample_run(spec, se, sn, reps=args.steps, warmup=args.warmup)
    per_step = float(np.mean(r["times"]))
    value = se / per_step
    sample = (f"{se} edges over a {sn}-node slice at the {args.workload} token shape "
              f"(F={spec['f']}, d={spec['d']}, H={spec['h']}), fp32, edge-chunked by {r['chunk']}; edges/s is per-edge linear")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "edges/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {spec['desc']} (N={spec['n']}, E={spec['e']}, F={spec['f']}, "
                               f"d={spec['d']}, H={spec['h']}), one AMPConv layer fwd+bwd, {args.graph} graph seed 7",
                   "mode": "fp32 (the reference's AMPConv on the host cores: " +
                           ("oracle/_ref/amp_conv.py, verbatim, behind the PyG stand-in)" if r["kind"] == "reference"
                            else "op-for-op port oracle/torch_port.py)"), "sample": sample},
        "cpu_baseline": {"value": value, "unit": "edges/s", "cores": r["cores"], "kind": r["kind"], "sample": sample},
        "e2e": {"value": value, "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------
# Parity check printed with every GPU line: the layer that is about to be timed (same code path, same world size) on small
# seeded graphs against the CPU oracle (oracle/numpy_oracle.py, the checker -- never the thing measured).  At N > 1 every
# rank checks its own rows of out / dX and the all-reduced parameter gradients; the JSON line carries the worst error over
# ranks and the run exits non-zero when a case misses the bf16 bar of BASELINE.json (2e-2).
# ------------------------------------------------------------------------------------------
PARITY_CASES = [
    dict(name="c4_tokens_skewed", n=600, e=4200, f=128, d=64, h=4, graph="skewed"),
    dict(name="c5_tokens_skewed", n=480, e=3000, f=100, d=64, h=8, graph="skewed"),
]
PARITY_TOL = 2e-2


def parity_check(dev, world, rank, mode="bf16"):
    from ampnet_b200 import AMPConv
    from ampnet_b200.loader import make_inputs
    from oracle import numpy_oracle
    cases, worst = [], 0.0
    for c in PARITY_CASES:
        n, e, f, d, h = c["n"], c["e"], c["f"], c["d"], c["h"]
        x, ei, p, d_out = make_inputs(n, e, f, d, h, graph=c["graph"], seed=2024)
        ref = numpy_oracle.backward(x, ei, p["in_proj_weight"], p["in_proj_bias"], p["out_proj_weight"], p["out_proj_bias"], h, d_out)
        conv = AMPConv(d, h, mode=mode).to(dev)
        mha = conv.multi_head_attention
        with torch.no_grad():
            mha.in_proj_weight.copy_(torch.from_numpy(p["in_proj_weight"]))
            mha.in_proj_bias.copy_(torch.from_numpy(p["in_proj_bias"]))
            mha.out_proj.weight.copy_(torch.from_numpy(p["out_proj_weight"]))
            mha.out_proj.bias.copy_(torch.from_numpy(p["out_proj_bias"]))
        eit = torch.from_numpy(ei).to(dev)
        if world > 1:
            from ampnet_b200 import distributed as D
            pg = D.PartitionedGraph(eit, n, world, rank)
            lo, hi = pg.lo, pg.hi
            xl = torch.from_numpy(x[lo:hi]).to(dev).requires_grad_(True)
            out = D.dist_amp_conv(xl, pg, mha.in_proj_weight, mha.in_proj_bias, mha.out_proj.weight, mha.out_proj.bias, h,
                                  key=("parity", c["name"]))
        else:
            lo, hi = 0, n
            xl = torch.from_numpy(x).to(dev).requires_grad_(True)
            out = conv(xl, eit)
        out.backward(torch.from_numpy(d_out[lo:hi]).to(dev))
        torch.cuda.synchronize()

        def rel(a, b, scale):
            if a.size == 0:          # a rank may own no node at all (a hub holds more than 1 / world of the edges)
                return 0.0
            return float(np.abs(a.astype(np.float64) - b).max() / max(float(np.abs(scale).max()), 1e-30))

        errs = {"out": rel(out.detach().cpu().numpy(), ref["out"][lo:hi], ref["out"]),
                "d_x": rel(xl.grad.cpu().numpy(), ref["d_x"][lo:hi], ref["d_x"])}
        for key, q in (("d_in_proj_weight", mha.in_proj_weight), ("d_in_proj_bias", mha.in_proj_bias),
                       ("d_out_proj_weight", mha.out_proj.weight), ("d_out_proj_bias", mha.out_proj.bias)):
            errs[key] = rel(q.grad.cpu().numpy(), ref[key], ref[key])
        m = max(errs.values())
        if world > 1:
            t = torch.tensor([m], device=dev, dtype=torch.float64)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            m = float(t.item())
        worst = max(worst, m)
        cases.append({"case": c["name"], "shape": {k: c[k] for k in ("n", "e", "f", "d", "h")}, "graph": c["graph"],
                      "max_rel": m, "ok": bool(m < PARITY_TOL)})
    from ampnet_b200 import functional as F_
    F_.check_status(sync=True)
    return {"ok": all(c["ok"] for c in cases), "max_rel": worst, "tol": PARITY_TOL, "against": "oracle/numpy_oracle.py (fp64)",
            "path": ("dist_amp_conv over %d ranks" % world) if world > 1 else "AMPConv module, 1 GPU", "cases": cases}


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def algorithmic_bytes(spec, mode):
    """Per-launch algorithmic bytes of the three attention kernels (DESIGN.md 'Kernels')."""
    n, e, f, d, h = spec["n"], spec["e"], spec["f"], spec["d"], spec["h"]
    b = 4 if mode == "fp32" else 2
    tile = f * d * b
    stat = h * f * 4
    return {
        "attn_fwd": e * (2 * tile + 12 + stat) + n * (tile + f * d * 4),
        "attn_bwd_dq": e * (2 * tile + 12 + 2 * stat) + n * (2 * tile + f * d * b),          # dQ rows leave as bf16 in bf16 mode
        "attn_bwd_dkv": e * (2 * tile + 12 + 2 * stat) + n * (2 * tile + 2 * f * d * b),
    }


def ncu_traffic(workload, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel`, from the committed ncu --set full capture
    of this workload (profiles/ncu_traffic.json); None when no capture of this workload has been committed."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        t = json.load(open(path)).get(workload)
        return (int(t["bytes"][kernel]), t["source"]) if t and kernel in t["bytes"] else (None, None)
    except Exception:
        return None, None


def run_ours(args):
    from ampnet_b200 import AMPConv, _lib
    from ampnet_b200 import functional as F_
    from ampnet_b200.graph import Graph, clear_cache

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: ampnet_b200 has no CPU path")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=240))
    spec = dict(WORKLOADS[args.workload])
    n, e, f, d, h = spec["n"], spec["e"], spec["f"], spec["d"], spec["h"]

    if world > 1:
        return run_ours_partitioned(args, spec, world, rank, dev)
    edge_index_np = make_problem(spec, args.graph, seed=rank)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    x_host = torch.empty((n, f * d), dtype=torch.float32, pin_memory=True)
    x = torch.randn((n, f * d), generator=gen, device=dev)
    x_host.copy_(x)
    d_out = torch.randn((n, f * d), generator=gen, device=dev)
    ei_host = torch.from_numpy(edge_index_np).pin_memory()
    edge_index = ei_host.to(dev)
    conv = AMPConv(d, h, mode=args.mode).to(dev)
    init_conv(conv)
    params = list(conv.parameters())

    def step(xin, ei, want_loss=False):
        """One forward + backward of the layer with the upstream gradient d_out (the metric's unit of work).  The scalar
        loss <out, d_out> is only formed where a result has to travel to the host (the e2e leg)."""
        for p in params:
            p.grad = None
        xin.grad = None
        out = conv(xin, ei)
        loss = torch.dot(out.detach().view(-1), d_out.view(-1)) if want_loss else None
        out.backward(d_out)
        return loss

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    parity = None if args.no_parity_check else parity_check(dev, 1, 0, mode=args.mode if args.mode != "fp32" else "auto")
    x.requires_grad_(True)
    for _ in range(args.warmup):
        step(x, edge_index)
    barrier()
    launches0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        ev0.record()
        for _ in range(args.steps):
            step(x, edge_index)
        ev1.record()
        barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = (_lib.launch_count() - launches0) // max(1, args.steps)
    if world > 1:
        t = torch.tensor([ms_total], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * e / (ms_step * 1e-3)

    # ---- end to end through the public API with HOST buffers: every step uploads x + edge_index from pinned memory,
    #      builds the graph views, runs fwd + bwd and reads the loss and the four parameter gradients back.
    #      The uploads are double-buffered (ampnet_b200.loader.HostFeed): step i+1's inputs travel over PCIe on a copy
    #      stream while step i computes; all K uploads are inside the timed region, the first one fully exposed.
    #      The serial variant (upload, then compute, one stream) is timed as well and reported beside it.
    from ampnet_b200.loader import HostFeed
    e2e_steps = max(1, args.steps)
    grads_host = [torch.empty_like(p, device="cpu").pin_memory() for p in params]
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()
    del x
    torch.cuda.empty_cache()
    feed = HostFeed(dev)

    def e2e_compute(x_dev, ei_dev):
        clear_cache()                      # a new edge_index every step: the graph views are rebuilt
        xin = x_dev.detach().requires_grad_(True)
        loss = step(xin, ei_dev, want_loss=True)
        loss_host.copy_(loss.detach(), non_blocking=True)
        for gh, p in zip(grads_host, params):
            gh.copy_(p.grad, non_blocking=True)

    def e2e_run(k, overlap):
        feed.submit(x_host, ei_host)
        for i in range(k):
            x_dev, ei_dev = feed.get()
            if overlap and i + 1 < k:
                feed.submit(x_host, ei_host)
            e2e_compute(x_dev, ei_dev)
            feed.release()
            if not overlap and i + 1 < k:
                feed.copy_stream.wait_stream(torch.cuda.current_stream(dev))   # serial: upload only after this step's work
                feed.submit(x_host, ei_host)

    e2e_run(2, True)
    barrier()
    ev0.record()
    e2e_run(e2e_steps, True)
    ev1.record()
    barrier()
    e2e_ms = ev0.elapsed_time(ev1) / e2e_steps
    serial_steps = max(1, min(args.steps, 3))
    ev0.record()
    e2e_run(serial_steps, False)
    ev1.record()
    barrier()
    e2e_serial_ms = ev0.elapsed_time(ev1) / serial_steps
    if world > 1:
        t = torch.tensor([e2e_ms, e2e_serial_ms], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        e2e_ms, e2e_serial_ms = float(t[0].item()), float(t[1].item())
    h2d = x_host.numel() * 4 + ei_host.numel() * 8
    d2h = 4 + sum(p.numel() * 4 for p in params)
    feed.submit(x_host, ei_host)           # a resident copy of x for the per-kernel timings below
    x_dev, _ = feed.get()
    feed.release()

    # ---- per-kernel durations of the attention kernels (CUDA events on the launching stream)
    kern_ms = profile_attention_kernels(conv, x_dev, edge_index, d_out, args.mode, reps=max(2, min(args.steps, 5)))
    alg = algorithmic_bytes(spec, args.mode)
    pk = peaks()
    dominant = max(kern_ms, key=kern_ms.get)
    achieved = alg[dominant] / (kern_ms[dominant] * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic(args.workload, dominant)
    # the unit that actually binds these kernels is MUFU (exp2): H*F^2 exponentials per edge and pass at 16 / clk / SM
    exps = float(e) * h * f * f
    sm_clk = 148 * 1.965e9
    mufu = {k: exps / (v * 1e-3) / sm_clk / 16.0 for k, v in kern_ms.items()}
    # the three bounds of SURVEY 8(d) side by side: HBM (the dominant kernel, `frac`), tensor (whole step, algorithmic FLOPs
    # 14 F^2 d E + 24 F d^2 N against the sustained bf16 peak) and MUFU exp2 (per kernel)
    flops = 14.0 * f * f * d * e + 24.0 * f * d * d * n
    tflops = flops / (ms_step * 1e-3) / 1e12
    roofline = {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / pk["hbm_gbs"], "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": pk["source"], "kernel_ms": kern_ms, "algorithmic_bytes": alg,
                "hbm_frac_per_kernel": {k: alg[k] / (v * 1e-3) / 1e9 / pk["hbm_gbs"] for k, v in kern_ms.items()},
                "tensor": {"achieved_tflops": tflops, "peak_tflops": pk["bf16_tflops"], "frac": tflops / pk["bf16_tflops"],
                           "scope": "whole step, algorithmic FLOPs (zero-padding of head_dim 8 and F < 128 not counted)"},
                "binding": "none of the three pipes saturates: these kernels are bound by the latency between their pipeline stages "
                           "(ncu: XU 52-63 %, issue slots 50-64 %, tensor 8-27 %, DRAM 20-24 %; profiles/r02_ncu_attn_c4s_full.txt)",
                "mufu_frac": {"note": "exponentials per second relative to the MUFU exp2 issue peak (16 per clk per SM at 1.965 GHz, "
                                      "148 SMs: tools/mufu_probe.cu measures 15.98 per clk per SM), the busiest pipe of all three kernels; "
                                      "the forward and the dK/dV pass evaluate a quarter of their exponentials on the FMA pipe instead",
                              **mufu}}

    if rank != 0:
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        se = min(spec["e"], args.cpu_sample_edges)
        sn = min(spec["n"], max(2, se // 2))
        r = cpu_sample_run(spec, se, sn, reps=2, warmup=1)
        cpu = {"value": se / float(np.mean(r["times"])), "unit": "edges/s", "cores": r["cores"], "kind": r["kind"],
               "sample": f"{se} edges over a {sn}-node slice at the {args.workload} token shape, fp32, "
                         + ("the reference's own amp_conv.py (oracle/_ref, verbatim)" if r["kind"] == "reference"
                            else "reference op chain (oracle/torch_port.py)") + f", edge-chunked by {r['chunk']}"}
    line = {
        "metric": METRIC, "value": value, "unit": "edges/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32" if args.mode == "fp32" else "bf16", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {spec['desc']} (N={n}, E={e}, F={f}, d={d}, H={h}), one AMPConv layer "
                               f"fwd+bwd, {args.graph} graph seed 7", "mode": args.mode,
                   "l2": "inputs larger than L2 (x = %.1f GB); no flush" % (n * f * d * 4 / 1e9),
                   "parallelism": "1 GPU (bench.py --gpus N partitions this same graph over N GPUs: strong scaling)"},
        "node_updates_per_s": world * n / (ms_step * 1e-3),
        "clocks": clocks.summary(),
        "e2e": {"value": world * e / (e2e_ms * 1e-3), "unit": "edges/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "steps": e2e_steps,
                "serial_ms_per_step": e2e_serial_ms,
                "includes": "every step: H2D of x and edge_index from pinned memory, CSR build, fwd, bwd, D2H of loss and 4 "
                            "param grads; uploads double-buffered (ampnet_b200.loader.HostFeed): step i+1's upload overlaps "
                            "step i's compute, the first upload is exposed; serial_ms_per_step = same work without overlap"},
        "gpu_launches": int(launches),
        "roofline": roofline,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if parity is not None:
        line["parity_check"] = parity
    emit(line)
    if parity is not None and not parity["ok"]:
        raise SystemExit("parity check against the oracle failed: " + json.dumps(parity))


def run_ours_partitioned(args, spec, world, rank, dev):
    """N > 1: the SAME graph, destination-partitioned over the ranks (strong scaling).  Exchange step = ring-phased pushes
    over peer memory overlapped with the phases' compute (ampnet_b200/distributed.py, transport "peer"); NCCL carries one
    barrier per step and the all-reduce of the parameter gradients."""
    import torch.distributed as dist
    from ampnet_b200 import AMPConv, _lib, distributed as D
    from ampnet_b200 import functional as F_
    n, e, f, d, h = spec["n"], spec["e"], spec["f"], spec["d"], spec["h"]
    parity = None if args.no_parity_check else parity_check(dev, world, rank)
    D.close_engines()
    edge_index_np = make_problem(spec, args.graph, seed=0)            # identical on every rank
    ei_host = torch.from_numpy(edge_index_np).pin_memory()
    edge_index = ei_host.to(dev)
    pg = D.PartitionedGraph(edge_index, n, world, rank).build_plan()
    transport = D.default_transport(torch.empty(0, device=dev), pg)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    x = torch.randn((pg.n_local, f * d), generator=gen, device=dev)
    d_out = torch.randn((pg.n_local, f * d), generator=gen, device=dev)
    x_host = torch.empty_like(x, device="cpu").pin_memory()
    x_host.copy_(x)
    conv = AMPConv(d, h, mode="bf16").to(dev)
    init_conv(conv)
    mha = conv.multi_head_attention
    params = list(conv.parameters())

    def step(xin, want_loss=False):
        for p in params:
            p.grad = None
        xin.grad = None
        out = D.dist_amp_conv(xin, pg, mha.in_proj_weight, mha.in_proj_bias, mha.out_proj.weight, mha.out_proj.bias, h,
                              key="bench")
        loss = torch.dot(out.detach().view(-1), d_out.view(-1)) if want_loss else None
        out.backward(d_out)
        return loss

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    x.requires_grad_(True)
    for _ in range(args.warmup):
        step(x)
    barrier()
    D.phase_summary()            # drop the warm-up marks (AMPNET_B200_DIST_TIMING=1)
    launches0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(dev.index) as clocks:
        ev0.record()
        for _ in range(args.steps):
            step(x)
        ev1.record()
        barrier()
    t = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    launches = (_lib.launch_count() - launches0) // max(1, args.steps)
    phases = D.phase_summary()
    if phases and rank == 0:
        sys.stderr.write("phase ms (rank 0): " + json.dumps({k: round(v, 2) for k, v in phases.items()}) + "\n")

    # end to end, same definition as the 1-GPU line: every step uploads this rank's rows of x and the edge_index from
    # pinned memory (double-buffered, ampnet_b200.loader.HostFeed), rebuilds the rank's CSR views from the uploaded
    # edges, runs fwd + bwd with the exchange and reads the loss and the parameter gradients back.  Static across steps:
    # which destinations a rank owns and the halo plan (the index lists exchanged once per graph).
    from ampnet_b200.loader import HostFeed
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()
    grads_host = [torch.empty_like(p, device="cpu").pin_memory() for p in params]
    del x
    torch.cuda.empty_cache()
    feed = HostFeed(dev)

    def e2e_run(k):
        feed.submit(x_host, ei_host)
        for i in range(k):
            x_dev, ei_dev = feed.get()
            if i + 1 < k:
                feed.submit(x_host, ei_host)
            if getattr(pg, "phase_graphs", None) is not None:
                lei = torch.stack([pg.local_edge_index[0], ei_dev[1, pg.edge_ids] - pg.lo])   # destinations from the upload
                pg.local_edge_index = lei.contiguous()
                pg.phase_graphs = D.PhaseGraphs(pg, pg.phase_plan)
                for eng in D._engines.values():
                    if eng.pg is pg:
                        eng.pgs = pg.phase_graphs
            else:
                pg.graph = None
            xin = x_dev.detach().requires_grad_(True)
            loss = step(xin, want_loss=True)
            loss_host.copy_(loss.detach(), non_blocking=True)
            for gh, p in zip(grads_host, params):
                gh.copy_(p.grad, non_blocking=True)
            feed.release()

    e2e_run(2)
    barrier()
    e2e_steps = max(1, args.steps)
    ev0.record()
    e2e_run(e2e_steps)
    ev1.record()
    barrier()
    t = torch.tensor([ev0.elapsed_time(ev1) / e2e_steps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    sizes = torch.tensor([x_host.numel() * 4 + ei_host.numel() * 8, 4 + sum(p.numel() * 4 for p in params)], device=dev,
                         dtype=torch.float64)
    dist.all_reduce(sizes)
    F_.check_status(sync=True)
    free_b, total_b = torch.cuda.mem_get_info(dev)     # includes the peer windows, which the library allocates itself
    mem = torch.tensor([torch.cuda.max_memory_allocated(dev) / 2 ** 30, (total_b - free_b) / 2 ** 30], device=dev, dtype=torch.float64)
    dist.all_reduce(mem, op=dist.ReduceOp.MAX)
    if rank != 0:
        return
    pk = peaks()
    # whole-step bounds: algorithmic FLOPs 14 F^2 d E + 24 F d^2 N against the sustained bf16 peak of all GPUs,
    # exponentials 3 H F^2 E against the MUFU issue peak (16 / clk / SM at 1.965 GHz)
    flops = 14.0 * f * f * d * e + 24.0 * f * d * d * n
    exps = 3.0 * h * f * f * e
    sec = ms_step * 1e-3
    line = {
        "metric": METRIC, "value": e / sec, "unit": "edges/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {spec['desc']} (N={n}, E={e}, F={f}, d={d}, H={h}), one AMPConv layer "
                               f"fwd+bwd, {args.graph} graph seed 7", "mode": "bf16",
                   "l2": "inputs larger than L2; no flush",
                   "parallelism": f"dst-partitioned over {world} GPUs, transport {transport}: "
                                  + ("ring-phased pushes of the referenced K|V rows into the peers' windows (CUDA IPC, copy "
                                     "engines over NVLink) overlapped with the phases' attention; dK|dV blocks pushed back per "
                                     "phase, fixed-order adds; NCCL: one barrier per step + all-reduce of the parameter gradients"
                                     if transport == "peer" else
                                     "halo exchange of referenced K|V rows (bf16 all-to-all), dK|dV halo rows back to their "
                                     "owners (bf16 all-to-all, fixed-order adds), all-reduce parameter gradients (NCCL)")},
        "node_updates_per_s": n / sec,
        "clocks": clocks.summary(),
        "e2e": {"value": e / (e2e_ms * 1e-3), "unit": "edges/s", "h2d_bytes_per_step": int(sizes[0].item()),
                "d2h_bytes_per_step": int(sizes[1].item()), "ms_per_step": e2e_ms, "steps": e2e_steps,
                "includes": "every step, every rank: H2D of the rank's rows of x and of edge_index from pinned memory "
                            "(double-buffered, step i+1's upload overlaps step i's compute, first upload exposed), rebuild of "
                            "the rank's CSR views, fwd, bwd (with the exchange), D2H of loss and 4 param grads; static: the "
                            "destination partition and the halo plan"},
        "gpu_launches": int(launches),
        "max_memory_gib_per_gpu": {"torch_allocator_peak": float(mem[0].item()), "device_in_use_at_end": float(mem[1].item())},
        "roofline": {"bound": "hbm", "kernel": "whole step (see the N=1 line for per-kernel numbers)", "achieved": None,
                     "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": None, "traffic": None, "peak_source": pk["source"],
                     "tensor_frac": flops / sec / 1e12 / (pk["bf16_tflops"] * world),
                     "mufu_frac": exps / sec / (148 * 1.965e9 * 16.0 * world)},
    }
    if phases:
        line["phase_ms_rank0"] = {k: round(v, 3) for k, v in phases.items()}
    if parity is not None:
        line["parity_check"] = parity
    emit(line)
    if parity is not None and not parity["ok"]:
        raise SystemExit("parity check against the oracle failed: " + json.dumps(parity))


def profile_attention_kernels(conv, x, edge_index, d_out, mode, reps):
    """Times each attention kernel alone with CUDA events on the stream it is launched on."""
    from ampnet_b200 import _lib
    from ampnet_b200 import functional as F_
    from ampnet_b200.graph import get_graph
    dev = x.device
    g = get_graph(edge_index, x.shape[0])
    mha = conv.multi_head_attention
    with torch.no_grad():
        stages = F_.profile_stages(x.detach(), g, mha.in_proj_weight, mha.in_proj_bias, mha.out_proj.weight,
                                   mha.out_proj.bias, conv.num_heads, d_out, mode, reps)
    return stages


_REAL_STDOUT = None


def _guard_stdout():
    """Only the JSON line may reach stdout: NCCL and friends print banners to fd 1, so route fd 1 to stderr and keep a
    private handle on the real stdout for emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _guard_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C4", choices=sorted(WORKLOADS))
    ap.add_argument("--graph", default="uniform", choices=["uniform", "skewed"])
    ap.add_argument("--mode", default=os.environ.get("AMPNET_B200_BENCH_MODE", "bf16"))
    ap.add_argument("--cpu-sample-edges", type=int, default=4096)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
