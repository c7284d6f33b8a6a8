"""autograd glue between torch tensors and the C ABI (include/ampconv.h).

``amp_conv(x, graph, in_proj_weight, in_proj_bias, out_proj_weight, out_proj_bias, num_heads)``
computes what the reference's ``AMPConv.forward`` computes (``src/ampnet/conv/amp_conv.py:24-51``)
and is differentiable w.r.t. ``x`` and the four parameters.  All launches go to the current CUDA
stream; nothing here synchronises.
"""
import ctypes

import torch

from . import _lib

MODES = ("fp32", "bf16", "auto")
LOG2E = 1.4426950408889634
LN2 = 0.6931471805599453


def bf16_supported(f, d, h):
    lib = _lib.load()
    return bool(lib.ampconv_attn_bf16_supported(int(f), int(d), int(h)))


def resolve_mode(mode, f, d, h):
    """-> "fp32" | "bf16".  The tcgen05 family covers embed 64 with head_dim 32, 16 and 8 (the latter -- the ogbn-products shape --
    as work items (node, head group) whose tiles are zero-padded to 16 columns per head on the way into shared memory)."""
    if mode not in MODES:
        raise ValueError(f"unknown mode {mode!r}; available: {MODES}")
    if mode == "fp32":
        return mode
    if bf16_supported(f, d, h):
        return "bf16"
    if mode == "bf16":
        raise ValueError(f"mode='bf16' (tcgen05 kernels) does not cover F={f}, embed_dim={d}, num_heads={h}; "
                         "use mode='auto' or 'fp32'")
    return "fp32"


def _stream(dev):
    return _lib.stream_ptr(torch.cuda.current_stream(dev))


# ---------------------------------------------------------------------------------------------------------------------
# Status word of the tcgen05 kernels.  Every mbarrier wait in them is bounded; on a time-out the kernel records
# `code | CTA << 16` in its workspace and drains instead of hanging the GPU -- the output of that launch is then garbage.
# The product path therefore copies the word to pinned host memory at the end of every forward / backward (asynchronously,
# no synchronisation) and raises at the next layer call (or at ``check_status(sync=True)``) if any copy came back non-zero.
# ---------------------------------------------------------------------------------------------------------------------
_pending_status = []     # (pinned int32[1], event, label)


class KernelProtocolError(RuntimeError):
    pass


def _post_status(ws, label):
    host = torch.zeros(1, dtype=torch.int32).pin_memory()
    host.copy_(ws[1:2], non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    _pending_status.append((host, ev, label))


def check_status(sync=False):
    """Raises KernelProtocolError if a tcgen05 launch reported a pipeline time-out.  sync=False only looks at status copies
    that have already arrived (free); sync=True waits for all of them."""
    keep, bad = [], None
    for host, ev, label in _pending_status:
        if sync:
            ev.synchronize()
        if not ev.query():
            keep.append((host, ev, label))
            continue
        v = int(host.item())
        if v != 0 and bad is None:
            bad = (v, label)
    _pending_status[:] = keep
    if bad is not None:
        raise KernelProtocolError(f"tcgen05 kernel pipeline time-out in {bad[1]}: wait id {bad[0] & 0xffff} in CTA {bad[0] >> 16}; "
                                  "the results of that call are invalid")


def _param_grad_ws(out_dim, in_dim, dev):
    nbytes = ctypes.c_size_t(0)
    _lib.call("ampconv_param_grad_workspace_bytes", _lib.i32(out_dim), _lib.i32(in_dim), ctypes.byref(nbytes))
    return torch.empty(nbytes.value, dtype=torch.uint8, device=dev)


class _Saved:
    """What the last forward of a layer keeps for backward and for the lazy side outputs."""

    def __init__(self, mode, graph, shape, qkv, agg, lse, bf16=None, inputs=None):
        self.mode, self.graph, self.shape = mode, graph, shape
        self.qkv, self.agg, self.lse = qkv, agg, lse
        self.bf16 = bf16          # (q, k, v, lse2, workspace) of the tensor-core family
        self.inputs = inputs      # (x, w_in, b_in): lets the fp32 views be rebuilt lazily

    def ensure_fp32_views(self):
        """fp32 qkv and natural-log lse (needed by the fp32 kernels) for a forward that ran in bf16 mode."""
        if self.qkv is None:
            x, w_in, b_in = self.inputs
            n, e, f, d, h = self.shape
            dev = x.device
            self.qkv = torch.empty((n * f, 3 * d), dtype=torch.float32, device=dev)
            with torch.cuda.device(dev):
                _lib.call("ampconv_qkv_proj_f32", x, w_in, b_in, self.qkv, _lib.i64(n * f), _lib.i32(d), _stream(dev))
            self.lse = (self.bf16[3][:, :, :f] * LN2).contiguous()
        return self


def _forward_fp32(x, graph, w_in, b_in, w_out, b_out, num_heads):
    n, width = x.shape
    d = w_in.shape[1]
    f = width // d
    dev = x.device
    e = graph.num_edges
    st = _stream(dev)
    rows = n * f
    qkv = torch.empty((rows, 3 * d), dtype=torch.float32, device=dev)
    agg = torch.empty((rows, d), dtype=torch.float32, device=dev)
    lse = torch.empty((e, num_heads, f), dtype=torch.float32, device=dev)
    out = torch.empty((n, width), dtype=torch.float32, device=dev)
    _lib.call("ampconv_qkv_proj_f32", x, w_in, b_in, qkv, _lib.i64(rows), _lib.i32(d), st)
    _lib.call("ampconv_attn_fwd_f32", qkv, graph.dst_rowptr, graph.dst_src, graph.inv_deg, agg, lse,
              _lib.i64(n), _lib.i64(e), _lib.i32(f), _lib.i32(d), _lib.i32(num_heads), st)
    _lib.call("ampconv_out_proj_f32", agg, w_out, b_out, graph.has_in, out,
              _lib.i64(n), _lib.i32(f), _lib.i32(d), st)
    return out, _Saved("fp32", graph, (n, e, f, d, num_heads), qkv, agg, lse)


def _forward_bf16(x, graph, w_in, b_in, w_out, b_out, num_heads):
    """tcgen05 forward: node-level projection to bf16 Q'/K/V, fused attention + mean on tensor cores."""
    n, width = x.shape
    d = w_in.shape[1]
    f = width // d
    dev = x.device
    e = graph.num_edges
    st = _stream(dev)
    rows = n * f
    hd = d // num_heads
    q = torch.empty((rows, d), dtype=torch.bfloat16, device=dev)
    k = torch.empty_like(q)
    v = torch.empty_like(q)
    agg = torch.empty((rows, d), dtype=torch.float32, device=dev)
    lse2 = torch.empty((e, num_heads, (f + 3) // 4 * 4), dtype=torch.float32, device=dev)
    out = torch.empty((n, width), dtype=torch.float32, device=dev)
    ws = torch.zeros(64, dtype=torch.int32, device=dev)
    _lib.call("ampconv_qkv_proj_tc", x, w_in, b_in, q, k, v, _lib.i64(rows), _lib.i32(d),
              _lib.f32(LOG2E / hd ** 0.5), ws, st)
    _lib.call("ampconv_attn_fwd_bf16", q, k, v, graph.dst_rowptr, graph.dst_src, graph.inv_deg, graph.order_dst, agg, lse2,
              _lib.i64(n), _lib.i64(e), _lib.i32(f), _lib.i32(d), _lib.i32(num_heads),
              ws, _lib.size_t(ws.numel() * 4), st)
    _lib.call("ampconv_out_proj_tc", agg, w_out, b_out, graph.has_in, out,
              _lib.i64(n), _lib.i32(f), _lib.i32(d), ws, st)
    _post_status(ws, "AMPConv forward (bf16)")
    saved = _Saved("bf16", graph, (n, e, f, d, num_heads), None, agg, None, bf16=(q, k, v, lse2, ws),
                   inputs=(x, w_in.detach().clone(), b_in.detach().clone()))
    return out, saved


def _backward_fp32(saved, x, w_in, w_out, d_out):
    saved.ensure_fp32_views()
    n, e, f, d, h = saved.shape
    g = saved.graph
    dev = x.device
    st = _stream(dev)
    rows = n * f
    d_out = d_out.contiguous()
    d_agg = torch.empty((rows, d), dtype=torch.float32, device=dev)
    d_w_out = torch.empty_like(w_out)
    d_b_out = torch.empty(d, dtype=torch.float32, device=dev)
    ws = _param_grad_ws(3 * d, d, dev)
    _lib.call("ampconv_out_proj_bwd_f32", d_out, saved.agg, w_out, g.inv_deg, g.has_in,
              d_agg, d_w_out, d_b_out, _lib.i64(n), _lib.i32(f), _lib.i32(d),
              ws, _lib.size_t(ws.numel()), st)
    d_qkv = torch.empty((rows, 3 * d), dtype=torch.float32, device=dev)
    delta = torch.empty_like(saved.lse)
    dims = (_lib.i64(n), _lib.i64(e), _lib.i32(f), _lib.i32(d), _lib.i32(h), st)
    _lib.call("ampconv_attn_bwd_dq_f32", saved.qkv, d_agg, saved.lse, g.dst_rowptr, g.dst_src, d_qkv, delta, *dims)
    _lib.call("ampconv_attn_bwd_dkv_f32", saved.qkv, d_agg, saved.lse, delta, g.src_rowptr, g.src_dst, g.src_pos,
              d_qkv, *dims)
    d_x = torch.empty_like(x)
    d_w_in = torch.empty_like(w_in)
    d_b_in = torch.empty(3 * d, dtype=torch.float32, device=dev)
    _lib.call("ampconv_qkv_proj_bwd_f32", x, d_qkv, w_in, d_x, d_w_in, d_b_in,
              _lib.i64(rows), _lib.i32(d), ws, _lib.size_t(ws.numel()), st)
    return d_x, d_w_in, d_b_in, d_w_out, d_b_out


def _backward_bf16(saved, x, w_in, w_out, d_out):
    """tcgen05 backward: dO tile in bf16, two recompute kernels (dQ by destination, dK/dV by source)."""
    n, e, f, d, h = saved.shape
    g = saved.graph
    dev = x.device
    st = _stream(dev)
    rows = n * f
    q, k, v, lse2, bws = saved.bf16
    d_out = d_out.contiguous()
    d_agg = torch.empty((rows, d), dtype=torch.bfloat16, device=dev)
    d_w_out = torch.empty_like(w_out)
    d_b_out = torch.empty(d, dtype=torch.float32, device=dev)
    ws = _param_grad_ws(3 * d, d, dev)
    _lib.call("ampconv_out_proj_bwd_input_tc", d_out, w_out, g.inv_deg, d_agg,
              _lib.i64(n), _lib.i32(f), _lib.i32(d), bws, st)
    _lib.call("ampconv_out_proj_bwd_params_tc", d_out, saved.agg, g.has_in, d_w_out, d_b_out,
              _lib.i64(n), _lib.i32(f), _lib.i32(d), ws, _lib.size_t(ws.numel()), bws, st)
    # gradient rows dQ | dK | dV as bf16: their consumers feed bf16 operands to the tensor cores anyway
    d_qkv = torch.empty((rows, 3 * d), dtype=torch.bfloat16, device=dev)
    delta = torch.empty_like(lse2)
    tail = (_lib.i64(n), _lib.i64(e), _lib.i32(f), _lib.i32(d), _lib.i32(h), bws, _lib.size_t(bws.numel() * 4), st)
    _lib.call("ampconv_attn_bwd_dq_bf16_h", q, k, v, d_agg, lse2, g.dst_rowptr, g.dst_src, g.order_dst, d_qkv, delta, *tail)
    _lib.call("ampconv_attn_bwd_dkv_bf16_h", q, k, v, d_agg, lse2, delta, g.src_rowptr, g.src_dst, g.src_pos,
              g.order_src, d_qkv, *tail)
    d_x = torch.empty_like(x)
    d_w_in = torch.empty_like(w_in)
    d_b_in = torch.empty(3 * d, dtype=torch.float32, device=dev)
    _lib.call("ampconv_qkv_proj_bwd_input_tc_h", d_qkv, w_in, d_x, _lib.i64(rows), _lib.i32(d), bws, st)
    _lib.call("ampconv_qkv_proj_bwd_params_tc_h", x, d_qkv, d_w_in, d_b_in,
              _lib.i64(rows), _lib.i32(d), ws, _lib.size_t(ws.numel()), bws, st)
    _post_status(bws, "AMPConv backward (bf16)")
    return d_x, d_w_in, d_b_in, d_w_out, d_b_out


BF16_BACKWARD = "tcgen05"   # "fp32" routes the backward of a bf16-mode forward through the fp32 kernels (debug aid)


class _AMPConvFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w_in, b_in, w_out, b_out, graph, num_heads, mode, holder):
        with torch.cuda.device(x.device):
            fwd = _forward_bf16 if mode == "bf16" else _forward_fp32
            out, saved = fwd(x, graph, w_in, b_in, w_out, b_out, num_heads)
        ctx.save_for_backward(x, w_in, w_out)
        ctx.saved_state = saved
        if holder is not None:
            holder["saved"] = saved
            holder["params"] = (w_out.detach().clone(), b_out.detach().clone())   # snapshot: the lazy side outputs must not see a later optimiser step
        return out

    @staticmethod
    def backward(ctx, d_out):
        x, w_in, w_out = ctx.saved_tensors
        with torch.cuda.device(x.device):
            bwd = _backward_fp32
            if BF16_BACKWARD == "tcgen05":
                bwd = _backward_bf16 if ctx.saved_state.mode == "bf16" else _backward_fp32
            d_x, d_w_in, d_b_in, d_w_out, d_b_out = bwd(ctx.saved_state, x, w_in, w_out, d_out)
        return d_x, d_w_in, d_b_in, d_w_out, d_b_out, None, None, None, None


def check_inputs(x, edge_index, embed_dim, num_heads):
    if x.dim() != 2:
        raise ValueError("x must have shape [N, F * embed_dim]")
    if x.shape[1] % embed_dim != 0:
        # the reference prints "Error, invalid configuration" and then fails in reshape (amp_conv.py:32-36)
        raise ValueError(f"invalid configuration: x.shape[1]={x.shape[1]} is not a multiple of embed_dim={embed_dim}")
    if embed_dim % num_heads != 0:
        raise ValueError("embed_dim must be divisible by num_heads")
    if not x.is_cuda:
        raise TypeError("ampnet_b200.AMPConv has no CPU path: x must be a CUDA tensor "
                        "(the CPU oracle is oracle/torch_port.py and is test infrastructure only)")
    if x.dtype != torch.float32:
        raise TypeError("x must be float32, like the reference's activations")
    if edge_index.device != x.device:
        raise TypeError("x and edge_index must be on the same device")


def check_params(x, w_in, b_in, w_out, b_out):
    """The C ABI takes raw fp32 device pointers: a module after .half() / .double(), or with parameters on another device,
    must raise here instead of being reinterpreted."""
    d = w_in.shape[1] if w_in.dim() == 2 else -1
    want = {"in_proj_weight": (w_in, (3 * d, d)), "in_proj_bias": (b_in, (3 * d,)),
            "out_proj.weight": (w_out, (d, d)), "out_proj.bias": (b_out, (d,))}
    for name, (t, shape) in want.items():
        if t.dtype != torch.float32:
            raise TypeError(f"{name} must be float32 (got {t.dtype}): the kernels read fp32 parameters")
        if t.device != x.device:
            raise TypeError(f"{name} is on {t.device}, x on {x.device}")
        if tuple(t.shape) != shape:
            raise ValueError(f"{name} has shape {tuple(t.shape)}, expected {shape}")


def amp_conv(x, graph, w_in, b_in, w_out, b_out, num_heads, mode="auto", holder=None):
    check_params(x, w_in, b_in, w_out, b_out)
    check_status()
    mode = resolve_mode(mode, x.shape[1] // w_in.shape[1], w_in.shape[1], num_heads)
    return _AMPConvFunction.apply(x.contiguous(), w_in.contiguous(), b_in.contiguous(), w_out.contiguous(),
                                  b_out.contiguous(), graph, num_heads, mode, holder)


def attention_weights(saved, edge_ids=None):
    """Head-averaged coefficients in original edge order (``attn_output_weights``): [E, F, F], or [len(edge_ids), F, F] for
    the listed edge ids (columns of ``edge_index``) -- the chunked form: at the ogbn-arxiv shape all of [E, F, F] is 76 GB."""
    saved.ensure_fp32_views()
    n, e, f, d, h = saved.shape
    g = saved.graph
    dev = saved.qkv.device
    with torch.cuda.device(dev):
        if edge_ids is None:
            w = torch.empty((e, f, f), dtype=torch.float32, device=dev)
            _lib.call("ampconv_attn_weights_f32", saved.qkv, saved.lse, g.dst_rowptr, g.dst_src, g.dst_eid, w,
                      _lib.i64(n), _lib.i64(e), _lib.i32(f), _lib.i32(d), _lib.i32(h), _stream(dev))
            return w
        edge_ids = torch.as_tensor(edge_ids, device=dev).to(torch.int64).reshape(-1)
        if edge_ids.numel() and (int(edge_ids.min()) < 0 or int(edge_ids.max()) >= e):
            raise IndexError("edge id out of range")
        slots = g.slot_of_eid[edge_ids].contiguous()
        w = torch.empty((edge_ids.numel(), f, f), dtype=torch.float32, device=dev)
        _lib.call("ampconv_attn_weights_slots_f32", saved.qkv, saved.lse, g.slot_dst, g.dst_src, slots,
                  _lib.i64(slots.numel()), w, _lib.i32(f), _lib.i32(d), _lib.i32(h), _stream(dev))
    return w


def attention_weights_chunks(saved, chunk_edges=4096):
    """Generator over (edge_ids, weights[len, F, F]) slices covering every edge in original order."""
    e = saved.shape[1]
    dev = saved.graph.device
    for lo in range(0, e, chunk_edges):
        ids = torch.arange(lo, min(e, lo + chunk_edges), device=dev)
        yield ids, attention_weights(saved, ids)


def edge_output(saved, w_out, b_out):
    """Per-edge attention output after out_proj [E, F, d] in original edge order (``attn_output``)."""
    saved.ensure_fp32_views()
    n, e, f, d, h = saved.shape
    g = saved.graph
    dev = saved.qkv.device
    o = torch.empty((e, f, d), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.call("ampconv_edge_output_f32", saved.qkv, saved.lse, g.dst_rowptr, g.dst_src, g.dst_eid,
                  w_out, b_out, o, _lib.i64(n), _lib.i64(e), _lib.i32(f), _lib.i32(d), _lib.i32(h), _stream(dev))
    return o


def profile_stages(x, graph, w_in, b_in, w_out, b_out, num_heads, d_out, mode, reps):
    """Per-kernel durations (ms) of the attention kernels, each timed alone with CUDA events on
    the stream it is launched on (used by bench.py for the roofline line)."""
    n, width = x.shape
    d = w_in.shape[1]
    f = width // d
    dev = x.device
    e = graph.num_edges
    rows = n * f
    mode = resolve_mode(mode, f, d, num_heads)
    with torch.cuda.device(dev):
        st = _stream(dev)
        fwd = _forward_bf16 if mode == "bf16" else _forward_fp32
        out, saved = fwd(x, graph, w_in, b_in, w_out, b_out, num_heads)
        d_agg = torch.empty((rows, d), dtype=torch.float32, device=dev)
        d_w_out = torch.empty_like(w_out)
        d_b_out = torch.empty(d, dtype=torch.float32, device=dev)
        ws = _param_grad_ws(3 * d, d, dev)
        _lib.call("ampconv_out_proj_bwd_f32", d_out.contiguous(), saved.agg, w_out, graph.inv_deg, graph.has_in,
                  d_agg, d_w_out, d_b_out, _lib.i64(n), _lib.i32(f), _lib.i32(d), ws, _lib.size_t(ws.numel()), st)
        d_qkv = torch.empty((rows, 3 * d), dtype=torch.float32, device=dev)
        dims = (_lib.i64(n), _lib.i64(e), _lib.i32(f), _lib.i32(d), _lib.i32(num_heads), st)
        calls = {}
        if mode == "bf16":
            q, k, v, lse2, bws = saved.bf16
            calls["attn_fwd"] = lambda: _lib.call(
                "ampconv_attn_fwd_bf16", q, k, v, graph.dst_rowptr, graph.dst_src, graph.inv_deg, graph.order_dst, saved.agg, lse2,
                _lib.i64(n), _lib.i64(e), _lib.i32(f), _lib.i32(d), _lib.i32(num_heads), bws,
                _lib.size_t(bws.numel() * 4), st)
        else:
            calls["attn_fwd"] = lambda: _lib.call("ampconv_attn_fwd_f32", saved.qkv, graph.dst_rowptr, graph.dst_src,
                                                  graph.inv_deg, saved.agg, saved.lse, *dims)
        if mode == "bf16":
            q, k, v, lse2, bws = saved.bf16
            d_agg16 = d_agg.to(torch.bfloat16)
            delta2 = torch.empty_like(lse2)
            tail = (_lib.i64(n), _lib.i64(e), _lib.i32(f), _lib.i32(d), _lib.i32(num_heads), bws,
                    _lib.size_t(bws.numel() * 4), st)
            d_qkv16 = torch.empty((rows, 3 * d), dtype=torch.bfloat16, device=dev)     # the product path's bf16 gradient rows
            calls["attn_bwd_dq"] = lambda: _lib.call("ampconv_attn_bwd_dq_bf16_h", q, k, v, d_agg16, lse2,
                                                     graph.dst_rowptr, graph.dst_src, graph.order_dst, d_qkv16, delta2, *tail)
            calls["attn_bwd_dkv"] = lambda: _lib.call("ampconv_attn_bwd_dkv_bf16_h", q, k, v, d_agg16, lse2, delta2,
                                                      graph.src_rowptr, graph.src_dst, graph.src_pos, graph.order_src, d_qkv16, *tail)
        else:
            delta = torch.empty_like(saved.lse)
            calls["attn_bwd_dq"] = lambda: _lib.call("ampconv_attn_bwd_dq_f32", saved.qkv, d_agg, saved.lse,
                                                     graph.dst_rowptr, graph.dst_src, d_qkv, delta, *dims)
            calls["attn_bwd_dkv"] = lambda: _lib.call("ampconv_attn_bwd_dkv_f32", saved.qkv, d_agg, saved.lse, delta,
                                                      graph.src_rowptr, graph.src_dst, graph.src_pos, d_qkv, *dims)
        result = {}
        for name, fn in calls.items():
            fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            result[name] = e0.elapsed_time(e1) / reps
    return result


def bf16_status(saved):
    """Status word of the tensor-core kernels of a bf16-mode forward (0 = ok); synchronises."""
    ws = saved.bf16[4]
    status = ctypes.c_int(0)
    with torch.cuda.device(ws.device):
        _lib.call("ampconv_bf16_status", ws, ctypes.byref(status), _stream(ws.device))
    return status.value
