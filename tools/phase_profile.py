"""Phase timers of the tcgen05 forward kernel (debug entry point): where one softmax warp spends its cycles."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ampnet_b200 import _lib, functional as F_
from ampnet_b200.graph import Graph
from oracle import cases

def main():
    dev = torch.device("cuda:0")
    n, e, f, d, h = 16934, 116624, 128, 64, 4
    ei = torch.from_numpy(cases.make_graph("uniform", n, e, seed=7)).to(dev)
    g = Graph(ei, n)
    rows = n * f
    q = torch.randn(rows, d, device=dev).to(torch.bfloat16) * 0.5
    k = torch.randn(rows, d, device=dev).to(torch.bfloat16)
    v = torch.randn(rows, d, device=dev).to(torch.bfloat16)
    agg = torch.empty(rows, d, device=dev)
    lse2 = torch.empty(e, h, f, device=dev)
    ws = torch.zeros(64, dtype=torch.int32, device=dev)
    prof = torch.zeros(16, dtype=torch.int64, device=dev)
    st = _lib.stream_ptr(torch.cuda.current_stream(dev))
    for _ in range(2):
        _lib.call("ampconv_attn_fwd_bf16_profile", q, k, v, g.dst_rowptr, g.dst_src, g.inv_deg, None, agg, lse2,
                  _lib.i64(n), _lib.i64(e), _lib.i32(f), _lib.i32(d), _lib.i32(h), ws, _lib.size_t(256), st, prof)
    torch.cuda.synchronize()
    p = prof.cpu().tolist()
    names = ["wait S", "load S", "max", "wait P slot", "exp+pack+scale", "st P + publish + lse", "wait O (node end)", "wait MUFU token",
             "node wait", "node epilogue"]
    items = max(1, p[10])
    tot = sum(p[:10])
    print(f"items={items} total cycles/item={tot / items:.0f}")
    for nm, c in zip(names, p[:10]):
        print(f"  {nm:18s} {c / items:8.1f} cyc/item  {100.0 * c / tot:5.1f}%")
    mi = max(1, p[15])
    if p[15] > 0: print(f"MMA thread: items={mi} total/item={p[12] / mi:.0f} issue-QK/item={p[13] / mi:.0f} issue-PV/item={p[14] / mi:.0f} "
          f"idle/item={(p[12] - p[13] - p[14]) / mi:.0f}")

def bwd():
    dev = torch.device("cuda:0")
    n, e, f, d, h = 16934, 116624, 128, 64, 4
    ei = torch.from_numpy(cases.make_graph("uniform", n, e, seed=7)).to(dev)
    g = Graph(ei, n)
    rows = n * f
    q = torch.randn(rows, d, device=dev).to(torch.bfloat16) * 0.5
    k = torch.randn(rows, d, device=dev).to(torch.bfloat16)
    v = torch.randn(rows, d, device=dev).to(torch.bfloat16)
    do = torch.randn(rows, d, device=dev).to(torch.bfloat16)
    agg = torch.empty(rows, d, device=dev)
    lse2 = torch.empty(e, h, f, device=dev)
    delta = torch.empty(e, h, f, device=dev)
    dqkv = torch.empty(rows, 3 * d, device=dev)
    ws = torch.zeros(64, dtype=torch.int32, device=dev)
    st = _lib.stream_ptr(torch.cuda.current_stream(dev))
    tail = (_lib.i64(n), _lib.i64(e), _lib.i32(f), _lib.i32(d), _lib.i32(h), ws, _lib.size_t(256), st)
    _lib.call("ampconv_attn_fwd_bf16", q, k, v, g.dst_rowptr, g.dst_src, g.inv_deg, None, agg, lse2, *tail)
    prof = torch.zeros(64, dtype=torch.int64, device=dev)
    lib = _lib.load()
    names = ["own tiles", "edge tiles + statistics", "scores X/Y", "accumulators (node end)", "delta partials", "P K ring"]
    for mode in ("dq", "dkv"):
        prof.zero_()
        lib.ampconv_debug_set_bwd_profile(ctypes.c_void_p(prof.data_ptr()))
        if mode == "dq":
            _lib.call("ampconv_attn_bwd_dq_bf16", q, k, v, do, lse2, g.dst_rowptr, g.dst_src, None, dqkv, delta, *tail)
        else:
            _lib.call("ampconv_attn_bwd_dkv_bf16", q, k, v, do, lse2, delta, g.src_rowptr, g.src_dst, g.src_pos, None, dqkv, *tail)
        torch.cuda.synchronize()
        lib.ampconv_debug_set_bwd_profile(ctypes.c_void_p(0))
        p = prof.cpu().tolist()
        for off, who in ((0, "elementwise warp 0 (group 0)"), (16, "elementwise warp 8 (group 1)")):
            items = max(1, p[off + 8])
            print(f"bwd {mode} {who}: items={items} total cycles/item={p[off + 6] / items:.0f}; blocked per item on: " +
                  ", ".join(f"{nm} {p[off + i] / items:.0f}" for i, nm in enumerate(names)) +
                  f"; busy {(p[off + 6] - sum(p[off:off + 6])) / items:.0f}")
        for off, who, nm in ((32, "X-side consumer warp", ["accumulators free", "ring slot free", "operands"]),
                             (40, "Y-side consumer warp", ["accumulators free", "ring slot free", "operands"]),
                             (48, "fold warp 0", ["delta partials", "P K slot", "accumulators", "own tiles"])):
            if p[off + 7] > 0:
                it = p[off + 7]
                print(f"    {who}: total/item {p[off + 6] / it:.0f}; blocked on: " +
                      ", ".join(f"{n_} {p[off + i] / it:.0f}" for i, n_ in enumerate(nm)) +
                      f"; busy {(p[off + 6] - sum(p[off:off + len(nm)])) / it:.0f}")


if __name__ == "__main__":
    import ctypes
    bwd()
    main()
