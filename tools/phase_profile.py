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
    names = ["wait X/Y", "compute (ld, exp, pack, st)", "publish", "delta exchange", "scale P + publish", "lse load", "node / accumulator wait",
             "node epilogue"]
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
        items = max(1, p[8])
        tot = sum(p[:8])
        print(f"bwd {mode}: items={items} total cycles/item={tot / items:.0f}")
        for nm, cyc in zip(names, p[:8]):
            print(f"  {nm:28s} {cyc / items:8.1f} cyc/item  {100.0 * cyc / tot:5.1f}%")
        if p[24] > 0:
            it2 = p[24]
            print("  last elementwise warp: " + ", ".join(f"{nm} {cyc / it2:.0f}" for nm, cyc in zip(names, p[16:24])))
        if p[36] > 0:
            it3 = p[36]
            print(f"  score-MMA warp /item: wait edge tiles {p[32] / it3:.0f}, wait set free {p[33] / it3:.0f}, "
                  f"issue+commit {p[34] / it3:.0f}, node wait {p[35] / it3:.0f}")
        if p[44] > 0:
            it4 = p[44]
            print(f"  T-MMA warp 0 /item: wait operands {p[40] / it4:.0f}, issue+commit {p[41] / it4:.0f}, "
                  f"until T complete {p[42] / it4:.0f}, other {p[43] / it4:.0f}")


if __name__ == "__main__":
    import ctypes
    bwd()
    main()
