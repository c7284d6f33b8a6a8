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
    names = ["wait S", "load S", "max", "wait P slot", "exp+pack+scale", "st P + publish + lse", "wait O (node end)", "-",
             "node wait", "node epilogue"]
    items = max(1, p[10])
    tot = sum(p[:10])
    print(f"items={items} total cycles/item={tot / items:.0f}")
    for nm, c in zip(names, p[:10]):
        print(f"  {nm:18s} {c / items:8.1f} cyc/item  {100.0 * c / tot:5.1f}%")
    mi = max(1, p[15])
    if p[15] > 0: print(f"MMA thread: items={mi} total/item={p[12] / mi:.0f} issue-QK/item={p[13] / mi:.0f} issue-PV/item={p[14] / mi:.0f} "
          f"idle/item={(p[12] - p[13] - p[14]) / mi:.0f}")

if __name__ == "__main__":
    main()
