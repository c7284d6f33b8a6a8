"""Destination-/source-sorted views of an ``edge_index`` built on the device once and reused by
both layers and by forward and backward (the reference redoes ``index_select``/scatter every call
through PyG ``propagate``; ``src/ampnet/conv/amp_conv.py:24-26``)."""
import ctypes
import weakref

import torch

from . import _lib


class Graph:
    """CSR by destination and by source for edge_index [2, E] (row 0 = source, row 1 = destination)."""

    def __init__(self, edge_index, num_nodes):
        if edge_index.dim() != 2 or edge_index.size(0) != 2:
            raise ValueError("edge_index must have shape [2, E]")
        if edge_index.dtype != torch.int64:
            raise TypeError("edge_index must be int64 (torch.long), like the reference's")
        if not edge_index.is_cuda:
            raise TypeError("ampnet_b200 runs on CUDA only: edge_index must be a CUDA tensor "
                            "(the CPU oracle lives in oracle/, not in the product)")
        edge_index = edge_index.contiguous()
        dev = edge_index.device
        e, n = edge_index.size(1), int(num_nodes)
        self.num_edges, self.num_nodes, self.device = e, n, dev
        i32 = dict(dtype=torch.int32, device=dev)
        self.dst_rowptr = torch.empty(n + 1, **i32)
        self.dst_src = torch.empty(e, **i32)
        self.dst_eid = torch.empty(e, **i32)
        self.src_rowptr = torch.empty(n + 1, **i32)
        self.src_dst = torch.empty(e, **i32)
        self.src_pos = torch.empty(e, **i32)
        self.inv_deg = torch.empty(n, dtype=torch.float32, device=dev)
        self.has_in = torch.empty(n, dtype=torch.float32, device=dev)
        nbytes = ctypes.c_size_t(0)
        with torch.cuda.device(dev):
            _lib.call("ampconv_graph_workspace_bytes", _lib.i64(e), _lib.i64(n), ctypes.byref(nbytes))
            ws = torch.empty(max(nbytes.value, 1), dtype=torch.uint8, device=dev)
            _lib.call("ampconv_graph_build", edge_index, _lib.i64(e), _lib.i64(n),
                      self.dst_rowptr, self.dst_src, self.dst_eid,
                      self.src_rowptr, self.src_dst, self.src_pos,
                      self.inv_deg, self.has_in, ws, _lib.size_t(ws.numel()),
                      _lib.stream_ptr(torch.cuda.current_stream(dev)))
        # longest-processing-time-first node orders for the persistent tcgen05 kernels (dynamic scheduler)
        self.order_dst = torch.argsort(self.dst_rowptr[1:] - self.dst_rowptr[:-1], descending=True).to(torch.int32)
        self.order_src = torch.argsort(self.src_rowptr[1:] - self.src_rowptr[:-1], descending=True).to(torch.int32)
        self._slot_dst = None
        self._slot_of_eid = None

    @property
    def slot_dst(self):
        """Destination node of every destination-sorted slot [E] (int32)."""
        if self._slot_dst is None:
            deg = (self.dst_rowptr[1:] - self.dst_rowptr[:-1]).to(torch.int64)
            self._slot_dst = torch.repeat_interleave(torch.arange(self.num_nodes, device=self.device, dtype=torch.int32), deg)
        return self._slot_dst

    @property
    def slot_of_eid(self):
        """Destination-sorted slot of every original edge id [E] (int32): the inverse of dst_eid."""
        if self._slot_of_eid is None:
            inv = torch.empty(self.num_edges, dtype=torch.int32, device=self.device)
            inv[self.dst_eid.to(torch.int64)] = torch.arange(self.num_edges, device=self.device, dtype=torch.int32)
            self._slot_of_eid = inv
        return self._slot_of_eid


_cache = {}


def _evict(key):
    _cache.pop(key, None)


def get_graph(edge_index, num_nodes):
    """Cached on the identity, version and shape of ``edge_index`` (both AMPConv layers of a model
    and every training step on a fixed graph share one build).  The key relies on torch's version counter: a buffer
    rewritten behind torch's back (``.data`` writes, a c10d receive into it, an external kernel) must be followed by
    ``clear_cache()`` -- or build a ``Graph`` explicitly and call ``functional.amp_conv`` with it."""
    key = (edge_index.data_ptr(), edge_index._version, tuple(edge_index.shape), int(num_nodes),
           str(edge_index.device))
    hit = _cache.get(key)
    if hit is not None and hit[0]() is not None:
        _cache[key] = _cache.pop(key)          # most recently used last
        return hit[1]
    g = Graph(edge_index, num_nodes)
    while len(_cache) >= 16:
        _cache.pop(next(iter(_cache)))         # evict the least recently used entry only
    # the weak reference ties the entry's validity to the storage the key's data_ptr came from
    _cache[key] = (weakref.ref(edge_index, lambda _r, k=key: _evict(k)), g)
    return g


def clear_cache():
    _cache.clear()
