"""GraphSAINT random-walk subgraph sampler without torch_geometric / torch_sparse, runnable on the GPU.

Restates the sampler the reference trains with (``experiments/cora_benchmark_graphsaint.py:81-82``:
``GraphSAINTRandomWalkSampler(all_data, batch_size=8, walk_length=150, num_steps=200, sample_coverage=100)``), whose
semantics the reference vendors at ``visualization/visualize_graphsaint_subgraphs.py:22-199``:

* a batch = ``batch_size`` uniform start nodes, each followed for ``walk_length`` steps along out-edges
  (``adj.random_walk``, ``:195-199``; a node without out-edges stays where it is);
* the mini-batch is the subgraph INDUCED by the unique visited nodes, relabelled in ascending node-id order
  (``saint_subgraph``, ``:107-135``); node-level tensors are sliced by the node ids, edge-level tensors by the edge ids;
* with ``sample_coverage > 0`` normalisation coefficients are estimated up front from repeated sampling
  (``__compute_norm__``, ``:137-173``): ``node_norm = num_samples / node_count / N`` (0 counts -> 0.1) and
  ``edge_norm = clamp(node_count[row] / edge_count, 0, 1e4)`` (NaN -> 0.1).

Everything is batched tensor code, so the loader runs where the graph lives (CPU tensors work as well; the CPU tests use
that).  Randomness comes from a ``torch.Generator``: the draws differ from PyG's, the distribution is the same."""
import torch


class SubgraphData:
    """Minimal stand-in for ``torch_geometric.data.Data``: attribute bag with ``to(device)``, ``num_nodes``, ``num_edges``."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def to(self, device):
        return SubgraphData(**{k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in self.__dict__.items()})

    @property
    def num_edges(self):
        return int(self.edge_index.size(1))

    def keys(self):
        return list(self.__dict__.keys())


class GraphSAINTRandomWalkSampler:
    def __init__(self, data, batch_size, walk_length, num_steps=1, sample_coverage=0, generator=None, log=False):
        assert data.edge_index is not None
        assert not hasattr(data, "node_norm") and not hasattr(data, "edge_norm")
        self.data = data
        self.batch_size, self.walk_length = int(batch_size), int(walk_length)
        self.num_steps, self.sample_coverage = int(num_steps), int(sample_coverage)
        self.N = int(data.num_nodes)
        self.E = int(data.edge_index.size(1))
        ei = data.edge_index
        self.device = ei.device
        self.generator = generator
        # CSR by source node (SparseTensor(row=edge_index[0], col=edge_index[1], value=edge id))
        order = torch.sort(ei[0], stable=True).indices
        self.col = ei[1][order].contiguous()
        self.eid = order.contiguous()
        self.row = ei[0][order].contiguous()
        self.rowptr = torch.zeros(self.N + 1, dtype=torch.long, device=self.device)
        self.rowptr[1:] = torch.cumsum(torch.bincount(ei[0], minlength=self.N), 0)
        self.node_norm = self.edge_norm = None
        if self.sample_coverage > 0:
            self.node_norm, self.edge_norm = self._compute_norm()

    def __len__(self):
        return self.num_steps

    # ------------------------------------------------------------------ sampling
    def _rand(self, n):
        return torch.rand(n, device=self.device, generator=self.generator)

    def sample_nodes(self):
        """Node ids visited by ``batch_size`` random walks of ``walk_length`` steps, flattened ([B * (L + 1)])."""
        start = (self._rand(self.batch_size) * self.N).long().clamp_(max=self.N - 1)
        walk = [start]
        cur = start
        for _ in range(self.walk_length):
            lo = self.rowptr[cur]
            deg = self.rowptr[cur + 1] - lo
            pick = lo + (self._rand(cur.numel()) * deg).long().clamp_(max=(deg - 1).clamp_(min=0))
            nxt = torch.where(deg > 0, self.col[pick.clamp_(max=max(self.E - 1, 0))], cur)
            walk.append(nxt)
            cur = nxt
        return torch.stack(walk, dim=1).view(-1)

    def induced_subgraph(self, node_idx):
        """(local edge_index [2, e] in CSR order, global edge ids [e]) of the subgraph induced by the sorted unique node_idx."""
        local = torch.full((self.N,), -1, dtype=torch.long, device=self.device)
        local[node_idx] = torch.arange(node_idx.numel(), device=self.device)
        keep = (local[self.row] >= 0) & (local[self.col] >= 0)
        return torch.stack([local[self.row[keep]], local[self.col[keep]]]), self.eid[keep]

    def sample(self):
        node_idx = torch.unique(self.sample_nodes())
        edge_index, edge_idx = self.induced_subgraph(node_idx)
        return node_idx, edge_index, edge_idx

    def _collate(self, node_idx, edge_index, edge_idx):
        out = {"num_nodes": int(node_idx.numel()), "edge_index": edge_index}
        for key, item in self.data.__dict__.items():
            if key in ("edge_index", "num_nodes"):
                continue
            if isinstance(item, torch.Tensor) and item.dim() > 0 and item.size(0) == self.N:
                out[key] = item[node_idx]
            elif isinstance(item, torch.Tensor) and item.dim() > 0 and item.size(0) == self.E:
                out[key] = item[edge_idx]
            else:
                out[key] = item
        if self.sample_coverage > 0:
            out["node_norm"] = self.node_norm[node_idx]
            out["edge_norm"] = self.edge_norm[edge_idx]
        return SubgraphData(**out)

    def __iter__(self):
        for _ in range(self.num_steps):
            yield self._collate(*self.sample())

    # ------------------------------------------------------------------ normalisation statistics
    def _compute_norm(self):
        node_count = torch.zeros(self.N, dtype=torch.float, device=self.device)
        edge_count = torch.zeros(self.E, dtype=torch.float, device=self.device)
        num_samples = total_sampled_nodes = 0
        while total_sampled_nodes < self.N * self.sample_coverage:
            for _ in range(self.num_steps):
                node_idx, _, edge_idx = self.sample()
                node_count[node_idx] += 1
                edge_count[edge_idx] += 1
                total_sampled_nodes += int(node_idx.numel())
            num_samples += self.num_steps
        t = torch.empty_like(edge_count).scatter_(0, self.eid, node_count[self.row])
        edge_norm = (t / edge_count).clamp_(0, 1e4)
        edge_norm[torch.isnan(edge_norm)] = 0.1
        node_count[node_count == 0] = 0.1
        node_norm = num_samples / node_count / self.N
        return node_norm, edge_norm


def cora_shaped_data(num_nodes=2708, num_features=1433, num_classes=7, num_undirected_edges=5278, nnz_per_node=18,
                     num_train=140, num_val=500, num_test=1000, seed=0, device="cpu"):
    """Synthetic stand-in with the shape of Planetoid Cora (the dataset itself cannot be downloaded here; SURVEY.md
    section 8c): binary bag-of-words features with at least one present feature per node, a symmetrised edge list
    (10 556 directed edges), class labels that correlate with the features (so that a model can learn something) and the
    140 / 500 / 1000 split."""
    g = torch.Generator().manual_seed(seed)
    y = torch.randint(0, num_classes, (num_nodes,), generator=g)
    # every class prefers its own slice of the vocabulary
    block = num_features // num_classes
    x = torch.zeros(num_nodes, num_features)
    own = (torch.rand(num_nodes, nnz_per_node, generator=g) < 0.7)
    col_own = y.unsqueeze(1) * block + torch.randint(0, block, (num_nodes, nnz_per_node), generator=g)
    col_any = torch.randint(0, num_features, (num_nodes, nnz_per_node), generator=g)
    x.scatter_(1, torch.where(own, col_own, col_any), 1.0)
    # homophilous edges: 80 % inside the class
    a = torch.randint(0, num_nodes, (num_undirected_edges,), generator=g)
    same = torch.rand(num_undirected_edges, generator=g) < 0.8
    perm = torch.argsort(y + torch.rand(num_nodes, generator=g))              # nodes grouped by class
    start = torch.searchsorted(y[perm].contiguous(), torch.arange(num_classes))
    cnt = torch.bincount(y, minlength=num_classes)
    pick = start[y[a]] + (torch.rand(num_undirected_edges, generator=g) * cnt[y[a]]).long().clamp_(max=num_nodes - 1)
    b = torch.where(same, perm[pick.clamp_(max=num_nodes - 1)], torch.randint(0, num_nodes, (num_undirected_edges,), generator=g))
    edge_index = torch.cat([torch.stack([a, b]), torch.stack([b, a])], dim=1)
    order = torch.randperm(num_nodes, generator=g)
    masks = {}
    o = 0
    for name, n in (("train_mask", num_train), ("val_mask", num_val), ("test_mask", num_test)):
        m = torch.zeros(num_nodes, dtype=torch.bool)
        m[order[o:o + n]] = True
        masks[name] = m
        o += n
    return SubgraphData(x=x, y=y, edge_index=edge_index, num_nodes=num_nodes, **masks).to(device)
