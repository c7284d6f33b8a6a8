"""B200 mirror of the reference's 2-layer model ``AMPGCN`` (``src/ampnet/module/amp_gcn.py:19-276``): same constructor
keywords, same sub-module / parameter names (checkpoints written by the reference scripts load with ``strict=True``),
same ``forward(data)`` and the same public attributes (``conv1_embedding``, ``conv2_embedding``,
``sampled_node_feat_indices``).  What changes:

* the two message-passing layers are ``ampnet_b200.AMPConv`` (CUDA, no CPU path);
* the tokeniser (``amp_gcn.py:120-183``) runs on the device: column z-score with scikit-learn's ``StandardScaler``
  semantics (population variance, zero-variance columns keep scale 1) instead of a device -> host -> device round trip,
  and one batched draw of the present features of every node instead of a Python loop over nodes
  (0.39 s of the reference's forward at Cora size).  Bit parity with ``np.random.choice`` is not
  possible, so ``forward`` accepts the indices explicitly (that is how the parity test feeds the reference's draw);
* ``dropout_adj`` is restated here (PyG is not a dependency).

Plotting helpers of the reference class (``amp_gcn.py:278-406``) are out of scope."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..conv import AMPConv


def dropout_adj(edge_index, p=0.5, training=True):
    """Drops every edge independently with probability p (training only); returns (edge_index, None) like PyG's
    ``torch_geometric.utils.dropout.dropout_adj`` as called at ``amp_gcn.py:241``."""
    if p < 0.0 or p > 1.0:
        raise ValueError(f"Dropout probability has to be between 0 and 1 (got {p})")
    if not training or p == 0.0:
        return edge_index, None
    keep = torch.rand(edge_index.size(1), device=edge_index.device) >= p
    return edge_index[:, keep], None


class AMPGCN(nn.Module):
    def __init__(self, device="cuda", embedding_dim=100, num_heads=2, num_node_features=1433, num_sampled_vectors=40,
                 output_dim=7, softmax_out=True, feat_emb_dim=99, val_emb_dim=1, downsample_feature_vectors=True,
                 average_pooling_flag=True, dropout_rate=0.1, dropout_adj_rate=0.1, feature_repeats=5, mode="auto"):
        super().__init__()
        assert embedding_dim == feat_emb_dim + val_emb_dim, \
            "Feature and value dimensions do not add up to total embedding dimension"
        if val_emb_dim != 1:
            raise ValueError("the tokeniser appends the z-scored value as ONE dimension (amp_gcn.py:145-146)")
        self.device = device
        self.conv1_embedding = None
        self.conv2_embedding = None
        self.emb_dim = embedding_dim
        self.num_sampled_vectors = num_sampled_vectors
        self.num_node_features = num_node_features
        self.output_dim = output_dim
        self.softmax_out = softmax_out
        self.feat_emb_dim = feat_emb_dim
        self.val_emb_dim = val_emb_dim
        self.downsampling_vectors = downsample_feature_vectors
        self.average_pooling_flag = average_pooling_flag
        self.dropout_rate = dropout_rate
        self.dropout_adj_rate = dropout_adj_rate
        self.feature_repeats = feature_repeats
        self.feature_embedding_table = nn.Embedding(num_embeddings=num_node_features, embedding_dim=feat_emb_dim)
        self._sampled_idx = None
        if not average_pooling_flag:
            self.cls_token = nn.Parameter(torch.zeros(1, 1, self.emb_dim))
            nn.init.normal_(self.cls_token, std=.02)
        self.conv1 = AMPConv(embed_dim=embedding_dim, num_heads=num_heads, mode=mode)
        self.drop1 = nn.Dropout(p=dropout_rate)
        self.conv2 = AMPConv(embed_dim=embedding_dim, num_heads=num_heads, mode=mode)
        self.drop2 = nn.Dropout(p=dropout_rate)
        self.final_linear_out = nn.Linear(in_features=self.emb_dim, out_features=output_dim)
        self.drop3 = nn.Dropout(p=dropout_rate)
        self.act_out = nn.Sigmoid()

    # ------------------------------------------------------------------ tokeniser (amp_gcn.py:120-183), on the device
    @property
    def sampled_node_feat_indices(self):
        """[N, S] numpy array of the sampled feature ids, as the reference stores it (amp_gcn.py:244); None without sampling."""
        return None if self._sampled_idx is None else self._sampled_idx.cpu().numpy()

    @staticmethod
    def _zscore(x):
        # sklearn.preprocessing.StandardScaler: per-column mean, population variance, scale 1 where the variance is ~0
        x64 = x.to(torch.float64)
        mean = x64.mean(dim=0, keepdim=True)
        var = x64.var(dim=0, unbiased=False, keepdim=True)
        scale = var.sqrt()
        scale = torch.where(scale < 10 * torch.finfo(torch.float64).eps, torch.ones_like(scale), scale)
        return ((x64 - mean) / scale).to(torch.float32)

    def normalize_features_and_add_feature_table_embedding(self, x, sampled_indices=None):
        z = self._zscore(x)
        n = x.shape[0]
        table = self.feature_embedding_table.weight
        if self.downsampling_vectors:
            if sampled_indices is None:
                present = (x != 0).to(torch.float32)
                if bool((present.sum(dim=1) == 0).any()):
                    raise ValueError("a node has no present feature to sample from (np.random.choice raises as well)")
                idx = torch.multinomial(present, self.num_sampled_vectors, replacement=True)
            else:
                idx = torch.as_tensor(sampled_indices, device=x.device).long()
                if tuple(idx.shape) != (n, self.num_sampled_vectors):
                    raise ValueError(f"sampled_indices must be [{n}, {self.num_sampled_vectors}]")
            tokens = torch.cat((table[idx], z.gather(1, idx).unsqueeze(-1)), dim=2)          # [N, S, d]
        else:
            idx = None
            emb = torch.tile(table, (self.feature_repeats, 1))                                  # [nf * repeats, d - 1]
            tokens = torch.cat((emb.unsqueeze(0).expand(n, -1, -1), z.unsqueeze(-1)), dim=2)    # [N, nf, d]
        return tokens.reshape(n, self.num_sampled_vectors * self.emb_dim), idx

    # ------------------------------------------------------------------ forward (amp_gcn.py:239-276)
    def forward(self, data, sampled_indices=None):
        x, edge_index = data.x.to(self.device), data.edge_index.to(self.device)
        edge_index = dropout_adj(edge_index=edge_index, p=self.dropout_adj_rate, training=self.training)[0]
        x, idx = self.normalize_features_and_add_feature_table_embedding(x.to(torch.float32), sampled_indices)
        self._sampled_idx = idx
        x = self.drop1(x)
        x = self.conv1(x, edge_index)
        self.conv1_embedding = x
        x = F.relu(x)
        x = self.drop2(x)
        x = self.conv2(x, edge_index)
        self.conv2_embedding = x
        x = F.relu(x)
        x = self.drop3(x)
        x = torch.reshape(x, (x.shape[0], x.shape[1] // self.emb_dim, self.emb_dim))
        x = x.mean(dim=1) if self.average_pooling_flag else x[:, 0]
        x = self.final_linear_out(x)
        return F.log_softmax(x, dim=1) if self.softmax_out else self.act_out(x)
