"""``AMPNetClassifier`` -- the reference's second caller of the hot path
(``src/ampnet/module/amp_net_classifier_Rahul.py:7-57``): two ``AMPConv`` layers over pre-built feature tokens
``data.x [N, F*d]``, ELU and dropout(0.6) between them, a linear read-out over the flattened tokens, log-softmax.

Same constructor arguments, attribute and parameter names as the reference class, including the members its forward never
uses (``layer_norm``, ``post_conv_linear1``, ``post_conv_linear2``: they are part of the reference's state_dict, so its
checkpoints load with ``strict=True``).  The two message-passing layers are ``ampnet_b200.AMPConv`` (CUDA only); the glue
(dropout, ELU, read-out) is plain torch, as in the reference.  ``mode`` selects the kernel family of both layers
("auto" | "bf16" | "fp32", see ``ampnet_b200/conv/amp_conv.py``).
"""
import torch
import torch.nn.functional as F
from torch.nn import LayerNorm, Linear

from ..conv import AMPConv

__all__ = ["AMPNetClassifier"]


class AMPNetClassifier(torch.nn.Module):
    def __init__(self, num_heads, embed_dim, n_original_features, out_dim, mode="auto"):
        super().__init__()
        self.conv1_embedding = None
        self.conv2_embedding = None
        self.num_heads = num_heads
        self.embed_dim = embed_dim
        self.out_dim = out_dim
        width = n_original_features * embed_dim
        self.layer_norm = LayerNorm(width, elementwise_affine=False)          # amp_net_classifier_Rahul.py:17-20 (unused)
        self.conv1 = AMPConv(embed_dim=embed_dim, num_heads=num_heads, mode=mode)
        self.post_conv_linear1 = Linear(in_features=width, out_features=width)   # :26-29 (unused by forward)
        self.conv2 = AMPConv(embed_dim=embed_dim, num_heads=num_heads, mode=mode)
        self.post_conv_linear2 = Linear(in_features=width, out_features=width)   # :35-38 (unused by forward)
        self.linear_out = Linear(in_features=width, out_features=out_dim)

    def forward(self, data):
        # amp_net_classifier_Rahul.py:45-57
        x, edge_index = data.x, data.edge_index
        x = F.dropout(x, p=0.6, training=self.training)
        x = self.conv1(x, edge_index)
        self.conv1_embedding = x
        x = F.elu(x)
        x = F.dropout(x, p=0.6, training=self.training)
        x = self.conv2(x, edge_index)          # same edge_index: the graph views built for conv1 are reused
        self.conv2_embedding = x
        x = F.elu(x)
        x = F.dropout(x, p=0.6, training=self.training)
        x = self.linear_out(x)
        return F.log_softmax(x, dim=1)
