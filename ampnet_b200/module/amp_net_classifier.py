"""``AMPNetClassifier`` -- the reference's second caller of the hot path
(``src/ampnet/module/amp_net_classifier_Rahul.py:7-57``): two ``AMPConv`` layers over pre-built feature tokens
``data.x [N, F*d]``, each preceded by dropout(0.6) and followed by ELU, then dropout, a linear read-out over the flattened
tokens and log-softmax.

Constructor arguments, attribute names and state_dict keys are the reference's, including the members its forward never
uses (``layer_norm``, ``post_conv_linear1``, ``post_conv_linear2``: they are part of the reference's checkpoints, which
therefore load with ``strict=True``).  The two message-passing layers are ``ampnet_b200.AMPConv`` (CUDA only); the glue is
plain torch, as in the reference.  ``mode`` selects the kernel family of both layers ("auto" | "bf16" | "fp32").
"""
import torch.nn.functional as F
from torch import nn

from ..conv import AMPConv

__all__ = ["AMPNetClassifier"]


class AMPNetClassifier(nn.Module):
    DROPOUT = 0.6          # hard-wired in the reference's forward (amp_net_classifier_Rahul.py:47,51,55)
    NUM_LAYERS = 2

    def __init__(self, num_heads, embed_dim, n_original_features, out_dim, mode="auto"):
        super().__init__()
        self.num_heads, self.embed_dim, self.out_dim = num_heads, embed_dim, out_dim
        width = n_original_features * embed_dim
        self.layer_norm = nn.LayerNorm(width, elementwise_affine=False)      # no parameters; kept for attribute parity
        for i in range(1, self.NUM_LAYERS + 1):
            # registration order conv_i, post_conv_linear_i = the reference's (same state_dict order, same init draws)
            self.add_module(f"conv{i}", AMPConv(embed_dim=embed_dim, num_heads=num_heads, mode=mode))
            self.add_module(f"post_conv_linear{i}", nn.Linear(width, width))
            setattr(self, f"conv{i}_embedding", None)
        self.linear_out = nn.Linear(width, out_dim)

    def forward(self, data):
        h, edge_index = data.x, data.edge_index           # the same edge_index for both layers: graph views built once
        for i in range(1, self.NUM_LAYERS + 1):
            h = F.dropout(h, p=self.DROPOUT, training=self.training)
            h = getattr(self, f"conv{i}")(h, edge_index)
            setattr(self, f"conv{i}_embedding", h)        # read by the reference's analysis code
            h = F.elu(h)
        h = F.dropout(h, p=self.DROPOUT, training=self.training)
        return F.log_softmax(self.linear_out(h), dim=1)
