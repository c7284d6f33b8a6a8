"""ampnet_b200 -- B200-native (sm_100a) implementation of AMPNet's AMPConv hot path.

Public surface mirrors the reference package (``src/ampnet/__init__.py``): ``AMPConv`` and the models built on it
(``AMPGCN``, ``AMPNetClassifier``).
Importing the package does not need a GPU; running a layer does, and fails loudly otherwise.
"""
from .conv import AMPConv, AMPConvV2
from .module import AMPGCN, AMPNetClassifier

__all__ = ["AMPConv", "AMPConvV2", "AMPGCN", "AMPNetClassifier"]
__version__ = "0.1.0"
