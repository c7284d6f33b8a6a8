from .amp_conv import AMPConv, AMPConvV2

__all__ = ["AMPConv", "AMPConvV2"]
