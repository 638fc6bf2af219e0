"""Flat-name shim for the reference's `my_pixelwise_xentropy`."""
from ee_semantic_segmentation_b200.my_pixelwise_xentropy import BrXEntropyLoss, _cross_entropy  # noqa: F401
