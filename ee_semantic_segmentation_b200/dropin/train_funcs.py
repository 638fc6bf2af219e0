"""Flat-name shim for the reference's `train_funcs` (train_epoch, train)."""
from ee_semantic_segmentation_b200.train_funcs import make_optimizer, poly_scheduler, train, train_epoch  # noqa: F401
