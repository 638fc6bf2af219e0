"""Flat-name shim for the step-level part of the reference's `train_funcs`."""
from ee_semantic_segmentation_b200.train_funcs import make_optimizer, poly_scheduler, train_epoch  # noqa: F401
