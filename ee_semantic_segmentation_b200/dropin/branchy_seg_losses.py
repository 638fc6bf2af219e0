"""Flat-name shim for the reference's `branchy_seg_losses` (imported there as `BSL`)."""
from ee_semantic_segmentation_b200.branchy_seg_losses import (BrSegLoss, DiceLoss, FocalLoss,  # noqa: F401
                                                              FocalTverskyLoss, JaccardLoss, LovaszSoftmax, TverskyLoss)
