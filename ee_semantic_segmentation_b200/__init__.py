"""ee_semantic_segmentation_b200 — B200-native (sm_100a) implementation of the early-exit
segmentation hot path of MateusGilbert/ee_semantic_segmentation, behind the reference's Python API.

Module names mirror the reference's flat files (`from_deepv3_new`, `ee_dnn_op_ne`, `eval_br_ent`,
`my_pixelwise_xentropy`, `branchy_seg_losses`, `lovaszsoftmax`, `seg_metrics`, `compute_mIoU`,
`eval_mIoU`, ...). `ee_semantic_segmentation_b200/dropin/` holds same-named top-level shims for code
that imports the reference's flat module names. The kernels live in csrc/ and are reached through
the C ABI in include/eeseg.h (ctypes, `_lib.py`); there is no CPU or PyTorch fallback for them."""
__version__ = "0.1.0"
