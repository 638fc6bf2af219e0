"""Flat-name shim: put ee_semantic_segmentation_b200/dropin on sys.path and code written against the
reference's module `ee_dnn_op` (e.g. pickled models referring to `from_deepv3.branchyDeepv3`) resolves to the
B200 implementation."""
from ee_semantic_segmentation_b200.ee_dnn_op import *  # noqa: F401,F403
