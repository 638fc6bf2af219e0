"""ctypes binding of libeeseg_b200.so (include/eeseg.h). No CPU fallback: if the library is missing
or a kernel is asked to run on a CPU tensor, the call raises."""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EESEG_LIB") or os.path.join(_HERE, "libeeseg_b200.so")   # EESEG_LIB: A/B builds while tuning
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "eeseg.h")

_lib = None

F32, BF16, U8 = 0, 1, 2

c_i, c_i64, c_f, c_p, c_sz = ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p, ctypes.c_size_t

# name -> (restype, argtypes); must list every symbol include/eeseg.h declares (tests check this)
PROTOTYPES = {
    "eeseg_abi_version": (c_i, []),
    "eeseg_last_error": (ctypes.c_char_p, []),
    "eeseg_launch_count": (c_i64, []),
    "eeseg_confusion_hist": (c_i, [c_p, c_i, c_i, c_p, c_i, c_i, c_i64, c_p, c_i, c_p]),
    "eeseg_exit_accumulate": (c_i, [c_p, c_p, c_p, c_i, c_i, c_i, c_i64, c_p, c_p, c_p, c_p]),
    "eeseg_exit_accumulate_u8": (c_i, [c_p, c_p, c_p, c_i, c_i, c_i, c_i64, c_p, c_p, c_p, c_p]),
    "eeseg_exit_gate_num_partials": (c_i, [c_i, c_i]),
    "eeseg_exit_gate_pixels": (c_i, [c_p, c_i, c_i, c_i64, c_i64, c_i64, c_i64, c_i, c_i, c_i, c_i,
                                     c_i, c_i, c_f, c_p, c_i, c_i64, c_p, c_p, c_p, c_p, c_p, c_p]),
    "eeseg_entropy_pool_mean": (c_i, [c_p, c_i, c_i, c_i, c_i, c_i, c_p, c_p]),
    "eeseg_exit_gate_decide": (c_i, [c_p, c_p, c_i, c_p, c_i, c_i64, c_f, c_i, c_i, c_p, c_p, c_p,
                                     c_p, c_p, c_p]),
    "eeseg_upsample_bilinear": (c_i, [c_p, c_i, c_i64, c_i64, c_i64, c_i64, c_i, c_i, c_i, c_i, c_i,
                                      c_i, c_p, c_i, c_i64, c_p]),
    "eeseg_upsample_bilinear_bwd": (c_i, [c_p, c_i, c_i64, c_i, c_i, c_i, c_i, c_p, c_p]),
    "eeseg_multi_exit_ce_workspace_bytes": (c_sz, [c_i, c_i, c_i64]),
    "eeseg_multi_exit_ce_fwd": (c_i, [c_p, c_i, c_i64, c_p, c_i, c_i, c_i, c_i64, c_i64, c_p, c_p,
                                      c_p, c_p, c_p, c_p]),
    "eeseg_multi_exit_ce_bwd": (c_i, [c_p, c_i, c_i64, c_p, c_i, c_i, c_i, c_i64, c_i64, c_p, c_p,
                                      c_p, c_p]),
    "eeseg_scale_exits": (c_i, [c_p, c_i, c_i64, c_i, c_i64, c_p, c_p, c_p]),
    "eeseg_lovasz_workspace_bytes": (c_sz, [c_i, c_i, c_i, c_i64]),
    "eeseg_lovasz_fwd_bwd": (c_i, [c_p, c_i, c_i64, c_p, c_i, c_i, c_i, c_i64, c_i, c_i64, c_i, c_i,
                                   c_p, c_p, c_p, c_sz, c_p]),
    "eeseg_conv_igemm_fwd": (c_i, [c_p, c_p, c_p, c_p, c_i64, c_i, c_i, c_i, c_i, c_i, c_i, c_i,
                                   c_i, c_i, c_i, c_i, c_p, c_i64, c_p, c_i, c_i64, c_p]),
    "eeseg_conv_igemm_wgrad_workspace_bytes": (c_sz, [c_i, c_i, c_i, c_i, c_i, c_i, c_i]),
    "eeseg_conv_igemm_wgrad": (c_i, [c_p, c_p, c_i64, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_p, c_p]),
    "eeseg_conv_igemm_wgrad_to_param_workspace_bytes": (c_sz, [c_i, c_i, c_i, c_i, c_i, c_i, c_i]),
    "eeseg_conv_igemm_wgrad_to_param": (c_i, [c_p, c_p, c_i64, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_i, c_p, c_p]),
    "eeseg_conv_igemm_dgrad_workspace_bytes": (c_sz, [c_i, c_i, c_i, c_i]),
    "eeseg_conv_igemm_dgrad": (c_i, [c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_i, c_i64, c_p, c_p]),
    "eeseg_conv_weight_rot180_t": (c_i, [c_p, c_i, c_i, c_i, c_i, c_p, c_p]),
    "eeseg_soft_overlap_workspace_bytes": (c_sz, [c_i, c_i, c_i, c_i64]),
    "eeseg_soft_overlap_fwd": (c_i, [c_p, c_i, c_i64, c_p, c_i, c_i, c_i, c_i64, c_p, c_p, c_p]),
    "eeseg_soft_overlap_bwd": (c_i, [c_p, c_i, c_i64, c_p, c_i, c_i, c_i, c_i64, c_p, c_p, c_p, c_p]),
    "eeseg_maxpool3x3s2_nhwc_train": (c_i, [c_p, c_i, c_i, c_i, c_i, c_p, c_p, c_p]),
    "eeseg_maxpool3x3s2_nhwc_bwd": (c_i, [c_p, c_p, c_i, c_i, c_i, c_i, c_p, c_p]),
    "eeseg_exit_stage_commit": (c_i, [c_p, c_f, c_i, c_i, c_i, c_p, c_p, c_i, c_i64, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p]),
    "eeseg_compact_rows": (c_i, [c_p, c_p, c_p, c_p, c_i, c_i, c_i64, c_p]),
    "eeseg_focal_workspace_bytes": (c_sz, [c_i, c_i, c_i64]),
    "eeseg_focal_fwd": (c_i, [c_p, c_i, c_i64, c_p, c_i, c_i, c_i, c_i64, c_f, c_p, c_p, c_i64, c_i64, c_p, c_p, c_p, c_p, c_p, c_p]),
    "eeseg_bn_train_workspace_bytes": (c_sz, [c_i]),
    "eeseg_bn_train_fwd": (c_i, [c_p, c_i64, c_i, c_p, c_p, c_p, c_p, c_f, c_f, c_i, c_p, c_p, c_p, c_p, c_p, c_p]),
    "eeseg_bn_train_bwd": (c_i, [c_p, c_p, c_p, c_i64, c_i, c_p, c_p, c_p, c_i, c_p, c_p, c_p, c_p, c_p, c_p]),
    "eeseg_bn_train_bwd_acc": (c_i, [c_p, c_p, c_p, c_i64, c_i, c_p, c_p, c_p, c_i, c_p, c_p, c_p, c_p, c_p, c_p]),
    "eeseg_conv_group_tiles": (c_i, [c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p]),
    "eeseg_conv_igemm_grouped": (c_i, [c_p, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i,
                                       c_p, c_i64, c_i, c_p, c_i, c_i, c_p]),
    "eeseg_conv_pair_clusters": (c_i, []),
    "eeseg_stem_space_to_depth": (c_i, [c_p, c_i, c_i, c_i, c_p, c_p]),
    "eeseg_stem_space_to_depth_any": (c_i, [c_p, c_i, c_p, c_p, c_i, c_i, c_i, c_p, c_p]),
    "eeseg_maxpool3x3s2_nhwc": (c_i, [c_p, c_i, c_i, c_i, c_i, c_p, c_p]),
    "eeseg_dense_bn_act": (c_i, [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_p, c_p]),
    "eeseg_dropout_fwd": (c_i, [c_p, c_i64, c_f, c_p, c_p, c_p, c_p]),
    "eeseg_dropout_bwd": (c_i, [c_p, c_p, c_i64, c_f, c_p, c_p]),
    "eeseg_sgd_chunk_bytes": (c_sz, []),
    "eeseg_sgd_multi": (c_i, [c_p, c_i, c_p, c_f, c_f, c_p]),
    "eeseg_dense_bwd": (c_i, [c_p, c_p, c_p, c_i, c_i, c_i, c_p, c_p, c_p]),
    "eeseg_weight_prep_tile_bytes": (c_sz, []),
    "eeseg_weight_prep_multi": (c_i, [c_p, c_i, c_i, c_p]),
    "eeseg_bn_rows_fwd": (c_i, [c_p, c_i, c_i, c_p, c_p, c_p, c_p, c_f, c_f, c_i, c_p, c_p, c_p, c_p]),
    "eeseg_bn_rows_bwd": (c_i, [c_p, c_p, c_p, c_i, c_i, c_p, c_p, c_p, c_i, c_p, c_p, c_p, c_p]),
    "eeseg_broadcast_rows_nhwc": (c_i, [c_p, c_i, c_i64, c_i, c_f, c_p, c_p]),
    "eeseg_global_avgpool_workspace_bytes": (c_sz, [c_i, c_i]),
    "eeseg_global_avgpool_nhwc": (c_i, [c_p, c_i, c_i64, c_i, c_p, c_p, c_p]),
}


# present only in the -DEESEG_TUNING build (include/eeseg_tuning.h); bound when the loaded library has them
TUNING_PROTOTYPES = {
    "eeseg_conv_debug_stats": (c_i, [c_p]),
    "eeseg_conv_probe": (c_i, [c_i, c_i]),
    "eeseg_conv_timing": (c_i, [c_p, c_i]),
}


def header_symbols():
    """Every function name declared in include/eeseg.h."""
    with open(HEADER_PATH) as f:
        src = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(eeseg_[a-z0-9_]+)\s*\(", src)))


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m ee_semantic_segmentation_b200.build` "
                "(there is no CPU or PyTorch fallback for the eeseg kernels)")
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        for name, (res, args) in TUNING_PROTOTYPES.items():
            if hasattr(l, name):
                fn = getattr(l, name)
                fn.restype, fn.argtypes = res, args
        if l.eeseg_abi_version() != 1:
            raise RuntimeError("libeeseg_b200.so ABI version mismatch")
        _lib = l
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().eeseg_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def launch_count():
    return int(lib().eeseg_launch_count())
